// The two thin convolutions at the waveform boundary of the Oobleck stack.  They hold < 0.3 % of the
// FLOPs but touch the longest tensors, and their channel counts (io_channels = 1 or 2) do not fit a
// tensor-core tile, so they get HBM-streaming CUDA-core kernels of their own:
//   conv_wave_out_kernel  decoder tail: SnakeBeta -> Conv1d(C -> io, k7, pad 3, bias=False) [-> tanh]
//                         (autoencoders.py:183-185); reads the fp32 residual stream [B, T, C], applies
//                         SnakeBeta once per staged element in fp32, accumulates in fp32 and writes the
//                         waveform directly in the API layout [B, io, T].  Keeping this layer in fp32
//                         removes what SURVEY.md (H1) measured as half of the bf16 error budget.
//   conv_wave_in_kernel   encoder head: Conv1d(io -> C, k7, pad 3) (autoencoders.py:133); reads the waveform
//                         in the API layout and writes the fp32 residual stream plus the SnakeBeta-activated
//                         bf16 operand of the first ResidualUnit, both channels-last.
#pragma once
#include <cuda_bf16.h>
#include <cstdint>

#include "conv_umma.cuh"    // snake_beta
#include "elementwise.cuh"  // ld_elem / st_elem

namespace kvae {

struct WaveOutParams {
  const float* x;         // [B, T, Cin] fp32 residual stream
  const float* pro_a;     // SnakeBeta exp(alpha) [Cin]
  const float* pro_inv_b;
  const float* w;         // [7][Cin][COUT] fp32
  void* y;                // [B, COUT, T]
  int y_f32;
  int T, Cin, tanh_out;
};

constexpr int kWaveOutTile = 128;

// grid (ceil(T/128), B), block 128.  Shared: act[Cin][135] (+ weights [7][Cin][COUT], partials [4][128][COUT]).
template <int COUT>
__global__ void __launch_bounds__(128) conv_wave_out_kernel(const WaveOutParams p) {
  constexpr int TT = kWaveOutTile, RS = TT + 6, RSP = RS + 1;   // 135: odd stride -> conflict-free both ways
  extern __shared__ float sm_wo[];
  float* act = sm_wo;                              // [Cin][RSP]
  float* ws = act + p.Cin * RSP;                   // [7][Cin][COUT]
  float* part = ws + 7 * p.Cin * COUT;             // [4][TT][COUT]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * TT;
  for (int i = threadIdx.x; i < 7 * p.Cin * COUT; i += 128) ws[i] = p.w[i];
  // stage SnakeBeta(x) for rows t0-3 .. t0+TT+2; lanes walk channels (coalesced 128 B per row segment)
  for (int r = warp; r < RS; r += 4) {
    const int t = t0 - 3 + r;
    const bool in = (t >= 0 && t < p.T);
    const float* xr = p.x + (static_cast<size_t>(b) * p.T + (in ? t : 0)) * p.Cin;
    for (int ci = lane; ci < p.Cin; ci += 32) {
      float v = 0.f;
      if (in) v = snake_beta<true>(xr[ci], p.pro_a[ci], p.pro_inv_b[ci]);
      act[ci * RSP + r] = v;
    }
  }
  __syncthreads();
  // each warp reduces a quarter of the channels for all 128 outputs; lane owns outputs lane + 32*j
  float acc[4][COUT];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[j][c] = 0.f;
  const int cpw = p.Cin / 4;
  for (int ci = warp * cpw; ci < (warp + 1) * cpw; ++ci) {
    const float* a = act + ci * RSP + lane;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      float wv[COUT];
#pragma unroll
      for (int c = 0; c < COUT; ++c) wv[c] = ws[(k * p.Cin + ci) * COUT + c];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xv = a[32 * j + k];
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[j][c] = fmaf(xv, wv[c], acc[j][c]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < COUT; ++c) part[(warp * TT + lane + 32 * j) * COUT + c] = acc[j][c];
  __syncthreads();
  for (int i = threadIdx.x; i < TT * COUT; i += 128) {
    const int c = i / TT, tt = i % TT;            // consecutive threads -> consecutive time (coalesced store)
    const int t = t0 + tt;
    if (t >= p.T) continue;
    float v = part[(0 * TT + tt) * COUT + c] + part[(1 * TT + tt) * COUT + c] + part[(2 * TT + tt) * COUT + c] +
              part[(3 * TT + tt) * COUT + c];
    if (p.tanh_out) v = tanhf(v);
    st_elem(p.y, (static_cast<size_t>(b) * COUT + c) * p.T + t, p.y_f32, v);
  }
}

inline size_t wave_out_smem(int Cin, int Cout) {
  return (static_cast<size_t>(Cin) * (kWaveOutTile + 7) + 7 * Cin * Cout + 4 * kWaveOutTile * Cout) * sizeof(float);
}

struct WaveInParams {
  const void* x;          // [B, CIN, T]
  int x_f32;
  const float* w;         // [7][CIN][Cout] fp32
  const float* bias;      // [Cout]
  float* out_raw;         // [B, T, Cout] fp32 or nullptr
  __nv_bfloat16* out_act; // [B, T, Cout] bf16 or nullptr
  const float* snake_a;   // epilogue SnakeBeta of the first ResidualUnit (or nullptr: plain cast)
  const float* snake_inv_b;
  int T, Cout;
};

constexpr int kWaveInTile = 64;

// grid (ceil(T/64), Cout/128, B), block 128: warp w handles rows w, w+4, ...; lane handles 4 out-channels.
template <int CIN>
__global__ void __launch_bounds__(128) conv_wave_in_kernel(const WaveInParams p) {
  constexpr int TT = kWaveInTile, RS = TT + 6;
  __shared__ float xs[CIN][RS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * TT;
  const int co = blockIdx.y * 128 + lane * 4;
  for (int i = threadIdx.x; i < CIN * RS; i += 128) {
    const int c = i / RS, r = i % RS;
    const int t = t0 - 3 + r;
    xs[c][r] = (t >= 0 && t < p.T) ? ld_elem(p.x, (static_cast<size_t>(b) * CIN + c) * p.T + t, p.x_f32) : 0.f;
  }
  float w[7][CIN][4];
#pragma unroll
  for (int k = 0; k < 7; ++k)
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p.w + (static_cast<size_t>(k) * CIN + c) * p.Cout + co));
      w[k][c][0] = v.x; w[k][c][1] = v.y; w[k][c][2] = v.z; w[k][c][3] = v.w;
    }
  const float4 bias = __ldg(reinterpret_cast<const float4*>(p.bias + co));
  float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sib = sa;
  if (p.snake_a) {
    sa = __ldg(reinterpret_cast<const float4*>(p.snake_a + co));
    sib = __ldg(reinterpret_cast<const float4*>(p.snake_inv_b + co));
  }
  __syncthreads();
  for (int r = warp; r < TT; r += 4) {
    const int t = t0 + r;
    if (t >= p.T) break;
    float v0 = bias.x, v1 = bias.y, v2 = bias.z, v3 = bias.w;
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        const float xv = xs[c][r + k];
        v0 = fmaf(xv, w[k][c][0], v0); v1 = fmaf(xv, w[k][c][1], v1);
        v2 = fmaf(xv, w[k][c][2], v2); v3 = fmaf(xv, w[k][c][3], v3);
      }
    const size_t o = (static_cast<size_t>(b) * p.T + t) * p.Cout + co;
    if (p.out_raw) *reinterpret_cast<float4*>(p.out_raw + o) = make_float4(v0, v1, v2, v3);
    if (p.out_act) {
      if (p.snake_a) {
        v0 = snake_beta<true>(v0, sa.x, sib.x); v1 = snake_beta<true>(v1, sa.y, sib.y);
        v2 = snake_beta<true>(v2, sa.z, sib.z); v3 = snake_beta<true>(v3, sa.w, sib.w);
      }
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1), h1 = __floats2bfloat162_rn(v2, v3);
      *reinterpret_cast<uint2*>(p.out_act + o) =
          make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
    }
  }
}

}  // namespace kvae
