// Implicit-GEMM 1-D convolution on the sm_100a tensor cores (tcgen05.mma, TMEM accumulators,
// TMA-fed operands).  One kernel serves the three weight-normalised convolution shapes of the
// Oobleck stack (reference: stable_audio_tools/models/autoencoders.py:39-114):
//   * Conv1d stride 1, dilation d       (ResidualUnit k7 / k1, first/last convs)   :49-53,133,141,168
//   * Conv1d stride s, kernel 2s        (EncoderBlock down-sampling)                :76-77
//   * ConvTranspose1d stride s          (DecoderBlock up-sampling, polyphase)       :98-100
//
// Data layout in HBM: activations are channels-last [B, T, C] bf16, so a tap shift is a whole
// 128-byte-row offset of a K-major operand tile; weights are pre-packed [tap][Cout][Cin] bf16.
//
// GEMM view per CTA: D[M = 128*MT time rows, N = NT out-channels] += A[M, 64 ch] * W[N, 64 ch]^T
// summed over (64-channel chunk) x (tap).  A "tap" is (input phase, row shift, weight slab); taps
// that read the same input phase share ONE staged A slab (rows = 128*MT + shift span) and address
// it through row-shifted shared-memory descriptors, so the activations are fetched once per
// chunk instead of once per tap.
//
// Warp roles (256 threads): warp 0 TMA producer, warp 1 UMMA issuer, warp 2 TMEM allocator,
// warps 4-7 epilogue (TMEM -> registers -> bias / residual / SnakeBeta -> global).
#pragma once
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>

#include "ptx.cuh"

namespace kvae {

constexpr int kMaxTaps = 32;
constexpr int kMaxPhases = 8;

struct Tap {
  int16_t a_phase;  // phase of the input view (strided conv), else 0
  int16_t a_row;    // slab start row relative to the tile's first output row (valid if first)
  int16_t shift;    // row offset of this tap inside the staged slab (>= 0)
  int16_t w_slab;   // packed-weight slab index
  int16_t first;    // 1: a new A slab is staged before this tap
  int16_t last;     // 1: last tap reading the current A slab
};

struct ConvParams {
  int B, Tq_out, P_out, Cout;
  int n_chunks;           // Cin / 64
  int MT, NT;             // 128-row sub-tiles per CTA; out-channels per CTA
  int RB, nbox;           // A slab = nbox TMA boxes of RB rows
  int SA, SB;             // ring depths
  int desc_mode;          // 0: descriptor base_offset = 0; 1: base_offset = (addr >> 7) & 7
  int tmem_cols;          // power of two >= MT*NT
  int tap_begin[kMaxPhases + 1];
  Tap taps[kMaxTaps];
  // epilogue
  const float* bias;        // [Cout] or nullptr
  const void* residual;     // [B, T_out, Cout] or nullptr
  int residual_f32;         // residual element type: 1 fp32, 0 bf16
  void* out_raw;            // pre-activation output or nullptr
  int out_raw_f32;
  int out_raw_cf;           // 1: out_raw is channels-first [B, Cout, T_out] (API layout), else [B, T_out, Cout]
  __nv_bfloat16* out_act;   // bf16 operand for the next conv (SnakeBeta applied if snake_a)
  const float* snake_a;     // exp(alpha)            [Cout] or nullptr (plain cast)
  const float* snake_inv_b; // 1/(exp(beta) + 1e-9)  [Cout]
};

__host__ __device__ inline size_t conv_umma_smem_bytes(const ConvParams& p) {
  return 1024 /*barriers*/ + 1024 /*alignment slack*/ +
         static_cast<size_t>(p.SA) * p.nbox * p.RB * 128 + static_cast<size_t>(p.SB) * p.NT * 128;
}

// sin with a two-constant Cody-Waite reduction to [-pi, pi] in front of the MUFU approximation:
// absolute error ~5e-7 for |x| up to ~1e4, at 4 extra FMA-pipe instructions (plain __sinf loses
// accuracy quickly outside [-pi, pi]).
__device__ __forceinline__ float sin_fast(float x) {
  const float k = rintf(x * 0.15915494309189535f);
  float r = fmaf(-k, 6.2831854820251465f, x);       // 2*pi rounded to fp32
  r = fmaf(-k, -1.7484555e-7f, r);                  // 2*pi - fp32(2*pi)
  return __sinf(r);
}

template <bool kFastSin>
__device__ __forceinline__ float snake_beta(float v, float a, float inv_b) {
  // reference blocks.py:301-302: x + (1/(beta + 1e-9)) * sin(x*alpha)^2
  const float s = kFastSin ? sin_fast(v * a) : sinf(v * a);
  return fmaf(inv_b * s, s, v);
}

__global__ void __launch_bounds__(256, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  // barrier block lives in the first 1 KB of the (1024-aligned) carve-out
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* smem = smem_raw + pad;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* a_empty = a_full + 8;
  uint64_t* b_full = a_full + 16;
  uint64_t* b_empty = a_full + 32;
  uint64_t* acc_full = a_full + 48;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 49);
  uint8_t* a_ring = smem + 1024;
  const uint32_t a_bytes = static_cast<uint32_t>(p.nbox) * p.RB * 128;
  const uint32_t b_bytes = static_cast<uint32_t>(p.NT) * 128;
  uint8_t* b_ring = a_ring + static_cast<size_t>(p.SA) * a_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates
  const int phi = blockIdx.x % p.P_out;
  const int q0 = (blockIdx.x / p.P_out) * (128 * p.MT);
  const int n0 = blockIdx.y * p.NT;
  const int b = blockIdx.z;
  const int t_lo = p.tap_begin[phi], t_hi = p.tap_begin[phi + 1];

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    for (int i = 0; i < p.SA; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < p.SB; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], 1); }
    ptx::mbar_init(acc_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        for (int t = t_lo; t < t_hi; ++t) {
          const Tap tap = p.taps[t];
          if (tap.first) {
            ptx::mbar_wait(&a_empty[as], aph ^ 1u);
            ptx::mbar_expect_tx(&a_full[as], a_bytes);
            for (int bx = 0; bx < p.nbox; ++bx)
              ptx::tma_load_4d(a_ring + static_cast<size_t>(as) * a_bytes + bx * p.RB * 128, &tmA,
                               &a_full[as], ch * 64, tap.a_phase, q0 + tap.a_row + bx * p.RB, b);
            if (++as == p.SA) { as = 0; aph ^= 1u; }
          }
          ptx::mbar_wait(&b_empty[bs], bph ^ 1u);
          ptx::mbar_expect_tx(&b_full[bs], b_bytes);
          ptx::tma_load_3d(b_ring + static_cast<size_t>(bs) * b_bytes, &tmW, &b_full[bs], ch * 64, n0,
                           tap.w_slab);
          if (++bs == p.SB) { bs = 0; bph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ UMMA issuer
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::idesc_bf16_f32(128, p.NT);
      const uint32_t a_base = ptx::smem_u32(a_ring);
      const uint32_t b_base = ptx::smem_u32(b_ring);
      int as = 0, bs = 0, cur = 0;
      uint32_t aph = 0, bph = 0;
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        for (int t = t_lo; t < t_hi; ++t) {
          const Tap tap = p.taps[t];
          if (tap.first) {
            ptx::mbar_wait(&a_full[as], aph);
            cur = as;
            if (++as == p.SA) { as = 0; aph ^= 1u; }
          }
          ptx::mbar_wait(&b_full[bs], bph);
          ptx::tc_fence_after();
          const uint32_t fresh = (ch == 0 && t == t_lo) ? 1u : 0u;
          for (int m = 0; m < p.MT; ++m) {
            const uint32_t a_tile = a_base + cur * a_bytes + (tap.shift + 128 * m) * 128;
            const uint32_t b_tile = b_base + bs * b_bytes;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t aa = a_tile + k * 32;
              const uint32_t bo = p.desc_mode ? ((aa >> 7) & 7u) : 0u;
              ptx::umma_f16(tmem_base + m * p.NT, ptx::smem_desc_sw128(aa, bo),
                            ptx::smem_desc_sw128(b_tile + k * 32, 0), idesc,
                            (fresh && k == 0) ? 0u : 1u);
            }
          }
          ptx::umma_commit(&b_empty[bs]);
          if (tap.last) ptx::umma_commit(&a_empty[cur]);
          if (++bs == p.SB) { bs = 0; bph ^= 1u; }
        }
      }
      ptx::umma_commit(acc_full);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    // TMEM lane = time row, so tcgen05.ld hands each thread one row x 32 channels.  Rows are Cout
    // elements apart in HBM, so that mapping would touch 32 different lines per warp instruction.
    // Each warp therefore transposes its 32x32 block through shared memory (the operand rings are
    // idle once acc_full fires) into "8 lanes x 16 B = one 128-byte row segment", so every residual
    // load and every output store is a fully used line.
    const int quad = warp & 3;
    const int T_out = p.Tq_out * p.P_out;
    float* stage = reinterpret_cast<float*>(a_ring) + quad * (32 * 33);
    const int tr = lane >> 3;         // row within a group of 4
    const int tc = (lane & 7) * 4;    // first of this lane's 4 channels inside the 32-channel block
    ptx::mbar_wait(acc_full, 0);
    ptx::tc_fence_after();
    for (int m = 0; m < p.MT; ++m) {
      const int qw = q0 + m * 128 + quad * 32;   // first row of this warp's block
      for (int c0 = 0; c0 < p.NT; c0 += 32) {
        uint32_t r[32];
        __syncwarp();  // tcgen05.ld is .sync.aligned; also orders the previous block's stage reads
        ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + m * p.NT + c0, r);
        ptx::tmem_ld_wait();
        const int cbase = n0 + c0;
        if (p.out_raw && p.out_raw_cf) {
          // channels-first API output (no residual / activation on this path): thread = time row,
          // lanes hold consecutive time steps -> 128 B per channel per warp, already coalesced
          const int q = qw + lane;
          if (q < p.Tq_out) {
            const size_t t_out = static_cast<size_t>(q) * p.P_out + phi;
            const size_t o0 = (static_cast<size_t>(b) * p.Cout + cbase) * T_out + t_out;
            if (p.out_raw_f32) {
              float* op = static_cast<float*>(p.out_raw) + o0;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float v = __uint_as_float(r[j]);
                if (p.bias) v += __ldg(p.bias + cbase + j);
                op[static_cast<size_t>(j) * T_out] = v;
              }
            } else {
              __nv_bfloat16* op = static_cast<__nv_bfloat16*>(p.out_raw) + o0;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float v = __uint_as_float(r[j]);
                if (p.bias) v += __ldg(p.bias + cbase + j);
                op[static_cast<size_t>(j) * T_out] = __float2bfloat16(v);
              }
            }
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = __uint_as_float(r[j]);
        __syncwarp();
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), sa4 = bias4, sib4 = bias4;
        if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + cbase + tc));
        if (p.snake_a) {
          sa4 = __ldg(reinterpret_cast<const float4*>(p.snake_a + cbase + tc));
          sib4 = __ldg(reinterpret_cast<const float4*>(p.snake_inv_b + cbase + tc));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + tr;
          const int q = qw + rr;
          if (q >= p.Tq_out) continue;
          const size_t o = (static_cast<size_t>(b) * T_out + static_cast<size_t>(q) * p.P_out + phi) * p.Cout +
                           cbase + tc;
          float v0 = stage[rr * 33 + tc] + bias4.x, v1 = stage[rr * 33 + tc + 1] + bias4.y;
          float v2 = stage[rr * 33 + tc + 2] + bias4.z, v3 = stage[rr * 33 + tc + 3] + bias4.w;
          if (p.residual) {
            if (p.residual_f32) {
              const float4 x = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(p.residual) + o));
              v0 += x.x; v1 += x.y; v2 += x.z; v3 += x.w;
            } else {
              const uint2 x = __ldg(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(p.residual) + o));
              v0 += __uint_as_float(x.x << 16); v1 += __uint_as_float(x.x & 0xffff0000u);
              v2 += __uint_as_float(x.y << 16); v3 += __uint_as_float(x.y & 0xffff0000u);
            }
          }
          if (p.out_raw) {
            if (p.out_raw_f32) {
              *reinterpret_cast<float4*>(static_cast<float*>(p.out_raw) + o) = make_float4(v0, v1, v2, v3);
            } else {
              __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1), h1 = __floats2bfloat162_rn(v2, v3);
              *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out_raw) + o) =
                  make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
            }
          }
          if (p.out_act) {
            if (p.snake_a) {
              v0 = snake_beta<true>(v0, sa4.x, sib4.x); v1 = snake_beta<true>(v1, sa4.y, sib4.y);
              v2 = snake_beta<true>(v2, sa4.z, sib4.z); v3 = snake_beta<true>(v3, sa4.w, sib4.w);
            }
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1), h1 = __floats2bfloat162_rn(v2, v3);
            *reinterpret_cast<uint2*>(p.out_act + o) =
                make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace kvae
