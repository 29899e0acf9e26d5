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

// sin(x) for the bf16-mode epilogues: the MUFU approximation as is (one FMUL by 1/(2*pi) + MUFU.SIN; the unit
// works on the fractional number of revolutions, so it is periodic by construction).  Its error is
// ~4e-7 + |x| * 2^-24 (the rounding of x/(2*pi)), i.e. < 1e-5 for |x| < 100 -- far below the bf16 rounding
// applied right after.  An explicit range reduction in front cost five more FMA-pipe instructions per
// element, and the fused epilogues at C = 128 are bound by exactly that pipe (64 lanes/clk/SM), see
// profiles/r01_conv_umma2_swap_k7_c128_B4.ncu-rep.
__device__ __forceinline__ float sin_fast(float x) { return __sinf(x); }

template <bool kFastSin>
__device__ __forceinline__ float snake_beta(float v, float a, float inv_b) {
  // reference blocks.py:301-302: x + (1/(beta + 1e-9)) * sin(x*alpha)^2
  const float s = kFastSin ? sin_fast(v * a) : sinf(v * a);
  return fmaf(inv_b * s, s, v);
}

// fp32-mode SnakeBeta inside the tensor-core epilogues: sinf costs ~40 instructions per element in epilogues that are
// issue-bound, so reduce the argument to [-pi, pi] with a two-constant Cody-Waite step (exact to ~1e-7 for |x| < 1e4)
// and use the MUFU sine there, where its absolute error is ~4e-7 -- an order of magnitude inside the 1e-5 budget.
__device__ __forceinline__ float snake_beta_rr(float v, float a, float inv_b) {
  const float t = v * a;
  const float k = rintf(t * 0.15915494309189535f);
  float r = fmaf(k, -6.2831854820251465f, t);          // 2*pi rounded to fp32
  r = fmaf(k, 1.7484555314695172e-7f, r);              // 2*pi - fp32(2*pi) = -1.7484555e-7, sign folded in
  const float s = __sinf(r);
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
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int T_out = p.Tq_out * p.P_out;
    ptx::mbar_wait(acc_full, 0);
    ptx::tc_fence_after();
    for (int m = 0; m < p.MT; ++m) {
      const int q = q0 + m * 128 + row;
      const bool valid = q < p.Tq_out;
      const size_t orow = (static_cast<size_t>(b) * T_out + static_cast<size_t>(q) * p.P_out + phi) *
                          p.Cout;
      for (int c0 = 0; c0 < p.NT; c0 += 32) {
        uint32_t r[32];
        __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the predicated body
        ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + m * p.NT + c0, r);
        ptx::tmem_ld_wait();
        if (valid) {
        const int cbase = n0 + c0;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + cbase + j));
            v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
          }
        }
        if (p.residual) {
          if (p.residual_f32) {
            const float4* rp = reinterpret_cast<const float4*>(
                static_cast<const float*>(p.residual) + orow + cbase);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 x = __ldg(rp + j);
              v[4 * j] += x.x; v[4 * j + 1] += x.y; v[4 * j + 2] += x.z; v[4 * j + 3] += x.w;
            }
          } else {
            const uint4* rp = reinterpret_cast<const uint4*>(
                static_cast<const __nv_bfloat16*>(p.residual) + orow + cbase);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 x = __ldg(rp + j);
              const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[8 * j + 2 * e] += __uint_as_float(w[e] << 16);
                v[8 * j + 2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
              }
            }
          }
        }
        if (p.out_raw && p.out_raw_cf) {
          // channels-first store: lanes hold consecutive time steps -> 128 B (fp32) per channel
          const size_t t_out = static_cast<size_t>(q) * p.P_out + phi;
          const size_t o0 = (static_cast<size_t>(b) * p.Cout + cbase) * T_out + t_out;
          if (p.out_raw_f32) {
            float* op = static_cast<float*>(p.out_raw) + o0;
#pragma unroll
            for (int j = 0; j < 32; ++j) op[static_cast<size_t>(j) * T_out] = v[j];
          } else {
            __nv_bfloat16* op = static_cast<__nv_bfloat16*>(p.out_raw) + o0;
#pragma unroll
            for (int j = 0; j < 32; ++j) op[static_cast<size_t>(j) * T_out] = __float2bfloat16(v[j]);
          }
        } else if (p.out_raw) {
          if (p.out_raw_f32) {
            float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out_raw) + orow + cbase);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            uint4* op =
                reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out_raw) + orow + cbase);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
                w[e] = *reinterpret_cast<uint32_t*>(&h);
              }
              op[j] = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
        if (p.out_act) {
          if (p.snake_a) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(p.snake_a + cbase + j));
              const float4 ib = __ldg(reinterpret_cast<const float4*>(p.snake_inv_b + cbase + j));
              v[j] = snake_beta<true>(v[j], a.x, ib.x);
              v[j + 1] = snake_beta<true>(v[j + 1], a.y, ib.y);
              v[j + 2] = snake_beta<true>(v[j + 2], a.z, ib.z);
              v[j + 3] = snake_beta<true>(v[j + 3], a.w, ib.w);
            }
          }
          uint4* op = reinterpret_cast<uint4*>(p.out_act + orow + cbase);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
              w[e] = *reinterpret_cast<uint32_t*>(&h);
            }
            op[j] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        }  // valid
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace kvae
