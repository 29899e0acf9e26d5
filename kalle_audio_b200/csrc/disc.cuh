// Oobleck discriminator (stable_audio_tools/models/discriminators.py:62-297) -- the kernels around its convolutions.
//
// Every sub-discriminator is a stack of strided Conv1d(k = 15, s = 4, p = 7) + SiLU.  That is literally so for the
// multi-scale nets; the multi-period nets are Conv2d(15 x 15, stride 4, padding 7) over the waveform folded to
// [N, C, H = ceil(T / n), W = n] with n <= 11 < 15, so most of the 225 taps only ever meet zero padding.  Here the
// W axis moves into the channels: x' [N, C * W, H], w' [(co, wo), (ci, wi), kh] = w[co, ci, kh, wi - 4 wo + 7] (0 where
// that column index falls outside the kernel) -- the SAME arithmetic as the Conv2d minus the products with padding
// zeros (15 x fewer multiply-adds once W has shrunk to 1, which is the case from the second or third layer on).  This
// file holds the fold / unfold of inputs and weights, SiLU, avg_pool1d(2), the score mean, the hinge losses and the
// feature-matching distance over all feature tensors in one launch, each with its backward (first half), and the fp32
// kernels of the convolutions themselves: k = 15 / stride 4 forward, data gradient, weight gradient, and the final 1 x 1
// conv (second half).  Convs of any other geometry run on the generic layer kernels (conv_direct_kernel / wgrad_direct_kernel).
#pragma once
#include <cstdint>

namespace kvae {

constexpr int kDiscThreads = 256;

__device__ __forceinline__ float block_sum_256(float v, float* red /*[8]*/) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kDiscThreads / 32; ++i) s += red[i];
  return s;
}

// ---- MultiPeriodDiscriminator.fold (:164-168) with the period axis moved into the channels
// forward:  y[b, c * n + w, h] = x[b, c, h * n + w]  (0 for h * n + w >= T);  y: [N, C * n, H]
__global__ void __launch_bounds__(kDiscThreads)
disc_period_fold_kernel(const float* __restrict__ x, float* __restrict__ y, int C, long long T, int n, long long H,
                        size_t total) {
  for (size_t idx = blockIdx.x * static_cast<size_t>(kDiscThreads) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * kDiscThreads) {
    const long long h = static_cast<long long>(idx % H);
    const size_t cw = idx / H;                 // b * C * n + c * n + w
    const int w = static_cast<int>(cw % n);
    const size_t bc = cw / n;                  // b * C + c
    const long long t = h * n + w;
    y[idx] = t < T ? x[bc * T + t] : 0.f;
  }
}
// backward: gx[b, c, t] = gy[b, c * n + t % n, t / n];  gx: [N, C, T]
__global__ void __launch_bounds__(kDiscThreads)
disc_period_unfold_kernel(const float* __restrict__ gy, float* __restrict__ gx, long long T, int n, long long H,
                          size_t total) {
  for (size_t idx = blockIdx.x * static_cast<size_t>(kDiscThreads) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * kDiscThreads) {
    const long long t = static_cast<long long>(idx % T);
    const size_t bc = idx / T;
    gx[idx] = gy[(bc * n + static_cast<size_t>(t % n)) * H + t / n];
  }
}

// ---- nn.functional.avg_pool1d(x, 2) between the scales (:137): y[r, t] = (x[r, 2t] + x[r, 2t + 1]) / 2, To = T / 2
__global__ void __launch_bounds__(kDiscThreads)
disc_avg_pool2_kernel(const float* __restrict__ x, float* __restrict__ y, long long T, long long To, size_t total) {
  for (size_t idx = blockIdx.x * static_cast<size_t>(kDiscThreads) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * kDiscThreads) {
    const long long t = static_cast<long long>(idx % To);
    const size_t r = idx / To;
    const float2 v = *reinterpret_cast<const float2*>(x + r * T + 2 * t);   // r * T + 2t is even when T is even ...
    y[idx] = (v.x + v.y) * 0.5f;
  }
}
__global__ void __launch_bounds__(kDiscThreads)
disc_avg_pool2_odd_kernel(const float* __restrict__ x, float* __restrict__ y, long long T, long long To, size_t total) {
  for (size_t idx = blockIdx.x * static_cast<size_t>(kDiscThreads) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * kDiscThreads) {
    const long long t = static_cast<long long>(idx % To);
    const size_t r = idx / To;
    y[idx] = (x[r * T + 2 * t] + x[r * T + 2 * t + 1]) * 0.5f;             // ... odd T: rows are not 8-byte aligned
  }
}
// backward: gx[r, t] = gy[r, t / 2] / 2 for t < 2 To, 0 for the dropped last sample of an odd-length row
__global__ void __launch_bounds__(kDiscThreads)
disc_avg_pool2_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, long long T, long long To, size_t total) {
  for (size_t idx = blockIdx.x * static_cast<size_t>(kDiscThreads) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * kDiscThreads) {
    const long long t = static_cast<long long>(idx % T);
    const size_t r = idx / T;
    gx[idx] = t < 2 * To ? 0.5f * gy[r * To + t / 2] : 0.f;
  }
}

// ---- Conv2d weight [Cout, Cin, K, K] (torch layout, already weight-norm folded) -> Conv1d weight over the folded
// channels: wf[(co, wo), (ci, wi), kh] = w[co, ci, kh, wi - stride * wo + pad]; bias_f[(co, wo)] = bias[co]
__global__ void __launch_bounds__(kDiscThreads)
disc_fold_weight2d_kernel(const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ wf,
                          float* __restrict__ bias_f, int Cout, int Cin, int K, int stride, int pad, int W, int Wo) {
  const size_t total = static_cast<size_t>(Cout) * Wo * Cin * W * K;
  for (size_t idx = blockIdx.x * static_cast<size_t>(kDiscThreads) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * kDiscThreads) {
    const int kh = static_cast<int>(idx % K);
    size_t r = idx / K;
    const int wi = static_cast<int>(r % W); r /= W;
    const int ci = static_cast<int>(r % Cin); r /= Cin;
    const int wo = static_cast<int>(r % Wo);
    const int co = static_cast<int>(r / Wo);
    const int kw = wi - stride * wo + pad;
    wf[idx] = (kw >= 0 && kw < K) ? w[((static_cast<size_t>(co) * Cin + ci) * K + kh) * K + kw] : 0.f;
  }
  if (bias && bias_f)
    for (int i = blockIdx.x * kDiscThreads + threadIdx.x; i < Cout * Wo; i += gridDim.x * kDiscThreads)
      bias_f[i] = bias[i / Wo];
}
// adjoint: dw[co, ci, kh, kw] = sum_wo dwf[(co, wo), (ci, kw + stride * wo - pad), kh]; dbias[co] = sum_wo dbias_f[(co, wo)]
__global__ void __launch_bounds__(kDiscThreads)
disc_unfold_weight2d_kernel(const float* __restrict__ dwf, const float* __restrict__ dbias_f, float* __restrict__ dw,
                            float* __restrict__ dbias, int Cout, int Cin, int K, int stride, int pad, int W, int Wo) {
  const size_t total = static_cast<size_t>(Cout) * Cin * K * K;
  for (size_t idx = blockIdx.x * static_cast<size_t>(kDiscThreads) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * kDiscThreads) {
    const int kw = static_cast<int>(idx % K);
    size_t r = idx / K;
    const int kh = static_cast<int>(r % K); r /= K;
    const int ci = static_cast<int>(r % Cin);
    const int co = static_cast<int>(r / Cin);
    float s = 0.f;
    for (int wo = 0; wo < Wo; ++wo) {
      const int wi = kw + stride * wo - pad;
      if (wi >= 0 && wi < W) s += dwf[(((static_cast<size_t>(co) * Wo + wo) * Cin + ci) * W + wi) * K + kh];
    }
    dw[idx] = s;
  }
  if (dbias && dbias_f)
    for (int co = blockIdx.x * kDiscThreads + threadIdx.x; co < Cout; co += gridDim.x * kDiscThreads) {
      float s = 0.f;
      for (int wo = 0; wo < Wo; ++wo) s += dbias_f[co * Wo + wo];
      dbias[co] = s;
    }
}

// ---- nn.SiLU between the convs (:76): a = f * sigmoid(f)
__device__ __forceinline__ float silu_f(float f) { return f / (1.f + expf(-f)); }
__global__ void __launch_bounds__(kDiscThreads)
disc_silu_kernel(const float* __restrict__ f, float* __restrict__ a, size_t n) {
  const size_t n4 = n / 4;
  const size_t stride = static_cast<size_t>(gridDim.x) * kDiscThreads;
  const size_t i0 = blockIdx.x * static_cast<size_t>(kDiscThreads) + threadIdx.x;
  for (size_t i = i0; i < n4; i += stride) {
    float4 v = reinterpret_cast<const float4*>(f)[i];
    v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w);
    reinterpret_cast<float4*>(a)[i] = v;
  }
  for (size_t i = n4 * 4 + i0; i < n; i += stride) a[i] = silu_f(f[i]);
}
// gf = ga * d silu(f) / df + gfeat, d silu / df = s (1 + f (1 - s)), s = sigmoid(f).  gfeat (the gradient arriving at the
// feature tensor itself, from the feature-matching distance) may be null; gf may alias ga.
__device__ __forceinline__ float silu_grad_f(float f) {
  const float s = 1.f / (1.f + expf(-f));
  return s * (1.f + f * (1.f - s));
}
__global__ void __launch_bounds__(kDiscThreads)
disc_silu_bwd_kernel(const float* __restrict__ f, const float* ga, const float* __restrict__ gfeat, float* gf, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(kDiscThreads) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * kDiscThreads) {
    float g = ga[i] * silu_grad_f(f[i]);
    if (gfeat) g += gfeat[i];
    gf[i] = g;
  }
}

// ---- score = x.reshape(N, -1).mean(-1) of the last conv's output (:115); one block per batch item
__global__ void __launch_bounds__(kDiscThreads)
disc_score_kernel(const float* __restrict__ y, float* __restrict__ score, long long inner, int accumulate) {
  __shared__ float red[8];
  const float* row = y + static_cast<size_t>(blockIdx.x) * inner;
  float s = 0.f;
  for (long long i = threadIdx.x; i < inner; i += kDiscThreads) s += row[i];
  s = block_sum_256(s, red);
  if (threadIdx.x == 0) {
    const float m = s / static_cast<float>(inner);
    score[blockIdx.x] = accumulate ? score[blockIdx.x] + m : m;
  }
}
// gy[b, i] = gscore[b] / inner + gfeat[b, i]   (either source may be null)
__global__ void __launch_bounds__(kDiscThreads)
disc_score_bwd_kernel(const float* __restrict__ gscore, const float* __restrict__ gfeat, float* __restrict__ gy,
                      long long inner, size_t total) {
  for (size_t i = blockIdx.x * static_cast<size_t>(kDiscThreads) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * kDiscThreads) {
    float g = gscore ? gscore[i / inner] / static_cast<float>(inner) : 0.f;
    if (gfeat) g += gfeat[i];
    gy[i] = g;
  }
}

// ---- get_hinge_losses (:11-14) on score [2B] = (reals | fakes): losses[0] = dis, losses[1] = gen.
// With g_losses (d L / d dis, d L / d gen) it writes g_score [2B] instead.  One block.
__global__ void __launch_bounds__(kDiscThreads)
disc_hinge_kernel(const float* __restrict__ score, int B, float* __restrict__ losses, const float* __restrict__ g_losses,
                  float* __restrict__ g_score) {
  __shared__ float red[8];
  if (g_losses) {
    const float gd = g_losses[0] / static_cast<float>(B), gg = g_losses[1] / static_cast<float>(B);
    for (int i = threadIdx.x; i < 2 * B; i += kDiscThreads) {
      const float s = score[i];
      g_score[i] = i < B ? ((1.f - s) > 0.f ? -gd : 0.f) : (((1.f + s) > 0.f ? gd : 0.f) - gg);
    }
    return;
  }
  float r = 0.f, f = 0.f, m = 0.f;
  for (int i = threadIdx.x; i < B; i += kDiscThreads) {
    r += fmaxf(1.f - score[i], 0.f);
    f += fmaxf(1.f + score[B + i], 0.f);
    m += score[B + i];
  }
  r = block_sum_256(r, red);
  f = block_sum_256(f, red);
  m = block_sum_256(m, red);
  if (threadIdx.x == 0) {
    losses[0] = r / static_cast<float>(B) + f / static_cast<float>(B);
    losses[1] = -(m / static_cast<float>(B));
  }
}

// ---- feature-matching distance (:285-295): sum over the feature tensors of mean |real - fake|, the real half being the
// first `half` floats of a tensor and the fake half the next `half` (batch-concatenated input).  All tensors in ONE
// launch: the table rides in the kernel parameters; block -> tensor by its first block index.
constexpr int kFmMax = 56;
constexpr int kFmChunk = kDiscThreads * 16;
struct FmTable {
  const float* feat[kFmMax];
  float* grad[kFmMax];          // backward only: [2 * half] each
  long long half[kFmMax];
  int blk_begin[kFmMax + 1];
  int n;
};

__device__ __forceinline__ int fm_find(const FmTable& t, int blk) {
  int k = 0;
  while (k + 1 < t.n && t.blk_begin[k + 1] <= blk) ++k;
  return k;
}

// partial[block] = sum |r - f| / half over this block's chunk
__global__ void __launch_bounds__(kDiscThreads)
disc_fm_partial_kernel(const __grid_constant__ FmTable t, float* __restrict__ partial) {
  __shared__ float red[8];
  const int k = fm_find(t, blockIdx.x);
  const long long half = t.half[k];
  const long long i0 = static_cast<long long>(blockIdx.x - t.blk_begin[k]) * kFmChunk;
  const long long i1 = min(i0 + kFmChunk, half);
  const float* r = t.feat[k];
  const float* f = r + half;
  float s = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += kDiscThreads) s += fabsf(r[i] - f[i]);
  s = block_sum_256(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s / static_cast<float>(half);
}
// loss (+)= sum of the partials in a fixed order (one block)
__global__ void __launch_bounds__(kDiscThreads)
disc_fm_finish_kernel(const float* __restrict__ partial, int n, float* __restrict__ loss, int accumulate) {
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += kDiscThreads) s += partial[i];
  s = block_sum_256(s, red);
  if (threadIdx.x == 0) loss[0] = accumulate ? loss[0] + s : s;
}
// grad[k][i] = g * sign(r - f) / half, grad[k][half + i] = -that   (g = d L / d distance, one device float)
__global__ void __launch_bounds__(kDiscThreads)
disc_fm_bwd_kernel(const __grid_constant__ FmTable t, const float* __restrict__ g_loss) {
  const int k = fm_find(t, blockIdx.x);
  const long long half = t.half[k];
  const long long i0 = static_cast<long long>(blockIdx.x - t.blk_begin[k]) * kFmChunk;
  const long long i1 = min(i0 + kFmChunk, half);
  const float* r = t.feat[k];
  const float* f = r + half;
  float* gr = t.grad[k];
  const float g = g_loss[0] / static_cast<float>(half);
  for (long long i = i0 + threadIdx.x; i < i1; i += kDiscThreads) {
    const float d = r[i] - f[i];
    const float v = d > 0.f ? g : (d < 0.f ? -g : 0.f);
    gr[i] = v;
    gr[half + i] = -v;
  }
}

// =====================================================================================================================
// The discriminator's strided convolution -- k = 15, stride 4, padding 7 on [N, C, T] fp32 (discriminators.py:70-71, 85-
// 100: every conv of every net but the last 1 x 1) -- as three register-tiled fp32 kernels.  The generic layer kernels
// (conv_direct_kernel / wgrad_direct_kernel) reach ~9 / ~4 TFLOP/s on this geometry in the API layout (one tap and 16
// channels per barrier pair, 16-row stages in the weight gradient); here each thread keeps a register window of the
// input that serves all 15 taps of several outputs, shared memory is read with 16-byte loads at >= 13 FMAs per load, the
// layouts are chosen so that those loads are bank-conflict free (see each kernel), and the operand tiles of the next
// channel chunk arrive through cp.async while the current one is multiplied (two stages; the first version filled one
// stage between two barriers and was latency-bound: 320 us for a 256-channel layer whatever the grid size,
// profiles/r02_disc_launches_conv15_v1_ncu.csv).
// =====================================================================================================================
constexpr int kDK = 15, kDS = 4, kDP = 7;

__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
  const int sz = valid ? 4 : 0;      // 0 source bytes: the destination is zero-filled, nothing is read
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst, const float* src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- forward: y[n, co, t] = bias[co] + sum_{ci, k} wp[k][ci][co] * x[n, ci, 4 t + k - 7]      (wp = pack_weights layout)
// Block: 64 co x 128 t; 256 threads = 16 t-groups (8 consecutive t) x 16 co-groups (4 co); lane % 16 = t-group.
// A thread's 8 outputs x 15 taps read the 43 consecutive inputs 32 tg + 4 j + k: one register window per input channel.
// The slab row is stored in 32-float segments 36 floats apart, so the 8 lanes of a quarter warp (8 t-groups) hit 32
// different banks with their 16-byte loads; lanes of different co-groups read the same slab address (broadcast).
constexpr int kCfCo = 64, kCfKc = 8;
template <int TT>
struct CfGeom {
  static constexpr int BT = 16 * TT;                                   // outputs t per block
  static constexpr int SLAB = 4 * BT + 12;                             // slab entries per channel
  static constexpr int ROW = ((SLAB - 1) + 4 * ((SLAB - 1) >> 5) + 4) & ~3;   // 4 floats of padding after every 32
  static constexpr int XS = kCfKc * ROW, WS = kCfKc * kDK * kCfCo;
  static constexpr int STAGE = XS + WS;                                // floats per pipeline stage
  static constexpr size_t SMEM = 2 * static_cast<size_t>(STAGE) * sizeof(float);
};

// TT = outputs per thread: 8 (128 per block) for long layers; 2 (32 per block) when the long tiles would leave SMs idle
template <int TT>
__global__ void __launch_bounds__(256, 2)
disc_conv15_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wp, const float* __restrict__ bias,
                       float* __restrict__ y, int Cin, int Cout, int T, int To) {
  using G = CfGeom<TT>;
  extern __shared__ __align__(16) float disc_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tg = lane & 15, cg = warp * 2 + (lane >> 4);
  const int tb0 = blockIdx.x * G::BT, co0 = blockIdx.y * kCfCo, n = blockIdx.z;
  const long long u0 = static_cast<long long>(kDS) * tb0 - kDP;      // input index of slab entry 0
  const float* xn = x + static_cast<size_t>(n) * Cin * T;
  const bool w16 = (Cout & 3) == 0;
  float acc[4][TT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TT; ++j) acc[i][j] = 0.f;

  // one warp per input channel of the chunk (8 warps, 8 channels): no index arithmetic beyond an add per copy -- the first
  // version derived (channel, position) from a flat index with divisions, and a sixth of all issued instructions were
  // that integer work (ncu: ALU pipe 23 % next to FMA 49 %, profiles/r02_disc_conv15_fwd_msd0_l2_l4.ncu-rep)
  static_assert(kCfKc == 8, "one warp per channel");
  auto load = [&](int c0, int stage) {
    float* xs = disc_smem + stage * G::STAGE;
    float* ws = xs + G::XS;
    if (c0 + warp >= Cin) return;
    const float* xr = xn + static_cast<size_t>(c0 + warp) * T;
    float* dst = xs + warp * G::ROW;
    for (int j = lane; j < G::SLAB; j += 32) {
      const long long u = u0 + j;
      const bool ok = u >= 0 && u < T;
      cp_async4(dst + j + 4 * (j >> 5), ok ? xr + u : xn, ok);
    }
    float* wd = ws + warp * kDK * kCfCo;
    if (w16) {
      const int co = (lane & 15) * 4;
      const bool ok = co0 + co < Cout;
      for (int k = lane >> 4; k < kDK; k += 2)
        cp_async16(wd + k * kCfCo + co, ok ? wp + (static_cast<size_t>(k) * Cin + c0 + warp) * Cout + co0 + co : wp, ok);
    } else {
      for (int k = 0; k < kDK; ++k)
        for (int co = lane; co < kCfCo; co += 32) {
          const bool ok = co0 + co < Cout;
          cp_async4(wd + k * kCfCo + co, ok ? wp + (static_cast<size_t>(k) * Cin + c0 + warp) * Cout + co0 + co : wp, ok);
        }
    }
  };

  const int n_chunks = (Cin + kCfKc - 1) / kCfKc;
  load(0, 0);
  cp_async_commit();
  for (int c = 0; c < n_chunks; ++c) {
    if (c + 1 < n_chunks) load((c + 1) * kCfKc, (c + 1) & 1);
    cp_async_commit();
    cp_async_wait<1>();          // everything but the newest group has landed: chunk c
    __syncthreads();
    const float* xs = disc_smem + (c & 1) * G::STAGE;
    const float* ws = xs + G::XS;
    const int nci = min(kCfKc, Cin - c * kCfKc);
    for (int ci = 0; ci < nci; ++ci) {
      float xw[4 * TT + 12];
#pragma unroll
      for (int q = 0; q < TT + 3; ++q) {
        const int jj = 4 * TT * tg + 4 * q;
        const float4 v = *reinterpret_cast<const float4*>(xs + ci * G::ROW + jj + 4 * (jj >> 5));
        xw[4 * q] = v.x; xw[4 * q + 1] = v.y; xw[4 * q + 2] = v.z; xw[4 * q + 3] = v.w;
      }
#pragma unroll
      for (int k = 0; k < kDK; ++k) {
        const float4 w4 = *reinterpret_cast<const float4*>(ws + (ci * kDK + k) * kCfCo + 4 * cg);
#pragma unroll
        for (int j = 0; j < TT; ++j) {
          const float xv = xw[4 * j + k];
          acc[0][j] = fmaf(w4.x, xv, acc[0][j]);
          acc[1][j] = fmaf(w4.y, xv, acc[1][j]);
          acc[2][j] = fmaf(w4.z, xv, acc[2][j]);
          acc[3][j] = fmaf(w4.w, xv, acc[3][j]);
        }
      }
    }
    __syncthreads();             // this stage is overwritten by the load issued in the next iteration
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + 4 * cg + i;
    if (co >= Cout) continue;
    const float b = bias ? bias[co] : 0.f;
    float* yr = y + (static_cast<size_t>(n) * Cout + co) * To;
#pragma unroll
    for (int j = 0; j < TT; ++j) {
      const int t = tb0 + TT * tg + j;
      if (t < To) yr[t] = acc[i][j] + b;
    }
  }
}

// ---- data gradient: with v = u + 7, phi = v % 4, q = v / 4:  gx[n, ci, u] = sum_{co, m} wT[phi + 4 m][co][ci] * gy[n, co, q - m]
// (wT = pack_weights layout of the weight read as a ConvTranspose1d weight).  Block: 64 ci x 256 v; 256 threads = 16
// v-groups (16 consecutive v = 4 q x 4 phases) x 16 ci-groups (4 ci).  Per output channel a thread loads the 8 gradients
// gy[q0 - 4 .. q0 + 3] (two 16-byte loads, consecutive lanes consecutive addresses) and, per tap, 4 weights.
constexpr int kDgKc = 8;
template <int VQ, int CPT>
struct DgGeom {
  static constexpr int V = 4 * VQ, BV = 16 * V, CI = 16 * CPT;
  static constexpr int GUSED = 16 * VQ + 4, GROW = 16 * VQ + 8;
  static constexpr int GS = kDgKc * GROW, WS = kDgKc * kDK * CI;
  static constexpr int STAGE = GS + WS;
  static constexpr size_t SMEM = 2 * static_cast<size_t>(STAGE) * sizeof(float);
};

// VQ = q positions (x 4 phases) per thread: 4 (256 outputs per block) or 1 (64 per block, when the long tiles would leave
// SMs idle); CPT = input channels per thread: 4 (64-channel tile) or 2 (32-channel tile, for the 32- and 96-channel layers)
template <int VQ, int CPT>
__global__ void __launch_bounds__(256, 2)
disc_conv15_dgrad_kernel(const float* __restrict__ gy, const float* __restrict__ wT, float* __restrict__ gx, int Cin,
                         int Cout, int T, int To) {
  using G = DgGeom<VQ, CPT>;
  extern __shared__ __align__(16) float disc_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int vg = lane & 15, cig = warp * 2 + (lane >> 4);
  const int vb0 = blockIdx.x * G::BV, ci0 = blockIdx.y * G::CI, n = blockIdx.z;
  const int qb0 = vb0 / kDS;
  const float* gn = gy + static_cast<size_t>(n) * Cout * To;
  const bool w16 = (Cin & 3) == 0;
  float acc[CPT][G::V];
#pragma unroll
  for (int i = 0; i < CPT; ++i)
#pragma unroll
    for (int j = 0; j < G::V; ++j) acc[i][j] = 0.f;

  static_assert(kDgKc == 8, "one warp per channel");
  auto load = [&](int c0, int stage) {      // one warp per output channel of the chunk
    float* gs = disc_smem + stage * G::STAGE;
    float* ws = gs + G::GS;
    if (c0 + warp >= Cout) return;
    const float* gr = gn + static_cast<size_t>(c0 + warp) * To;
    float* dst = gs + warp * G::GROW;
    for (int i = lane; i < G::GUSED; i += 32) {
      const int t = qb0 - 4 + i;
      const bool ok = t >= 0 && t < To;
      cp_async4(dst + i, ok ? gr + t : gn, ok);
    }
    float* wd = ws + warp * kDK * G::CI;
    if (w16) {
      constexpr int VPR = G::CI / 4;          // 16-byte pieces per (co, k) row
      const int ci = (lane % VPR) * 4;
      const bool ok = ci0 + ci < Cin;
      for (int k = lane / VPR; k < kDK; k += 32 / VPR)
        cp_async16(wd + k * G::CI + ci, ok ? wT + (static_cast<size_t>(k) * Cout + c0 + warp) * Cin + ci0 + ci : wT, ok);
    } else {
      for (int k = 0; k < kDK; ++k)
        for (int ci = lane; ci < G::CI; ci += 32) {
          const bool ok = ci0 + ci < Cin;
          cp_async4(wd + k * G::CI + ci, ok ? wT + (static_cast<size_t>(k) * Cout + c0 + warp) * Cin + ci0 + ci : wT, ok);
        }
    }
  };

  const int n_chunks = (Cout + kDgKc - 1) / kDgKc;
  load(0, 0);
  cp_async_commit();
  for (int c = 0; c < n_chunks; ++c) {
    if (c + 1 < n_chunks) load((c + 1) * kDgKc, (c + 1) & 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* gs = disc_smem + (c & 1) * G::STAGE;
    const float* ws = gs + G::GS;
    const int nco = min(kDgKc, Cout - c * kDgKc);
    for (int co = 0; co < nco; ++co) {
      float gw[VQ + 4];
      if constexpr (VQ == 4) {
        const float4 a = *reinterpret_cast<const float4*>(gs + co * G::GROW + 4 * vg);
        const float4 b = *reinterpret_cast<const float4*>(gs + co * G::GROW + 4 * vg + 4);
        gw[0] = a.x; gw[1] = a.y; gw[2] = a.z; gw[3] = a.w; gw[4] = b.x; gw[5] = b.y; gw[6] = b.z; gw[7] = b.w;
      } else {
#pragma unroll
        for (int e = 0; e < VQ + 4; ++e) gw[e] = gs[co * G::GROW + VQ * vg + e];
      }
#pragma unroll
      for (int k = 0; k < kDK; ++k) {
        float w[CPT];
        if constexpr (CPT == 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(ws + (co * kDK + k) * G::CI + 4 * cig);
          w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
        } else {
          const float2 w2 = *reinterpret_cast<const float2*>(ws + (co * kDK + k) * G::CI + 2 * cig);
          w[0] = w2.x; w[1] = w2.y;
        }
        const int phi = k & 3, m = k >> 2;
#pragma unroll
        for (int qq = 0; qq < VQ; ++qq) {
          const float g = gw[4 + qq - m];
#pragma unroll
          for (int i = 0; i < CPT; ++i) acc[i][4 * qq + phi] = fmaf(w[i], g, acc[i][4 * qq + phi]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    const int ci = ci0 + CPT * cig + i;
    if (ci >= Cin) continue;
    float* gr = gx + (static_cast<size_t>(n) * Cin + ci) * T;
#pragma unroll
    for (int j = 0; j < G::V; ++j) {
      const int u = vb0 + G::V * vg + j - kDP;
      if (u >= 0 && u < T) gr[u] = acc[i][j];
    }
  }
}

// ---- data gradient of the io-side layer (2 .. 24 folded input channels, <= 128 output channels): too few channels for a
// channel tile.  One thread per output sample and group of 4 input channels; the group's weights [co][k][4 ci] sit in
// shared memory; per output channel a thread reads the <= 4 gradients q - m that reach its sample.
constexpr int kDsMaxCo = 128;

__global__ void __launch_bounds__(kDiscThreads)
disc_conv15_dgrad_small_kernel(const float* __restrict__ gy, const float* __restrict__ w, float* __restrict__ gx, int Cin,
                               int Cout, int T, int To) {
  __shared__ __align__(16) float ws[kDsMaxCo * kDK * 4];
  const int ci0 = 4 * blockIdx.y, n = blockIdx.z;
  for (int idx = threadIdx.x; idx < Cout * kDK * 4; idx += kDiscThreads) {
    const int i = idx & 3, r = idx >> 2;
    const int k = r % kDK, co = r / kDK;
    ws[idx] = (ci0 + i < Cin) ? w[(static_cast<size_t>(co) * Cin + ci0 + i) * kDK + k] : 0.f;
  }
  __syncthreads();
  const int u = blockIdx.x * kDiscThreads + threadIdx.x;
  if (u >= T) return;
  const int v = u + kDP, phi = v & 3, q = v >> 2;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const float* gn = gy + static_cast<size_t>(n) * Cout * To;
  for (int co = 0; co < Cout; ++co) {
    const float* gr = gn + static_cast<size_t>(co) * To;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int k = phi + 4 * m, t = q - m;
      if (k < kDK && t >= 0 && t < To) {
        const float g = gr[t];
        const float4 w4 = *reinterpret_cast<const float4*>(&ws[(co * kDK + k) * 4]);
        a0 = fmaf(w4.x, g, a0); a1 = fmaf(w4.y, g, a1); a2 = fmaf(w4.z, g, a2); a3 = fmaf(w4.w, g, a3);
      }
    }
  }
  float* gr = gx + (static_cast<size_t>(n) * Cin + ci0) * T + u;
  if (ci0 < Cin) gr[0] = a0;
  if (ci0 + 1 < Cin) gr[static_cast<size_t>(T)] = a1;
  if (ci0 + 2 < Cin) gr[2 * static_cast<size_t>(T)] = a2;
  if (ci0 + 3 < Cin) gr[3 * static_cast<size_t>(T)] = a3;
}

// ---- weight gradient: dw[co][ci][k] += sum_{n, t} gy[n, co, t] * x[n, ci, 4 t + k - 7]      (torch layout, atomics)
// Block: 64 co x 16 ci x 15 taps over a slice of the (n, 64-sample t block) items; 256 threads = 16 ci (lane % 16) x 16
// co-groups (4 co): 60 accumulators per thread.  Per 4 outputs t a thread loads 4 x 4 gradients and a 28-float input
// window (11 16-byte loads for 240 FMAs).  The input slab rows are 276 floats apart (= 20 mod 32), so the 8 channels of
// a quarter warp hit 32 different banks; gradient rows are read by all 16 lanes of a co-group at once (broadcast).
constexpr int kWgTc = 64, kWgCo = 64, kWgCi = 16;
constexpr int kWgXRow = 276, kWgXUsed = kDS * kWgTc + 12;   // 268 slab entries per channel
constexpr int kWgGRow = 68;
constexpr int kWgGs = kWgCo * kWgGRow, kWgXs = kWgCi * kWgXRow;
constexpr int kWgStage = kWgGs + kWgXs;
constexpr size_t kWgSmem = 2 * static_cast<size_t>(kWgStage) * sizeof(float);

__global__ void __launch_bounds__(256, 2)
disc_conv15_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ dw, int Cin,
                         int Cout, int T, int To, int n_items) {
  extern __shared__ __align__(16) float disc_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cil = lane & 15, cg = warp * 2 + (lane >> 4);
  const int n_ci_tiles = (Cin + kWgCi - 1) / kWgCi;
  const int co0 = (blockIdx.x / n_ci_tiles) * kWgCo, ci0 = (blockIdx.x % n_ci_tiles) * kWgCi;
  const int tiles_t = (To + kWgTc - 1) / kWgTc;
  float acc[4][kDK];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < kDK; ++k) acc[i][k] = 0.f;

  auto load = [&](int item, int stage) {      // warps walk whole rows: an add per copy, no divisions
    float* gs = disc_smem + stage * kWgStage;
    float* xs = gs + kWgGs;
    const int n = item / tiles_t, t0 = (item - n * tiles_t) * kWgTc;
    for (int row = warp; row < kWgCo; row += 8) {
      const int co = co0 + row;
      const float* gr = gy + (static_cast<size_t>(n) * Cout + min(co, Cout - 1)) * To;
      for (int col = lane; col < kWgTc; col += 32) {
        const int t = t0 + col;
        const bool ok = co < Cout && t < To;
        cp_async4(gs + row * kWgGRow + col, ok ? gr + t : gy, ok);
      }
    }
    const long long u0 = static_cast<long long>(kDS) * t0 - kDP;
    for (int ci = warp; ci < kWgCi; ci += 8) {
      const int c = ci0 + ci;
      const float* xr = x + (static_cast<size_t>(n) * Cin + min(c, Cin - 1)) * T;
      for (int j = lane; j < kWgXUsed; j += 32) {
        const long long u = u0 + j;
        const bool ok = c < Cin && u >= 0 && u < T;
        cp_async4(xs + ci * kWgXRow + j, ok ? xr + u : x, ok);
      }
    }
  };

  int item = blockIdx.y, stage = 0;
  if (item < n_items) load(item, 0);
  cp_async_commit();
  for (; item < n_items; item += gridDim.y, stage ^= 1) {
    if (item + static_cast<int>(gridDim.y) < n_items) load(item + gridDim.y, stage ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* gs = disc_smem + stage * kWgStage;
    const float* xs = gs + kWgGs;
#pragma unroll 1
    for (int tau0 = 0; tau0 < kWgTc; tau0 += 4) {
      float g[4][4], xw[28];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(gs + (4 * cg + i) * kWgGRow + tau0);
        g[i][0] = v.x; g[i][1] = v.y; g[i][2] = v.z; g[i][3] = v.w;
      }
#pragma unroll
      for (int q = 0; q < 7; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(xs + cil * kWgXRow + kDS * tau0 + 4 * q);
        xw[4 * q] = v.x; xw[4 * q + 1] = v.y; xw[4 * q + 2] = v.z; xw[4 * q + 3] = v.w;
      }
#pragma unroll
      for (int tt = 0; tt < 4; ++tt)
#pragma unroll
        for (int k = 0; k < kDK; ++k) {
          const float xv = xw[4 * tt + k];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i][k] = fmaf(g[i][tt], xv, acc[i][k]);
        }
    }
    __syncthreads();
  }
  const int c = ci0 + cil;
  if (c < Cin) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int co = co0 + 4 * cg + i;
      if (co >= Cout) continue;
      float* d = dw + (static_cast<size_t>(co) * Cin + c) * kDK;
#pragma unroll
      for (int k = 0; k < kDK; ++k) atomicAdd(d + k, acc[i][k]);
    }
  }
}

// ---- the nets' last layer: a 1 x 1 convolution onto a few channels (256 -> 1 per folded column).  Pure streaming:
// forward  y[n, co, t] = b[co] + sum_ci w[co][ci] a[n, ci, t]  (one thread per (n, t), channels walked with coalesced loads),
// data gradient  ga[n, ci, t] = sum_co w[co][ci] gy[n, co, t],  weight gradient  dw[co][ci] = sum_{n, t} gy[n, co, t] a[n, ci, t].
constexpr int kC1MaxCo = 8;

// block: 32 consecutive t x 8 channel slices (ci = slice, slice + 8, ...), reduced through shared memory
__global__ void __launch_bounds__(kDiscThreads)
disc_conv1x1_fwd_kernel(const float* __restrict__ a, const float* __restrict__ w, const float* __restrict__ bias,
                        float* __restrict__ y, int Cin, int Cout, long long T) {
  __shared__ float red[8][kC1MaxCo][33];
  const int tl = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const long long t = blockIdx.x * 32ll + tl;
  const int n = blockIdx.y;
  float acc[kC1MaxCo];
#pragma unroll
  for (int co = 0; co < kC1MaxCo; ++co) acc[co] = 0.f;
  if (t < T) {
    const float* an = a + static_cast<size_t>(n) * Cin * T + t;
#pragma unroll 4
    for (int ci = sl; ci < Cin; ci += 8) {
      const float v = an[static_cast<size_t>(ci) * T];
#pragma unroll
      for (int co = 0; co < kC1MaxCo; ++co)
        if (co < Cout) acc[co] = fmaf(w[co * Cin + ci], v, acc[co]);
    }
  }
#pragma unroll
  for (int co = 0; co < kC1MaxCo; ++co) red[sl][co][tl] = acc[co];
  __syncthreads();
  for (int idx = threadIdx.x; idx < Cout * 32; idx += kDiscThreads) {
    const int co = idx >> 5, tt = idx & 31;
    const long long to = blockIdx.x * 32ll + tt;
    if (to >= T) continue;
    float v = bias ? bias[co] : 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) v += red[q][co][tt];
    y[(static_cast<size_t>(n) * Cout + co) * T + to] = v;
  }
}

// one thread per element of ga [N, Cin, T]
__global__ void __launch_bounds__(kDiscThreads)
disc_conv1x1_dgrad_kernel(const float* __restrict__ gy, const float* __restrict__ w, float* __restrict__ ga, int Cin,
                          int Cout, long long T, size_t total) {
  for (size_t idx = blockIdx.x * static_cast<size_t>(kDiscThreads) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * kDiscThreads) {
    const long long t = static_cast<long long>(idx % T);
    const size_t r = idx / T;
    const int ci = static_cast<int>(r % Cin);
    const size_t n = r / Cin;
    float v = 0.f;
    for (int co = 0; co < Cout; ++co) v = fmaf(w[co * Cin + ci], gy[(n * Cout + co) * T + t], v);
    ga[idx] = v;
  }
}

// dbias[c] += sum over this block's slice of the (n, t) positions of gy[n, c, t]; grid (C, splits), dbias zeroed before
__global__ void __launch_bounds__(kDiscThreads)
disc_bias_grad_kernel(const float* __restrict__ gy, float* __restrict__ db, int N, int C, long long T) {
  __shared__ float red[8];
  const int c = blockIdx.x;
  const long long total = static_cast<long long>(N) * T;
  const long long per = (total + gridDim.y - 1) / gridDim.y;
  const long long r0 = blockIdx.y * per, r1 = min(r0 + per, total);
  float s = 0.f;
  for (long long r = r0 + threadIdx.x; r < r1; r += kDiscThreads) {
    const long long n = r / T, t = r - n * T;
    s += gy[(static_cast<size_t>(n) * C + c) * T + t];
  }
  s = block_sum_256(s, red);
  if (threadIdx.x == 0) atomicAdd(db + c, s);
}

// one block per (ci, co): dw[co][ci] = sum over (n, t)
__global__ void __launch_bounds__(kDiscThreads)
disc_conv1x1_wgrad_kernel(const float* __restrict__ a, const float* __restrict__ gy, float* __restrict__ dw, int N, int Cin,
                          int Cout, long long T) {
  __shared__ float red[8];
  const int ci = blockIdx.x, co = blockIdx.y;
  float s = 0.f;
  for (int n = 0; n < N; ++n) {
    const float* ar = a + (static_cast<size_t>(n) * Cin + ci) * T;
    const float* gr = gy + (static_cast<size_t>(n) * Cout + co) * T;
    for (long long t = threadIdx.x; t < T; t += kDiscThreads) s = fmaf(ar[t], gr[t], s);
  }
  s = block_sum_256(s, red);
  if (threadIdx.x == 0) dw[co * Cin + ci] = s;
}

}  // namespace kvae
