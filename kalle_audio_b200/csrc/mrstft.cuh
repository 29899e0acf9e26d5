// Multi-resolution STFT loss of the autoencoder training wrapper, value and gradients in one pass over the signals
// (reference: /root/reference/stable_audio_tools/training/losses/auraloss.py -- FIRFilter "aw" 70-162, STFTLoss 220-441,
// MultiResolutionSTFTLoss 443-531, SumAndDifferenceSTFTLoss 534-606; used at training/autoencoders.py:123-175 as
// module(input = reals, target = decoded)).  SURVEY section 8(f) item 4, loss half.
//
//   prep      [B, C, T] -> M mono signals (identity view, or sum / difference of a stereo pair), A-weighting FIR
//   forward   per resolution: a block transforms 2048 / n frames at once -- input and target ride through ONE complex
//             radix-2 FFT in shared memory (z = x + i y, spectra separated by the Hermitian split) -- and accumulates
//             sum (ym - xm)^2, sum ym^2 and sum |log xm - log ym| per signal (double atomics)
//   finish    loss = mean over resolutions of  w_sc * mean_m ||ym - xm|| / ||ym||  +  w_log * mean |log xm - log ym|
//   backward  per resolution the frames are transformed again (nothing the size of a spectrogram is ever stored), the
//             magnitude gradients become two Hermitian spectra, ONE inverse FFT returns both time-domain gradients, which
//             are windowed and scattered (atomicAdd) through the reflect padding; then the transposed FIR / sum-difference
// The reference runs 2 x 7 torch.stft calls, ~60 elementwise kernels over spectrogram-sized tensors and their autograd
// mirror; HBM traffic here is the signals themselves, once per resolution and direction.
#pragma once
#include <cuda_bf16.h>
#include <cstdint>

#include "elementwise.cuh"

namespace kvae {

constexpr int kStftMaxN = 2048;      // largest fft size; a block holds kStftMaxN complex points (16 KB)
constexpr int kStftMaxRes = 16;
constexpr int kFirMaxTaps = 129;
constexpr int kFirTile = 1024;

struct StftRes { int n, log2n, hop, frames; };

// ---------------------------------------------------------------- prep: view / sum-difference + FIR (zero padding)
// signal m: sum_diff ? (m < B ? x[b,0] + x[b,1] : x[b,0] - x[b,1], b = m % B) : x viewed as [B*C, T]
__device__ __forceinline__ float stft_src(const void* x, int f32, int sum_diff, int B, long long T, int m, long long t) {
  if (t < 0 || t >= T) return 0.f;
  if (!sum_diff) return ld_elem(x, static_cast<size_t>(m) * T + t, f32);
  const int b = m % B;
  const float l = ld_elem(x, (static_cast<size_t>(b) * 2) * T + t, f32), r = ld_elem(x, (static_cast<size_t>(b) * 2 + 1) * T + t, f32);
  return m < B ? l + r : l - r;
}
// grid (ceil(T / kFirTile), M); out[m][t] = sum_k src[t + k - pad] * taps[k]   (F.conv1d = cross-correlation)
__global__ void __launch_bounds__(256) mrstft_prep_kernel(const void* x, int f32, int sum_diff, int B, long long T,
                                                          const float* taps, int ntaps, float* out) {
  __shared__ float s[kFirTile + kFirMaxTaps];
  __shared__ float tp[kFirMaxTaps];
  const int m = blockIdx.y;
  const long long t0 = static_cast<long long>(blockIdx.x) * kFirTile;
  const int pad = ntaps / 2;
  if (ntaps == 0) {
    for (int i = threadIdx.x; i < kFirTile; i += 256)
      if (t0 + i < T) out[static_cast<size_t>(m) * T + t0 + i] = stft_src(x, f32, sum_diff, B, T, m, t0 + i);
    return;
  }
  for (int i = threadIdx.x; i < ntaps; i += 256) tp[i] = taps[i];
  for (int i = threadIdx.x; i < kFirTile + ntaps - 1; i += 256) s[i] = stft_src(x, f32, sum_diff, B, T, m, t0 + i - pad);
  __syncthreads();
  for (int i = threadIdx.x; i < kFirTile; i += 256) {
    if (t0 + i >= T) break;
    float acc = 0.f;
    for (int k = 0; k < ntaps; ++k) acc = fmaf(s[i + k], tp[k], acc);
    out[static_cast<size_t>(m) * T + t0 + i] = acc;
  }
}
// transpose of prep: g [M, T] -> grad [B, C, T] fp32:  FIR^T then the transposed view / sum-difference
// FIR^T: h[m][t] = sum_k g[m][t + pad - k] * taps[k]
__global__ void __launch_bounds__(256) mrstft_prep_T_kernel(const float* g, int sum_diff, int B, int C, long long T,
                                                            const float* taps, int ntaps, float scale, float* grad) {
  __shared__ float s[2][kFirTile + kFirMaxTaps];
  __shared__ float tp[kFirMaxTaps];
  const int pad = ntaps / 2;
  const long long t0 = static_cast<long long>(blockIdx.x) * kFirTile;
  const int nsig = sum_diff ? 2 : 1;                       // signals this block combines
  const int b = blockIdx.y;                                // sum_diff: batch item; else: signal m
  for (int i = threadIdx.x; i < ntaps; i += 256) tp[i] = taps[i];
  for (int q = 0; q < nsig; ++q) {
    const size_t m = sum_diff ? static_cast<size_t>(q) * B + b : b;
    for (int i = threadIdx.x; i < kFirTile + (ntaps ? ntaps - 1 : 0); i += 256) {
      const long long t = t0 + i - pad;                    // s[q][i] = g[t0 + i - pad]
      s[q][i] = (t >= 0 && t < T) ? g[m * T + t] : 0.f;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kFirTile; i += 256) {
    if (t0 + i >= T) break;
    float h[2] = {0.f, 0.f};
    for (int q = 0; q < nsig; ++q) {
      if (ntaps == 0) { h[q] = s[q][i]; continue; }
      float acc = 0.f;
      for (int k = 0; k < ntaps; ++k) acc = fmaf(s[q][i + 2 * pad - k], tp[k], acc);   // g[t + pad - k]
      h[q] = acc;
    }
    if (sum_diff) {
      grad[(static_cast<size_t>(b) * 2) * T + t0 + i] = scale * (h[0] + h[1]);
      grad[(static_cast<size_t>(b) * 2 + 1) * T + t0 + i] = scale * (h[0] - h[1]);
    } else {
      grad[static_cast<size_t>(b) * T + t0 + i] = scale * h[0];
    }
  }
}

// ---------------------------------------------------------------- shared-memory FFT over `nf` frames of n points
// radix-2 decimation in time: input in bit-reversed order, output in natural order.  tw[j] = exp(-2 pi i j / n);
// inverse = conjugated twiddles.  All threads of the block call it.
__device__ __forceinline__ void stft_fft(float2* z, const float2* tw, int n, int log2n, int nf, bool inverse) {
  const int half_n = n >> 1;
  for (int s = 1; s <= log2n; ++s) {
    const int half = 1 << (s - 1);
    const int tstride = half_n >> (s - 1);
    for (int idx = threadIdx.x; idx < nf * half_n; idx += blockDim.x) {
      const int f = idx / half_n, j = idx - f * half_n;    // half_n is a power of two: shifts
      const int k = j & (half - 1);
      const int i0 = ((j - k) << 1) + k, i1 = i0 + half;
      float2 w = tw[k * tstride];
      if (inverse) w.y = -w.y;
      float2* zf = z + f * n;
      const float2 a = zf[i0], b = zf[i1];
      const float2 bw = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
      zf[i0] = make_float2(a.x + bw.x, a.y + bw.y);
      zf[i1] = make_float2(a.x - bw.x, a.y - bw.y);
    }
    __syncthreads();
  }
}
__device__ __forceinline__ long long stft_reflect(long long p, long long T) {
  if (p < 0) p = -p;
  if (p >= T) p = 2 * (T - 1) - p;
  return p;
}
// loads frames [f0, f0 + nf) of signal m of X and Y as z = w * (x + i y), bit-reversed; builds the twiddle table
__device__ __forceinline__ void stft_load(float2* z, float2* tw, const float* X, const float* Y, const float* win,
                                          const StftRes r, long long T, int m, int f0, int nf) {
  const int n = r.n;
  for (int j = threadIdx.x; j < (n >> 1); j += blockDim.x) {
    float sn, cs;
    sincospif(-2.0f * static_cast<float>(j) / static_cast<float>(n), &sn, &cs);
    tw[j] = make_float2(cs, sn);
  }
  for (int idx = threadIdx.x; idx < nf * n; idx += blockDim.x) {
    const int f = idx / n, i = idx - f * n;
    const long long p = stft_reflect(static_cast<long long>(f0 + f) * r.hop + i - (n >> 1), T);
    const float w = __ldg(win + i);
    const int ir = static_cast<int>(__brev(static_cast<unsigned>(i)) >> (32 - r.log2n));
    z[f * n + ir] = make_float2(w * X[static_cast<size_t>(m) * T + p], w * Y[static_cast<size_t>(m) * T + p]);
  }
  __syncthreads();
}
// spectra of the two real signals from Z = FFT(x + i y):  X[f] = (Z[f] + conj Z[n-f]) / 2,  Y[f] = (Z[f] - conj Z[n-f]) / 2i
__device__ __forceinline__ void stft_split(const float2 zf, const float2 zc, float2& xs, float2& ys) {
  xs = make_float2(0.5f * (zf.x + zc.x), 0.5f * (zf.y - zc.y));
  ys = make_float2(0.5f * (zf.y + zc.y), 0.5f * (zc.x - zf.x));
}

constexpr float kStftEps = 1e-8f;

// grid (ceil(frames / fpb), M), block 256; sums [M][3] doubles of this resolution: S1, S2, L1
__global__ void __launch_bounds__(256) mrstft_fwd_kernel(const float* X, const float* Y, const float* win, StftRes r,
                                                         long long T, double* sums) {
  __shared__ float2 z[kStftMaxN];
  __shared__ float2 tw[kStftMaxN / 2];
  __shared__ double red[3][8];
  const int n = r.n, fpb = kStftMaxN / n;
  const int m = blockIdx.y, f0 = blockIdx.x * fpb;
  const int nf = min(fpb, r.frames - f0);
  stft_load(z, tw, X, Y, win, r, T, m, f0, nf);
  stft_fft(z, tw, n, r.log2n, nf, false);
  const int F = (n >> 1) + 1;
  float s1 = 0.f, s2 = 0.f, l1 = 0.f;
  for (int idx = threadIdx.x; idx < nf * F; idx += blockDim.x) {
    const int f = idx / F, k = idx - f * F;
    const float2 zf = z[f * n + k], zc = z[f * n + ((n - k) & (n - 1))];
    float2 xs, ys;
    stft_split(zf, zc, xs, ys);
    const float xm = sqrtf(fmaxf(xs.x * xs.x + xs.y * xs.y, kStftEps)), ym = sqrtf(fmaxf(ys.x * ys.x + ys.y * ys.y, kStftEps));
    const float d = ym - xm;
    s1 = fmaf(d, d, s1);
    s2 = fmaf(ym, ym, s2);
    l1 += fabsf(logf(xm) - logf(ym));
  }
  double a = s1, b = s2, c = l1;
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; red[2][threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(sums + static_cast<size_t>(m) * 3 + threadIdx.x, t);
  }
}

// coefficients of one resolution: c_sc[m] = w_sc * wg[m] / (R * Mg), c_log[m] = w_log * wg[m] / (R * Mg * F * frames)
struct StftCoef { float c_sc, c_log; };

// loss += sum_m c_sc[m] * sqrt(S1 / S2) + c_log[m] * L1   (one block per resolution; atomicAdd into *loss)
__global__ void mrstft_finish_kernel(const double* sums, int M, int Mg, float w_sc, float w_log, float wg0, float wg1, int n_groups,
                                     double inv_R, double bins, float* loss) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const double wg = (n_groups == 2 && m >= Mg) ? wg1 : wg0;
    const double S1 = sums[m * 3], S2 = sums[m * 3 + 1], L1 = sums[m * 3 + 2];
    acc += wg * inv_R / Mg * (w_sc * sqrt(S1) / sqrt(S2) + w_log * L1 / bins);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
    atomicAdd(loss, static_cast<float>(t));
  }
}

// backward of one resolution: gX, gY [M, T] += d loss / d (filtered signals)
__global__ void __launch_bounds__(256) mrstft_bwd_kernel(const float* X, const float* Y, const float* win, StftRes r, long long T,
                                                         const double* sums, int Mg, float w_sc, float w_log, float wg0, float wg1,
                                                         int n_groups, float inv_R, float* gX, float* gY) {
  __shared__ float2 z[kStftMaxN];
  __shared__ float2 tw[kStftMaxN / 2];
  const int n = r.n, fpb = kStftMaxN / n;
  const int m = blockIdx.y, f0 = blockIdx.x * fpb;
  const int nf = min(fpb, r.frames - f0);
  stft_load(z, tw, X, Y, win, r, T, m, f0, nf);
  stft_fft(z, tw, n, r.log2n, nf, false);
  const int F = (n >> 1) + 1;
  const float wg = (n_groups == 2 && m >= Mg) ? wg1 : wg0;
  const double S1 = sums[static_cast<size_t>(m) * 3], S2 = sums[static_cast<size_t>(m) * 3 + 1];
  const float r1 = static_cast<float>(sqrt(S1)), r2 = static_cast<float>(sqrt(S2));
  const float c_sc = w_sc * wg * inv_R / Mg;
  const float c_log = w_log * wg * inv_R / (static_cast<float>(Mg) * F * r.frames);
  const float inv12 = (r1 > 0.f) ? 1.f / (r1 * r2) : 0.f;         // S1 = 0: x == y, the sc term has a zero (sub)gradient
  const float k3 = r1 / (r2 * r2 * r2);
  // bins k and n - k are rewritten together: W = Hx + i Hy with Hx, Hy the Hermitian extensions of the half spectra
  for (int idx = threadIdx.x; idx < nf * F; idx += blockDim.x) {
    const int f = idx / F, k = idx - f * F;
    const int kc = (n - k) & (n - 1);
    const float2 zf = z[f * n + k], zc = z[f * n + kc];
    float2 xs, ys;
    stft_split(zf, zc, xs, ys);
    const float px = xs.x * xs.x + xs.y * xs.y, py = ys.x * ys.x + ys.y * ys.y;
    const float xm = sqrtf(fmaxf(px, kStftEps)), ym = sqrtf(fmaxf(py, kStftEps));
    const float dl = logf(xm) - logf(ym);
    const float sg = (dl > 0.f) ? 1.f : (dl < 0.f ? -1.f : 0.f);
    float gx = c_sc * (xm - ym) * inv12 + c_log * sg / xm;
    float gy = c_sc * ((ym - xm) * inv12 - k3 * ym) - c_log * sg / ym;
    if (!(px > kStftEps)) gx = 0.f;                                  // clamp(min = eps): no gradient below it
    if (!(py > kStftEps)) gy = 0.f;
    const float2 Gx = make_float2(gx * xs.x / xm, gx * xs.y / xm), Gy = make_float2(gy * ys.x / ym, gy * ys.y / ym);
    if (k == 0 || k == (n >> 1)) {
      z[f * n + k] = make_float2(Gx.x, Gy.x);                         // real bins: only the real parts reach the signal
    } else {
      // W[k] = Gx/2 + i Gy/2,  W[n-k] = conj(Gx)/2 + i conj(Gy)/2
      z[f * n + k] = make_float2(0.5f * (Gx.x - Gy.y), 0.5f * (Gx.y + Gy.x));
      z[f * n + kc] = make_float2(0.5f * (Gx.x + Gy.y), 0.5f * (Gy.x - Gx.y));
    }
  }
  __syncthreads();
  // natural -> bit-reversed order in place, then the inverse transform (unnormalised): z[i] = dx[i] + i dy[i]
  for (int idx = threadIdx.x; idx < nf * n; idx += blockDim.x) {
    const int f = idx / n, i = idx - f * n;
    const int ir = static_cast<int>(__brev(static_cast<unsigned>(i)) >> (32 - r.log2n));
    if (i < ir) { const float2 t = z[f * n + i]; z[f * n + i] = z[f * n + ir]; z[f * n + ir] = t; }
  }
  __syncthreads();
  stft_fft(z, tw, n, r.log2n, nf, true);
  for (int idx = threadIdx.x; idx < nf * n; idx += blockDim.x) {
    const int f = idx / n, i = idx - f * n;
    const long long p = stft_reflect(static_cast<long long>(f0 + f) * r.hop + i - (n >> 1), T);
    const float w = __ldg(win + i);
    const float2 v = z[f * n + i];
    if (gX) atomicAdd(gX + static_cast<size_t>(m) * T + p, w * v.x);
    if (gY) atomicAdd(gY + static_cast<size_t>(m) * T + p, w * v.y);
  }
}

}  // namespace kvae
