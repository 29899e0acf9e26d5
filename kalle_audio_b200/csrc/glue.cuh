// LM <-> VAE glue of one autoregressive step (reference model_sigmaVAE.py:123-145, Llasa.infer):
//     mean   = distribution_linear(last_hidden)        Linear(H -> D) -> GELU (erf) -> Linear(D -> D)     (:42-50, :126)
//     latent = sample(mean, 'fix') = mean + std*noise   two separately rounded operations, as torch         (:153-157)
//     kl_end = KL( N(mean, std) || N(1, e) ).sum(-1) / D   the stop criterion                               (:134-139)
//     embed  = audio_linear(latent)                     Linear(D -> H), the next input embedding            (:143)
// In the reference this is ~20 eager launches per generated frame around three tiny GEMVs (batch 1).  Here it is ONE
// launch: a thread-block cluster of 8 CTAs per batch row splits every matrix by rows / columns, so the ~1 MB of
// fp32 weights is pulled from L2 by 8 SMs at once, and the three stages exchange their small vectors through
// distributed shared memory (cluster barriers instead of kernel boundaries).
//   stage 1: CTA r multiplies its H/8 slice of the hidden row with the matching columns of W1 -> D partial sums;
//            after a cluster barrier CTA r reduces the 8 partials of ITS D/8 outputs, adds the bias, applies GELU and
//            writes the slice into every peer's copy of g
//   stage 2: CTA r computes its D/8 rows of W2 g + b2 = mean, draws the latent, its share of the KL sum, and writes
//            the latent slice into every peer
//   stage 3: CTA r computes its H/8 rows of Wa latent + ba
#pragma once
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "elementwise.cuh"

namespace kvae {

constexpr int kGlueCluster = 8;
constexpr int kGlueThreads = 256;

struct GlueParams {
  const void* hidden;     // [B, H]
  int hidden_f32;
  const float* w1;        // [D, H]
  const float* b1;        // [D]
  const float* w2;        // [D, D]
  const float* b2;        // [D]
  const float* wa;        // [H, D]
  const float* ba;        // [H]
  const void* noise;      // [B, D] in the output dtype
  void* mean;             // [B, D]
  void* latent;           // [B, D]
  void* embed;            // [B, H]
  float* kl_end;          // [B]
  int out_f32;
  int H, D;
  float std;
};

__host__ __device__ inline size_t glue_smem_bytes(int H, int D) {
  return static_cast<size_t>(H / kGlueCluster + 3 * D + kGlueCluster) * sizeof(float);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kGlueThreads, 1) lm_glue_step_kernel(const GlueParams p) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float gsm[];
  const int r = static_cast<int>(cluster.block_rank());
  const int b = blockIdx.x / kGlueCluster;
  const int Hs = p.H / kGlueCluster, Ds = p.D / kGlueCluster;
  float* hid = gsm;                 // [Hs]   this CTA's slice of the hidden row
  float* part = hid + Hs;           // [D]    stage-1 partial sums over this CTA's H slice
  float* g = part + p.D;            // [D]    GELU(W1 h + b1), assembled from all CTAs
  float* lat = g + p.D;             // [D]    sampled latent, assembled from all CTAs
  float* klp = lat + p.D;           // [8]    per-CTA KL partial sums (rank 0's copy is the one that is read)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarps = kGlueThreads / 32;

  for (int i = threadIdx.x; i < Hs; i += kGlueThreads)
    hid[i] = ld_elem(p.hidden, static_cast<size_t>(b) * p.H + r * Hs + i, p.hidden_f32);
  __syncthreads();
  // ---- stage 1: partial W1 h over the H slice, all D outputs
  for (int d = warp; d < p.D; d += kWarps) {
    const float* w = p.w1 + static_cast<size_t>(d) * p.H + r * Hs;
    float acc = 0.f;
    for (int i = lane; i < Hs; i += 32) acc = fmaf(__ldg(w + i), hid[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) part[d] = acc;
  }
  cluster.sync();
  // reduce this CTA's D slice over the 8 partials (fixed order: deterministic), bias, GELU, broadcast
  for (int i = threadIdx.x; i < Ds; i += kGlueThreads) {
    const int d = r * Ds + i;
    float acc = __ldg(p.b1 + d);
    for (int q = 0; q < kGlueCluster; ++q) acc += cluster.map_shared_rank(part, q)[d];
    const float ge = 0.5f * acc * (1.f + erff(acc * 0.70710678118654752f));
    for (int q = 0; q < kGlueCluster; ++q) cluster.map_shared_rank(g, q)[d] = ge;
  }
  cluster.sync();
  // ---- stage 2: this CTA's rows of W2 g + b2 -> mean, latent, KL share
  float kl_acc = 0.f;
  for (int i = warp; i < Ds; i += kWarps) {
    const int d = r * Ds + i;
    const float* w = p.w2 + static_cast<size_t>(d) * p.D;
    float acc = 0.f;
    for (int j = lane; j < p.D; j += 32) acc = fmaf(__ldg(w + j), g[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      float m = acc + __ldg(p.b2 + d);
      const size_t o = static_cast<size_t>(b) * p.D + d;
      if (!p.out_f32) m = __bfloat162float(__float2bfloat16(m));       // the Linear's output in the module dtype
      st_elem(p.mean, o, p.out_f32, m);
      float t = __fmul_rn(p.std, ld_elem(p.noise, o, p.out_f32));      // torch: std * randn_like(mean), then mean + .
      if (!p.out_f32) t = __bfloat162float(__float2bfloat16(t));
      float z = __fadd_rn(m, t);
      if (!p.out_f32) z = __bfloat162float(__float2bfloat16(z));
      st_elem(p.latent, o, p.out_f32, z);
      for (int q = 0; q < kGlueCluster; ++q) cluster.map_shared_rank(lat, q)[d] = z;
      // KL( N(m, s) || N(1, e) ) = log(e / s) + (s^2 + (m - 1)^2) / (2 e^2) - 1/2
      const float e2 = 7.38905609893065f;
      kl_acc += (1.f - logf(p.std)) + (p.std * p.std + (m - 1.f) * (m - 1.f)) / (2.f * e2) - 0.5f;
    }
  }
  // lane 0 of every warp holds a share: reduce across the CTA through shared memory, then to rank 0
  __shared__ float klw[kWarps];
  if (lane == 0) klw[warp] = kl_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kWarps; ++w) s += klw[w];
    cluster.map_shared_rank(klp, 0)[r] = s;
  }
  cluster.sync();
  if (r == 0 && threadIdx.x == 0 && p.kl_end) {
    float s = 0.f;
    for (int q = 0; q < kGlueCluster; ++q) s += klp[q];
    p.kl_end[b] = s / static_cast<float>(p.D);
  }
  // ---- stage 3: this CTA's rows of Wa latent + ba
  for (int i = warp; i < Hs; i += kWarps) {
    const int h = r * Hs + i;
    const float* w = p.wa + static_cast<size_t>(h) * p.D;
    float acc = 0.f;
    for (int j = lane; j < p.D; j += 32) acc = fmaf(__ldg(w + j), lat[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) st_elem(p.embed, static_cast<size_t>(b) * p.H + h, p.out_f32, acc + __ldg(p.ba + h));
  }
}

inline cudaError_t launch_lm_glue_step(const GlueParams& p, int B, cudaStream_t st) {
  const size_t smem = glue_smem_bytes(p.H, p.D);
  cudaError_t e = cudaFuncSetAttribute(lm_glue_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(B * kGlueCluster);
  cfg.blockDim = dim3(kGlueThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kGlueCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, lm_glue_step_kernel, p);
}

}  // namespace kvae
