// Development probe: HBM -> shared -> HBM copy of a [rows, 128] fp16 tensor through per-warp TMA pipelines, the
// transfer pattern of the conv epilogues (skip block in, stream / operand block out).  Question it answers: how much
// HBM bandwidth do 16-row x 64-byte boxes (one TMEM lane quadrant = 32 channels) reach compared with 128- / 256-byte
// rows?   usage: tma_copy_probe <box channels 32|64|128> <box rows> <slots per warp> <warps per CTA> [rows]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <string>

#include "../kalle_audio_b200/csrc/ptx.cuh"
using namespace kvae;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   ptx::smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(ptx::smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}

__global__ void __launch_bounds__(1024, 1)
copy_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmD, int bc, int br, int S,
            long long n_blocks) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  const int warp = ptx::warp_idx(), lane = threadIdx.x & 31, W = blockDim.x >> 5;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);            // [W][S]
  const uint32_t blk_bytes = bc * 2 * br;
  uint8_t* ring = smem + 1024 + static_cast<size_t>(warp) * S * blk_bytes;
  if (threadIdx.x == 0) {
    for (int i = 0; i < W * S; ++i) ptx::mbar_init(&bars[i], 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (lane != 0) return;
  uint64_t* my = bars + warp * S;
  const int cpb = 128 / bc;                                       // channel blocks per row block
  const long long stride = static_cast<long long>(gridDim.x) * W;
  const long long first = static_cast<long long>(blockIdx.x) * W + warp;
  auto issue = [&](long long idx, int slot) {
    ptx::mbar_expect_tx(&my[slot], blk_bytes);
    tma_load_2d(ring + slot * blk_bytes, &tmS, &my[slot], static_cast<int>(idx % cpb) * bc, static_cast<int>(idx / cpb) * br);
  };
  for (int j = 0; j < S; ++j)
    if (first + j * stride < n_blocks) issue(first + j * stride, j);
  uint32_t ph = 0;
  int slot = 0;
  for (long long idx = first; idx < n_blocks; idx += stride) {
    ptx::mbar_wait(&my[slot], (ph >> slot) & 1u);
    ph ^= 1u << slot;
    tma_store_2d(&tmD, ring + slot * blk_bytes, static_cast<int>(idx % cpb) * bc, static_cast<int>(idx / cpb) * br);
    ptx::bulk_commit();
    ptx::bulk_wait_read<0>();
    if (idx + S * stride < n_blocks) issue(idx + S * stride, slot);
    if (++slot == S) slot = 0;
  }
  ptx::bulk_wait<0>();
}

int main(int argc, char** argv) {
  if (argc < 5) { printf("usage: tma_copy_probe <box channels> <box rows> <slots> <warps> [rows]\n"); return 1; }
  const int bc = atoi(argv[1]), br = atoi(argv[2]), S = atoi(argv[3]), W = atoi(argv[4]);
  const long long rows = argc > 5 ? atoll(argv[5]) : 4LL * 442368;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  PFN_enc enc = reinterpret_cast<PFN_enc>(p);
  uint16_t *src, *dst;
  const size_t bytes = static_cast<size_t>(rows) * 256;
  CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&dst, bytes));
  CK(cudaMemset(src, 0x3c, bytes)); CK(cudaMemset(dst, 0, bytes));
  CUtensorMap tmS, tmD;
  cuuint64_t dims[2] = {128, (cuuint64_t)rows};
  cuuint64_t strides[1] = {256};
  cuuint32_t box[2] = {(cuuint32_t)bc, (cuuint32_t)br};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = bc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : (bc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE);
  for (CUtensorMap* m : {&tmS, &tmD}) {
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, m == &tmS ? (void*)src : (void*)dst, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 3; }
  }
  const size_t smem = 2048 + static_cast<size_t>(W) * S * bc * 2 * br;
  if (smem > 227 * 1024) { printf("does not fit shared memory\n"); return 4; }
  CK(cudaFuncSetAttribute(copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const long long n_blocks = (rows / br) * (128 / bc);
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) copy_kernel<<<sms, W * 32, smem>>>(tmS, tmD, bc, br, S, n_blocks);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 5; ++i) copy_kernel<<<sms, W * 32, smem>>>(tmS, tmD, bc, br, S, n_blocks);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
  uint16_t h[4];
  CK(cudaMemcpy(h, dst + (rows - 1) * 128 + 124, 8, cudaMemcpyDeviceToHost));
  printf("TMACOPY box %3d ch x %3d rows (%4d B rows, %5d B boxes) slots %d warps %2d: %.3f ms  %.0f GB/s (read+write)  check %04x\n", bc, br,
         bc * 2, bc * 2 * br, S, W, ms, 2.0 * bytes / ms * 1e-6, h[3]);
  return 0;
}
