// conv_ru: one-kernel ResidualUnit for 128-channel stages (reference autoencoders.py:39-62):
//     x' = x + conv_k1( SnakeBeta2( conv_k7_dil_d( SnakeBeta1(x) ) ) )
// The producer of x already wrote a = SnakeBeta1(x) (bf16 operand) and x (fp32 residual stream); this kernel
// never materialises the intermediate h = SnakeBeta2(conv_k7(a)) in HBM.  Per output row it moves
// 256 B (a) + 512 B (x) in and 512 B (x') + 256 B (a' = next layer's operand) out = 1536 B instead of the
// 2048 B of the two-kernel form -- these stages are HBM-bound, so that is the speed-up.
//
// Orientation (see conv_umma2.cuh "swap"): out-channels on the MMA M side (TMEM lanes), 256 time rows on N.
//   GEMM1  D1[co, t]  = sum_{tap,ci} W7[tap][co, ci] * a[t + shift(tap), ci]      (14 x 4 MMAs of 128x256x16)
//   EPI1   h[t, c]    = bf16( SnakeBeta2( D1[c, t] + bias7[c] ) )  -> shared memory, K-major, 128-row halves
//   GEMM2  D2[co2, t] = sum_c W1[co2, c] * h[t, c]                                 (2 halves x 2 x 4 MMAs of 128x128x16)
//   EPI2   x'[t, c]   = D2[c, t] + bias1[c] + x[t, c]  -> fp32 stream (TMA store); a' = bf16(SnakeBeta_next(x')) (TMA store)
// TMEM: D1 = columns [0,256), D2 = [256,512).  Warp roles (640 threads): 0 TMA producer, 1 UMMA issuer,
// 2 TMEM allocator, 4-19 epilogue.  All 16 epilogue warps (4 per TMEM lane quadrant) do BOTH epilogues, one after
// the other: their EPI1 share of a tile (2 of the 8 16-row blocks of each half), then their EPI2 share (4 of the
// 16 16-row items).  With separate EPI1 / EPI2 warp groups the EPI1 warps idled 70 % of the time waiting for GEMM1
// while the 8 EPI2 warps, one latency-bound item at a time, took ~14 K cycles per tile and held D2 back
// (profiles/r01_conv_ru_unified_epilogue_B16.ncu-rep, warp-stall samples); now EPI2 of tile i runs on 16 warps
// underneath GEMM1 of tile i+1.  Every warp prefetches its next skip-connection block by TMA and stores in place.
// The residual stream is fp16 in HBM (inference plans only use this kernel).
// Shared memory (227 KB): activation slab ring 2 x <=40 KB, weight ring 3-4 x 16 KB (W7 taps, then the two W1
// chunks of the tile), h half 32 KB, epilogue staging 16 x 3 KB.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

#include "conv_umma2.cuh"

namespace kvae {

struct RuParams {
  int B, T;                 // rows per clip (stride-1 conv: T_in == T_out)
  int RB, nbox;             // activation slab = nbox TMA boxes of RB rows (256 + 6*dilation rows)
  int SA, SB;
  int q_tiles, total_tiles; // 256-row tiles
  uint32_t tap_shift16[7];  // (tap*dilation*128 B) >> 4
  int slab_row0;            // slab start row relative to the tile's first row (= -3*dilation)
  const float* bias7;       // [128]
  const float* s2_a;        // SnakeBeta between the two convs
  const float* s2_inv_b;
  const float* bias1;       // [128]
  int raw_out;              // 1: fp32 stream out via tmR
  int act_out;              // 1: bf16 operand out via tmO
  const float* sn_a;        // SnakeBeta folded into the operand output (nullptr: plain cast)
  const float* sn_inv_b;
  int raw_f16;              // 1: residual stream (x in, x' out) is fp16 in HBM, 0: fp32
  int dbg;                  // ablation switches for tools/umma_probe (KVAE_RU_DBG); 0 in production
  const void* a_ptr;        // conv_ru2_kernel: base addresses of the operand / stream tensors for L2 prefetch
  const void* x_ptr;
  int pf;                   // bit 0: prefetch the operand rows of the tile after next into L2, bit 1: its skip rows
  int k0, k1;               // conv_ru2_kernel: GEMM2 half 0 / 1 of tile i goes in before slab k0 / k1 of tile i+1's GEMM1
  int act_f16;              // conv_ru2_kernel: the operand output is fp16 instead of bf16 (the unit in front of the decoder tail,
                            // whose GEMM is kind::f16 -- conv_edge.cuh, WaveOutTcParams::preact)
};

constexpr int kRuC = 128;
constexpr int kRuHBytes = 2 * 128 * 128;   // h half: 2 K-chunks x [128 rows x 128 B]

constexpr int kRuRawBlk = 16 * 128;        // 16 rows x 32 fp32, SWIZZLE_128B
constexpr int kRuActBlk = 16 * 64;         // 16 rows x 32 bf16, SWIZZLE_64B
constexpr int kRuThreads = 640;

constexpr int kRuEpiWarps = 16;
// per epilogue warp: 2 fp16 stream blocks (skip in / stream out, in place) + 1 bf16 operand block
__host__ __device__ inline int ru_stage_bytes_per_warp(int act_out) { return 2 * kRuActBlk + (act_out ? kRuActBlk : 0); }
__host__ __device__ inline size_t ru_smem_bytes(const RuParams& p) {
  // (dbg 131 = barrier-protocol-only epilogue without skip loads: the staging area is never touched, so ring-depth
  // experiments may spend it on weight stages)
  const size_t staging = (p.dbg & 131) == 131 ? 0 : kRuEpiWarps * ru_stage_bytes_per_warp(p.act_out);
  return 1024 + 1024 + static_cast<size_t>(p.SA) * p.nbox * p.RB * 128 + static_cast<size_t>(p.SB) * kRuC * 128 +
         kRuHBytes + staging;
}

// kEpi selects the epilogue register mapping: 0 = one channel per thread (tcgen05.ld.32x32b, 2-byte shared-memory
// accesses), 1 = mma-style fragments (tcgen05.ld.16x256b + ldmatrix/stmatrix.trans + packed fp32 pairs).
template <int kEpi>
__global__ void __launch_bounds__(kRuThreads, 1)
conv_ru_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW7,
               const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmX,
               const __grid_constant__ RuParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* a_empty = a_full + 8;
  uint64_t* b_full = a_full + 16;
  uint64_t* b_empty = a_full + 32;
  uint64_t* d1_full = a_full + 48;
  uint64_t* d1_empty = a_full + 49;
  uint64_t* h_full = a_full + 50;
  uint64_t* h_empty = a_full + 51;
  uint64_t* d2_full = a_full + 52;
  uint64_t* d2h_full = a_full + 46;   // GEMM2 of half 0 has retired: rows 0..127 of D2 are final (EPI2 starts on them)
  uint64_t* d2_empty = a_full + 53;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 54);
  uint64_t* res_full = a_full + 56;   // [16 epilogue warps][2 slots]
  uint8_t* a_ring = smem + 1024;
  const uint32_t a_bytes = static_cast<uint32_t>(p.nbox) * p.RB * 128;
  constexpr uint32_t b_bytes = kRuC * 128;
  uint8_t* b_ring = a_ring + static_cast<size_t>(p.SA) * a_bytes;
  uint8_t* h_buf = b_ring + static_cast<size_t>(p.SB) * b_bytes;
  uint8_t* stage_base = h_buf + kRuHBytes;

  const int warp = ptx::warp_idx();
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW7);
    ptx::prefetch_tmap(&tmW1);
    ptx::prefetch_tmap(&tmX);
    if (p.raw_out) ptx::prefetch_tmap(&tmR);
    if (p.act_out) ptx::prefetch_tmap(&tmO);
    for (int i = 0; i < p.SA; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < p.SB; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], 1); }
    ptx::mbar_init(d1_full, 1);
    ptx::mbar_init(d1_empty, kRuEpiWarps);
    ptx::mbar_init(h_full, kRuEpiWarps);
    ptx::mbar_init(h_empty, 1);
    ptx::mbar_init(d2_full, 1);
    ptx::mbar_init(d2h_full, 1);
    ptx::mbar_init(d2_empty, kRuEpiWarps);
    for (int i = 0; i < 2 * kRuEpiWarps; ++i) ptx::mbar_init(&res_full[i], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t d1 = tmem_base, d2 = tmem_base + 256;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int b = tile / p.q_tiles;
        const int q0 = (tile % p.q_tiles) * 256;
        for (int ch = 0; ch < 2; ++ch) {
          ptx::mbar_wait(&a_empty[as], aph ^ 1u);
          if (p.dbg & 64) {                                  // ablation: no activation loads
            ptx::mbar_arrive(&a_full[as]);
          } else {
            ptx::mbar_expect_tx(&a_full[as], a_bytes);
            for (int bx = 0; bx < p.nbox; ++bx)
              ptx::tma_load_4d(a_ring + static_cast<size_t>(as) * a_bytes + bx * p.RB * 128, &tmA, &a_full[as], ch * 64, 0,
                               q0 + p.slab_row0 + bx * p.RB, b);
          }
          if (++as == p.SA) { as = 0; aph ^= 1u; }
          for (int t = 0; t < 7; ++t) {
            ptx::mbar_wait(&b_empty[bs], bph ^ 1u);
            if (p.dbg & 32) {                                // ablation: no weight loads
              ptx::mbar_arrive(&b_full[bs]);
            } else {
              ptx::mbar_expect_tx(&b_full[bs], b_bytes);
              ptx::tma_load_3d(b_ring + static_cast<size_t>(bs) * b_bytes, &tmW7, &b_full[bs], ch * 64, 0, t);
            }
            if (++bs == p.SB) { bs = 0; bph ^= 1u; }
          }
        }
        for (int kc = 0; kc < 2; ++kc) {   // the k=1 conv's weights ride the same ring
          ptx::mbar_wait(&b_empty[bs], bph ^ 1u);
          if (p.dbg & 32) {
            ptx::mbar_arrive(&b_full[bs]);
          } else {
            ptx::mbar_expect_tx(&b_full[bs], b_bytes);
            ptx::tma_load_3d(b_ring + static_cast<size_t>(bs) * b_bytes, &tmW1, &b_full[bs], kc * 64, 0, 0);
          }
          if (++bs == p.SB) { bs = 0; bph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ UMMA issuer
    if (ptx::elect_one()) {
      const uint32_t idesc1 = ptx::idesc_bf16_f32(128, 256);
      const uint32_t idesc2 = ptx::idesc_bf16_f32(128, 128);
      const uint64_t desc_hi = (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
      const uint32_t a_lo0 = ((ptx::smem_u32(a_ring) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_lo0 = ((ptx::smem_u32(b_ring) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t h_lo0 = ((ptx::smem_u32(h_buf) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t a_stage16 = a_bytes >> 4, b_stage16 = b_bytes >> 4;
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0, d1e_ph = 0, d2e_ph = 0, hf_ph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        // ---- GEMM1 into D1
        ptx::mbar_wait(d1_empty, d1e_ph ^ 1u);
        d1e_ph ^= 1u;
        ptx::tc_fence_after();
        uint32_t accum = 0;
        bool ready = false;
        for (int ch = 0; ch < 2; ++ch) {
          ptx::mbar_wait(&a_full[as], aph);
          const int cur = as;
          const uint32_t cur_lo = a_lo0 + as * a_stage16;
          if (++as == p.SA) { as = 0; aph ^= 1u; }
#pragma unroll 1
          for (int t = 0; t < 7; ++t) {
            if (!ready) ptx::mbar_wait(&b_full[bs], bph);
            ptx::tc_fence_after();
            const uint32_t al = cur_lo + p.tap_shift16[t];
            const uint32_t bl = b_lo0 + bs * b_stage16;
            const int cb = bs;
            if (++bs == p.SB) { bs = 0; bph ^= 1u; }
            ready = ptx::mbar_try_wait(&b_full[bs], bph);      // poll the NEXT stage while this tap's MMAs issue
            if (!(p.dbg & 16) || t == 0) {
            ptx::umma_f16(d1, desc_hi | bl, desc_hi | al, idesc1, accum);
            ptx::umma_f16(d1, desc_hi | (bl + 2), desc_hi | (al + 2), idesc1, 1u);
            ptx::umma_f16(d1, desc_hi | (bl + 4), desc_hi | (al + 4), idesc1, 1u);
            ptx::umma_f16(d1, desc_hi | (bl + 6), desc_hi | (al + 6), idesc1, 1u);
            }
            accum = 1u;
            ptx::umma_commit(&b_empty[cb]);
          }
          ptx::umma_commit(&a_empty[cur]);
        }
        ptx::umma_commit(d1_full);
        // ---- GEMM2 into D2, one 128-row half at a time as EPI1 hands over h
        const int s0 = bs;
        const uint32_t ph0 = bph;
        if (++bs == p.SB) { bs = 0; bph ^= 1u; }
        const int s1 = bs;
        const uint32_t ph1 = bph;
        if (++bs == p.SB) { bs = 0; bph ^= 1u; }
        ptx::mbar_wait(&b_full[s0], ph0);
        ptx::mbar_wait(&b_full[s1], ph1);
        ptx::mbar_wait(d2_empty, d2e_ph ^ 1u);
        d2e_ph ^= 1u;
        const uint32_t w_lo[2] = {b_lo0 + s0 * b_stage16, b_lo0 + s1 * b_stage16};
        for (int half = 0; half < 2; ++half) {
          ptx::mbar_wait(h_full, hf_ph);
          hf_ph ^= 1u;
          ptx::tc_fence_after();
          const uint32_t dd = d2 + half * 128;
#pragma unroll
          for (int kc = 0; kc < 2; ++kc) {
            const uint32_t hl = h_lo0 + kc * (16384 >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_f16(dd, desc_hi | (w_lo[kc] + 2 * k), desc_hi | (hl + 2 * k), idesc2, (kc | k) ? 1u : 0u);
          }
          ptx::umma_commit(h_empty);
          if (half == 0) ptx::umma_commit(d2h_full);
        }
        ptx::umma_commit(&b_empty[s0]);
        ptx::umma_commit(&b_empty[s1]);
        ptx::umma_commit(d2_full);
      }
    }
  } else if (warp >= 4 && kEpi == 1) {
    // ------------------------------------------------------------ epilogue warps, fragment mapping (see ptx.cuh,
    // tmem_ld_16x256b_*): a thread holds (time, time+1) PAIRS of one channel, so bias / SnakeBeta run as packed
    // fp32 pairs, a converted pair is one stmatrix.trans register, and the [time][channel] tiles in shared memory
    // are read and written as 16-byte rows (ldmatrix / stmatrix) instead of one 2-byte access per element.
    const int e = warp - 4;                            // 0..15
    const int quad = warp & 3;                         // TMEM lane quadrant = channels quad*32 .. +31
    const int sub = e >> 2;                            // 0..3
    const int g = lane >> 2;                           // fragment row: channel within a group of 8
    const int rr = lane & 7, mj = lane >> 3;           // ldmatrix/stmatrix: this lane addresses row rr of matrix mj
    const int cbase = quad * 32;
    auto dup = [](float f) { return ptx::f2_pack(f, f); };
    // EPI1 share of a 128-column half: 16 channels (lane half hl) x 64 time columns (block cb)
    const int hl = sub & 1, cb = sub >> 1;
    const int c1 = cbase + hl * 16 + g;                // this thread's EPI1 channels: c1 and c1 + 8
    const uint32_t d1_lane = d1 + (static_cast<uint32_t>(cbase + hl * 16) << 16) + cb * 64;
    const int hch = cbase + hl * 16 + (mj & 1) * 8;    // first channel of the 8-channel group of matrix mj
    const uint32_t h_lane = ptx::smem_u32(h_buf) + (hch >> 6) * 16384 + (cb * 64 + (mj >> 1) * 8 + rr) * 128 +
                            ((((hch & 63) >> 3) ^ rr) << 4);
    // EPI2 staging: fp16 stream blocks [16 rows x 64 B] and the bf16 operand block, SWIZZLE_64B; lane half 1 = ^ 32
    uint8_t* ring = stage_base + e * ru_stage_bytes_per_warp(p.act_out);
    const int brow = (mj >> 1) * 8 + rr;
    const uint32_t blk_lane = brow * 64 + (((mj & 1) ^ ((brow >> 1) & 3)) << 4);
    const uint32_t ring_lane = ptx::smem_u32(ring) + blk_lane;
    const uint32_t ablk_lane = ring_lane + 2 * kRuActBlk;
    uint8_t* ablk = ring + 2 * kRuActBlk;
    uint64_t* my_res_full = res_full + e * 2;
    const bool use_skip = !(p.dbg & 1);
    const bool has_snake = p.sn_a != nullptr;
    auto issue_skip = [&](int tb, int tq0, int item, int slot) {   // lane 0 only; (tb, tq0) = clip, first row of the tile
      ptx::mbar_expect_tx(&my_res_full[slot], kRuActBlk);
      ptx::tma_load_4d(ring + slot * kRuActBlk, &tmX, &my_res_full[slot], cbase, 0, tq0 + item * 16, tb);
    };
    if (lane == 0 && use_skip && static_cast<int>(blockIdx.x) < p.total_tiles)
      issue_skip(blockIdx.x / p.q_tiles, (blockIdx.x % p.q_tiles) * 256, sub, 0);
    int slot = 0;
    uint32_t d1f_ph = 0, he_ph = 0, d2f_ph = 0, res_ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int b = tile / p.q_tiles;
      const int q0 = (tile % p.q_tiles) * 256;
      const int ntile = tile + gridDim.x;                // this CTA's next tile (its first skip block is prefetched below)
      const int nb = ntile / p.q_tiles, nq0 = (ntile % p.q_tiles) * 256;
      // ---- EPI1 share: D1 -> bias, SnakeBeta -> h
      {
        const uint64_t kb[2] = {dup(__ldg(p.bias7 + c1)), dup(__ldg(p.bias7 + c1 + 8))};
        const uint64_t ka[2] = {dup(__ldg(p.s2_a + c1)), dup(__ldg(p.s2_a + c1 + 8))};
        const uint64_t kib[2] = {dup(__ldg(p.s2_inv_b + c1)), dup(__ldg(p.s2_inv_b + c1 + 8))};
        ptx::mbar_wait_parked(d1_full, d1f_ph);
        d1f_ph ^= 1u;
        ptx::tc_fence_after();
        for (int half = 0; half < 2; ++half) {
          uint32_t r[32], pk[16];
          __syncwarp();
          if (!(p.dbg & 128)) {
          ptx::tmem_ld_16x256b_x8(d1_lane + half * 128, r);
          ptx::tmem_ld_wait();
          }
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            if (p.dbg & 128) { pk[q] = 0u; continue; }       // ablation: barrier protocol only                 // q = 2n + hi: column group n, channel c1 + 8 hi
            const uint64_t v = ptx::f2_add(ptx::f2_pack_u(r[2 * q], r[2 * q + 1]), kb[q & 1]);
            float t0, t1, y0, y1;
            ptx::f2_unpack(ptx::f2_mul(v, ka[q & 1]), t0, t1);
            const uint64_t sn = ptx::f2_pack(sin_fast(t0), sin_fast(t1));
            ptx::f2_unpack(ptx::f2_fma(ptx::f2_mul(kib[q & 1], sn), sn, v), y0, y1);
            const __nv_bfloat162 hb2 = __floats2bfloat162_rn(y0, y1);
            pk[q] = *reinterpret_cast<const uint32_t*>(&hb2);
          }
          ptx::mbar_wait_parked(h_empty, he_ph ^ 1u);      // GEMM2 of the previous half has consumed h
          he_ph ^= 1u;
#pragma unroll
          for (int m = 0; m < 4; ++m)
            if (!(p.dbg & 128)) ptx::stmatrix_x4_trans(h_lane + m * 2048, pk[4 * m], pk[4 * m + 1], pk[4 * m + 2], pk[4 * m + 3]);
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(h_full);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(d1_empty);       // GEMM1 of the next tile may start: it runs under EPI2 below
      // ---- EPI2 share: D2 -> bias, + skip -> stream / operand out
      uint64_t kb[4], ka[4], kib[4];                     // index 2 * lane half + hi: channel cbase + 16 L + 8 hi + g
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ch = cbase + (i >> 1) * 16 + (i & 1) * 8 + g;
        kb[i] = dup(__ldg(p.bias1 + ch));
        ka[i] = has_snake ? dup(__ldg(p.sn_a + ch)) : 0ull;
        kib[i] = has_snake ? dup(__ldg(p.sn_inv_b + ch)) : 0ull;
      }
      ptx::mbar_wait_parked(d2h_full, d2f_ph);           // rows 0..127 first: GEMM2 of half 1 may still be running
      ptx::tc_fence_after();
#pragma unroll 1
      for (int item = sub; item < 16; item += 4) {
        const int r0 = q0 + item * 16;
        if (item == sub + 8) {                             // rows 128..255
          ptx::mbar_wait_parked(d2_full, d2f_ph);
          ptx::tc_fence_after();
        }
        if (lane == 0) {
          // every store issued so far has read its shared-memory source: the other stream slot and the operand
          // block are free again; fetch the NEXT item's skip block into the other slot
          ptx::bulk_wait_read<0>();
          if (use_skip) {
            if (item + 4 < 16) issue_skip(b, q0, item + 4, slot ^ 1);
            else if (ntile < p.total_tiles) issue_skip(nb, nq0, sub, slot ^ 1);
          }
        }
        if (p.dbg & 128) { if (use_skip) ptx::mbar_wait_parked(&my_res_full[slot], (res_ph >> slot) & 1u); res_ph ^= (1u << slot); slot ^= 1; continue; }
        uint32_t r[16], sk[8];
        __syncwarp();
        const uint32_t t2 = d2 + (static_cast<uint32_t>(cbase) << 16) + item * 16;
        ptx::tmem_ld_16x256b_x2(t2, r);
        ptx::tmem_ld_16x256b_x2(t2 + (16u << 16), r + 8);
        const uint32_t rb = ring_lane + slot * kRuActBlk;
        if (use_skip) {
          ptx::mbar_wait_parked(&my_res_full[slot], (res_ph >> slot) & 1u);
          ptx::ldmatrix_x4_trans(rb, sk[0], sk[1], sk[2], sk[3]);
          ptx::ldmatrix_x4_trans(rb ^ 32u, sk[4], sk[5], sk[6], sk[7]);
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) sk[q] = 0u;
        }
        res_ph ^= (1u << slot);
        ptx::tmem_ld_wait();
        uint64_t v[8];                                     // q = 4 L + 2 n + hi
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float2 sf = __half22float2(*reinterpret_cast<const __half2*>(&sk[q]));
          v[q] = ptx::f2_add(ptx::f2_add(ptx::f2_pack_u(r[2 * q], r[2 * q + 1]), kb[(q >> 2) * 2 + (q & 1)]),
                             ptx::f2_pack(sf.x, sf.y));
        }
        if (p.raw_out) {
          uint32_t w[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float y0, y1;
            ptx::f2_unpack(v[q], y0, y1);
            w[q] = ptx::f2h2_sat(y0, y1);
          }
          ptx::stmatrix_x4_trans(rb, w[0], w[1], w[2], w[3]);
          ptx::stmatrix_x4_trans(rb ^ 32u, w[4], w[5], w[6], w[7]);
        }
        if (p.act_out) {
          uint32_t w[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float y0, y1;
            if (has_snake) {
              const int ci = (q >> 2) * 2 + (q & 1);
              float t0, t1;
              ptx::f2_unpack(ptx::f2_mul(v[q], ka[ci]), t0, t1);
              const uint64_t sn = ptx::f2_pack(sin_fast(t0), sin_fast(t1));
              ptx::f2_unpack(ptx::f2_fma(ptx::f2_mul(kib[ci], sn), sn, v[q]), y0, y1);
            } else {
              ptx::f2_unpack(v[q], y0, y1);
            }
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(y0, y1);
            w[q] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          ptx::stmatrix_x4_trans(ablk_lane, w[0], w[1], w[2], w[3]);
          ptx::stmatrix_x4_trans(ablk_lane ^ 32u, w[4], w[5], w[6], w[7]);
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          uint8_t* const rblk = ring + slot * kRuActBlk;
          if (p.raw_out && !(p.dbg & 2)) ptx::tma_store_4d(&tmR, rblk, cbase, 0, r0, b);
          if (p.act_out && !(p.dbg & 2)) ptx::tma_store_4d(&tmO, ablk, cbase, 0, r0, b);
          ptx::bulk_commit();
        }
        slot ^= 1;
      }
      d2f_ph ^= 1u;
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(d2_empty);
    }
    if (lane == 0) ptx::bulk_wait<0>();
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue warps (kEpi == 0): EPI1 share, then EPI2 share, per tile
    const int e = warp - 4;                            // 0..15
    const int quad = warp & 3;                         // TMEM lane quadrant = channels quad*32 .. +31
    const int sub = e >> 2;                            // 0..3: which blocks / items of the quadrant this warp takes
    const int c = quad * 32 + lane;                    // this thread's channel
    const int cbase = quad * 32;
    const float bias7 = __ldg(p.bias7 + c);
    const float s2a = __ldg(p.s2_a + c), s2ib = __ldg(p.s2_inv_b + c);
    const float bias1 = __ldg(p.bias1 + c);
    float sa = 1.f, sib = 0.f;
    if (p.sn_a) { sa = __ldg(p.sn_a + c); sib = __ldg(p.sn_inv_b + c); }
    // EPI1: element (row r, channel c) of the K-major SWIZZLE_128B h tile: chunk tile c/64, 16-byte group (c%64)/8
    uint8_t* hb[8];
#pragma unroll
    for (int x = 0; x < 8; ++x)
      hb[x] = h_buf + (c >> 6) * 16384 + (((((c & 63) >> 3)) ^ x) << 4) + (c & 7) * 2;
    // EPI2 staging: fp16 stream blocks [16 rows x 64 B] and the bf16 operand block, SWIZZLE_64B
    uint8_t* ring = stage_base + e * ru_stage_bytes_per_warp(p.act_out);
    uint8_t* ablk = ring + 2 * kRuActBlk;
    uint64_t* my_res_full = res_full + e * 2;
    const uint32_t acol = (lane & 7) * 2, achunk = lane >> 3;
    const bool use_skip = !(p.dbg & 1);
    auto issue_skip = [&](int tile, int item, int slot) {   // lane 0 only
      if (tile < p.total_tiles && use_skip) {
        ptx::mbar_expect_tx(&my_res_full[slot], kRuActBlk);
        ptx::tma_load_4d(ring + slot * kRuActBlk, &tmX, &my_res_full[slot], cbase, 0, (tile % p.q_tiles) * 256 + item * 16,
                         tile / p.q_tiles);
      }
    };
    if (lane == 0) issue_skip(blockIdx.x, sub, 0);
    int slot = 0;
    uint32_t d1f_ph = 0, he_ph = 0, d2f_ph = 0, res_ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int b = tile / p.q_tiles;
      const int q0 = (tile % p.q_tiles) * 256;
      // ---- EPI1 share: D1 -> bias, SnakeBeta -> h
      ptx::mbar_wait_parked(d1_full, d1f_ph);
      d1f_ph ^= 1u;
      ptx::tc_fence_after();
      for (int half = 0; half < 2; ++half) {
        {
          // this warp's two 16-row blocks of the half: both TMEM loads in flight, the arithmetic done in registers
          // BEFORE waiting for the h buffer (GEMM2 of the previous half may still be reading it), then the stores
          uint32_t r0[16], r1[16];
          __syncwarp();
          const uint32_t t0 = d1 + (static_cast<uint32_t>(quad * 32) << 16) + half * 128 + sub * 16;
          ptx::tmem_ld_32x16(t0, r0);
          ptx::tmem_ld_32x16(t0 + 64, r1);
          ptx::tmem_ld_wait();
          __nv_bfloat16 v0[16], v1[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a = (p.dbg & 4) ? __uint_as_float(r0[j]) + bias7 : snake_beta<true>(__uint_as_float(r0[j]) + bias7, s2a, s2ib);
            const float b = (p.dbg & 4) ? __uint_as_float(r1[j]) + bias7 : snake_beta<true>(__uint_as_float(r1[j]) + bias7, s2a, s2ib);
            v0[j] = __float2bfloat16(a);
            v1[j] = __float2bfloat16(b);
          }
          ptx::mbar_wait_parked(h_empty, he_ph ^ 1u);      // GEMM2 of the previous half has consumed h
          he_ph ^= 1u;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            *reinterpret_cast<__nv_bfloat16*>(hb[j & 7] + (sub * 16 + j) * 128) = v0[j];
            *reinterpret_cast<__nv_bfloat16*>(hb[j & 7] + ((sub + 4) * 16 + j) * 128) = v1[j];
          }
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(h_full);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(d1_empty);       // GEMM1 of the next tile may start: it runs under EPI2 below
      // ---- EPI2 share: D2 -> bias, + skip -> stream / operand out
      ptx::mbar_wait_parked(d2h_full, d2f_ph);           // rows 0..127 first: GEMM2 of half 1 may still be running
      ptx::tc_fence_after();
#pragma unroll 1
      for (int item = sub; item < 16; item += 4) {
        const int r0 = q0 + item * 16;
        if (item == sub + 8) {                             // rows 128..255
          ptx::mbar_wait_parked(d2_full, d2f_ph);
          ptx::tc_fence_after();
        }
        if (lane == 0) {
          // every store issued so far has read its shared-memory source: the other stream slot and the operand
          // block are free again; fetch the NEXT item's skip block into the other slot
          ptx::bulk_wait_read<0>();
          int nt = tile, ni = item + 4;
          if (ni >= 16) { nt = tile + gridDim.x; ni = sub; }
          issue_skip(nt, ni, slot ^ 1);
        }
        if (use_skip) ptx::mbar_wait_parked(&my_res_full[slot], (res_ph >> slot) & 1u);
        res_ph ^= (1u << slot);
        uint32_t r[16];
        __syncwarp();
        ptx::tmem_ld_32x16(d2 + (static_cast<uint32_t>(quad * 32) << 16) + item * 16, r);
        ptx::tmem_ld_wait();
        uint8_t* const rblk = ring + slot * kRuActBlk;
        uint8_t* hbase[4];                             // the 4 swizzle phases of a SWIZZLE_64B block
#pragma unroll
        for (int x = 0; x < 4; ++x) hbase[x] = rblk + ((achunk ^ x) << 4) + acol;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          v[j] = __uint_as_float(r[j]) + bias1 + __half2float(*reinterpret_cast<const __half*>(hbase[(j >> 1) & 3] + j * 64));
        if (p.raw_out) {
#pragma unroll
          for (int j = 0; j < 16; ++j) *reinterpret_cast<uint16_t*>(hbase[(j >> 1) & 3] + j * 64) = ptx::f2h_sat(v[j]);
        }
        if (p.act_out) {
          if (p.sn_a && !(p.dbg & 8)) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = snake_beta<true>(v[j], sa, sib);
          }
          uint8_t* abase[4];
#pragma unroll
          for (int x = 0; x < 4; ++x) abase[x] = ablk + ((achunk ^ x) << 4) + acol;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            *reinterpret_cast<__nv_bfloat16*>(abase[(j >> 1) & 3] + j * 64) = __float2bfloat16(v[j]);
        }
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (p.raw_out && !(p.dbg & 2)) ptx::tma_store_4d(&tmR, rblk, cbase, 0, r0, b);
          if (p.act_out && !(p.dbg & 2)) ptx::tma_store_4d(&tmO, ablk, cbase, 0, r0, b);
          ptx::bulk_commit();
        }
        slot ^= 1;
      }
      d2f_ph ^= 1u;
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(d2_empty);
    }
    if (lane == 0) ptx::bulk_wait<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace kvae
