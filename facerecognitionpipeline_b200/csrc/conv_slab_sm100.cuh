// conv_slab_sm100.cuh — 3x3 stride-1 convolution with ACTIVATION SLAB REUSE on CTA pairs.
//
// profiles/r01b + tools/microbench_gemm.py: the im2col implicit GEMM re-fetches every input pixel
// nine times (once per filter tap) and, for Cout <= 128, is bound by L2->SM operand ingest
// (~42 B/clk/SM), not by the tensor pipe: 377 TF (Cout 64) / 718 TF (Cout 128).
//
// Here a CTA loads, ONCE per tile, the input rows it needs including the 1-pixel halo
// ("slab": (R+2) x (W+2) pixels x 64 channels, one tiled-TMA box whose out-of-bounds part is the
// convolution's zero padding) and the nine taps read it through nine shifted UMMA descriptors:
// for 128B-swizzled K-major operands tcgen05.mma derives the swizzle from the shared-memory address
// (tools/probe_shift.py: any row offset works with a plain descriptor), so tap (r,s) is simply the
// slab base + (r*(W+2)+s) rows.  Accumulator row i is the padded-row-major position i of the tile:
// positions with (i mod (W+2)) >= W or i >= R*(W+2) are halo columns / unused rows and are
// skipped by the epilogue (12.5 % of the MMA rows at W = 112/56/28 with R = 1/2/4).
// Activation ingest drops ~8x; weights stay resident in shared memory when they fit (Cin = 64).
//
// CTA pair (cluster 2, tcgen05 cta_group::2) as in gemm2_sm100.cuh: the leader issues the MMAs
// (M = 256 = two tiles), each CTA owns one tile's slab, half of the weight tile and 128 TMEM lanes.
#pragma once
#include "gemm2_sm100.cuh"

namespace frb {

struct SlabParams {
  int B, H, W, R;        // images, spatial size (H == W of the layer), rows per tile
  int N;                 // Cout
  int b_stages;          // weight ring depth; == 9 * CHUNKS => weights resident in shared memory
  int nbuf;              // slab-unit buffers in flight (a unit = one 64-channel chunk of one tile's slab)
  int slab_bytes;        // per buffer, multiple of 1024
  int box_bytes;         // bytes one slab TMA box delivers
  const float* bias;     // [bias_cases][N]
  int bias_cases;
  const float* prelu;    // [N] or nullptr
  const __nv_bfloat16* residual;  // identity shortcut (same shape as out) or nullptr
  __nv_bfloat16* out;    // [B][H][W][N]
  int debug;             // profiling experiments only (FRB_SLAB_DEBUG); 0 in production
  long long* trace;      // [grid][128] globaltimer stamps when non-null (profiling only)
  int* progress;         // inter-layer dataflow counters (ptx.cuh); nullptr = whole-grid dependency
  int wait_target;
  int sig_fence;         // experiments only: 0 drops the release fence (UNSAFE)
  int stage_out;         // 1: the epilogue stages a tile's output in shared memory and ONE TMA store writes it (tmOut)
  int res_tma;           // 1 (only with stage_out): RES_TMA instantiation - the shortcut tile arrives by TMA (tmRes) in the
                         // tile's staging buffer; the map describes the residual as [pixels][64 channels] like tmOut
};

constexpr int kSlabStageBuf = 16384;         // one staging buffer: 112 pixels x 64 channels (swizzled 128-byte rows)
__host__ __device__ __forceinline__ int slab_stage_bufs(int stage_out, int res_tma) { return stage_out ? (res_tma ? 3 : 2) : 0; }

__device__ __forceinline__ void tma2_load_4d(const CUtensorMap* m, uint32_t bar_cluster_addr, void* dst, int c0, int c1,
                                             int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "l"(kTmaMemDescDefault)
      : "memory");
}

// same with an explicit L2 eviction-priority descriptor (CUTLASS TMA::CacheHintSm90 values)
constexpr uint64_t kTmaEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kTmaEvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void tma2_load_4d_hint(const CUtensorMap* m, uint32_t bar_cluster_addr, void* dst, int c0, int c1,
                                                  int c2, int c3, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "l"(hint)
      : "memory");
}

__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define SLAB_TRACE(slot) do { if (p.trace) p.trace[blockIdx.x * 128 + (slot)] = gtimer(); } while (0)

// L2 prefetch of a tiled box (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

constexpr int kSlabMaxBStages = 18;
constexpr int kSlabMaxBuf = 6;

// K blocks are visited chunk-major: unit (tile, cc) runs its nine taps against weight K block tap*CHUNKS+cc,
// so a slab buffer holds ONE 64-channel chunk and is released after 36 MMAs.
// RES_TMA (its own instantiation, so the layers without a shortcut keep exactly the code they had): the shortcut tile
// comes in by TMA (tmRes) into the tile's staging buffer - see the epilogue.
template <int BLOCK_N, int CHUNKS, bool RES_TMA = false>
__global__ void __launch_bounds__(kGemm2Threads, 1)
conv_slab_sm100_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                       const SlabParams p) {
  constexpr int kBBytes = (BLOCK_N / 2) * kBlockK * 2;
  constexpr int kAcc = (512 / BLOCK_N) < 4 ? (512 / BLOCK_N) : 4;  // TMEM accumulator stages
  constexpr int kNumKb = 9 * CHUNKS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_slab = smem;                                  // [nbuf][slab_bytes]
  uint8_t* smem_b = smem + p.nbuf * p.slab_bytes;             // [b_stages][kBBytes]
  uint8_t* smem_stage = smem_b + p.b_stages * kBBytes;        // [2 or 3][16 KB] output staging (only with p.stage_out)
  const int n_stage = slab_stage_bufs(p.stage_out, RES_TMA ? 1 : 0);
  uint8_t* tail = smem_stage + n_stage * kSlabStageBuf;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint64_t* slab_full = bars;                             // [6]  leader only
  uint64_t* slab_empty = bars + kSlabMaxBuf;              // [6]  per CTA
  uint64_t* b_full = bars + 2 * kSlabMaxBuf;              // [18] leader only
  uint64_t* b_empty = b_full + kSlabMaxBStages;           // [18] per CTA
  uint64_t* tmem_full_bar = b_empty + kSlabMaxBStages;    // [4] per CTA
  uint64_t* tmem_empty_bar = tmem_full_bar + 4;           // [4] leader only, 16 arrivals
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 4);
  uint64_t* res_full = tmem_empty_bar + 5;                // [3] per CTA: the shortcut tile has landed in staging buffer i
  float* s_bias = reinterpret_cast<float*>(tail + 1024);      // [9][BLOCK_N]
  float* s_prelu = s_bias + 9 * (BLOCK_N + kBiasPad);

  // shuffle-broadcast makes the warp index provably warp-uniform for ptxas (uniform branches / registers)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int crank = static_cast<int>(cluster_ctarank());
  const bool leader = (crank == 0);
  if (threadIdx.x == 0) SLAB_TRACE(0);
  const int Wp = p.W + 2;
  const int tiles_per_img = p.H / p.R;
  const int num_tiles = p.B * tiles_per_img;
  const int total_pairs = (num_tiles + 1) / 2;
  const int first_pair = blockIdx.x >> 1;
  const int pair_step = gridDim.x >> 1;
  const bool resident = (p.b_stages == kNumKb);
  const int n_iters = (total_pairs - first_pair + pair_step - 1) / pair_step;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmB);
    if (p.stage_out) prefetch_tmap(&tmOut);
    for (int i = 0; i < kSlabMaxBuf; ++i) {
      mbar_init(&slab_full[i], 1);
      mbar_init(&slab_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 16);
    }
    for (int i = 0; i < kSlabMaxBStages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 3; ++i) mbar_init(&res_full[i], 1);
    if (RES_TMA) prefetch_tmap(&tmRes);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem2_alloc(tmem_ptr_smem, kAcc * BLOCK_N);
    tmem2_relinquish();
  }
  for (int i = threadIdx.x; i < p.bias_cases * BLOCK_N; i += kGemm2Threads) s_bias[(i / BLOCK_N) * (BLOCK_N + kBiasPad) + (i % BLOCK_N)] = p.bias[i];
  if (p.prelu != nullptr)
    for (int i = threadIdx.x; i < BLOCK_N; i += kGemm2Threads) s_prelu[i] = p.prelu[i];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  if (threadIdx.x == 0) SLAB_TRACE(1);
  pdl_launch_dependents();
  // Everything above (and the resident weight loads below) is independent of the previous layer; the producer
  // waits for it just before its first activation load, the epilogue warps before their first residual read / store.
  if (warp != 0 && (p.progress == nullptr || p.wait_target < 0)) pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; converged warp, elected issue) =====================
    // Slab units are issued as far ahead as buffers allow, interleaved with the weight K blocks, so the
    // slab depth is not throttled by the weight ring.
    const int n_units = n_iters * CHUNKS;
    const uint32_t slab_full_leader0 = mapa_u32(smem_u32(&slab_full[0]), 0);
    const uint32_t b_full_leader0 = mapa_u32(smem_u32(&b_full[0]), 0);
    int next = 0, nbuf_i = 0;        // next unit to issue and its buffer
    uint32_t nphase = 0;             // parity of that buffer's use count
    int u_it = 0, u_cc = 0;          // (iteration, chunk) of unit `next`
    auto issue_unit = [&]() {
      int tile = (first_pair + u_it * pair_step) * 2 + crank;
      if (tile >= num_tiles) tile = 0;  // padding tile of an odd count: stores are masked
      const int img = tile / tiles_per_img;
      const int h0 = (tile - img * tiles_per_img) * p.R;
      if (p.progress != nullptr && p.wait_target >= 0 && u_cc == 0) wait_images(p.progress, img, img, p.wait_target);
      if (leader && elect_one()) mbar_arrive_expect_tx(&slab_full[nbuf_i], 2 * ((p.debug & 32) ? 128 * p.W * (p.R + 2) : p.box_bytes));
      if (elect_one()) {
        if (p.debug & 16) {  // experiment: one TMA per input row instead of one (R+2)-row box
          for (int rr = 0; rr < p.R + 2; ++rr)
            tma2_load_4d(&tmX, slab_full_leader0 + 8 * nbuf_i, smem_slab + nbuf_i * p.slab_bytes + rr * Wp * 128, u_cc * kBlockK,
                         -1, h0 - 1 + rr, img);
        } else if (p.debug & 32) {  // experiment: fully in-bounds box (wrong numerics, timing only)
          int hh = h0 - 1;
          hh = hh < 0 ? 0 : (hh > p.H - p.R - 2 ? p.H - p.R - 2 : hh);
          tma2_load_4d(&tmX, slab_full_leader0 + 8 * nbuf_i, smem_slab + nbuf_i * p.slab_bytes, u_cc * kBlockK, 0, hh, img);
        } else if (p.debug & 512) {  // experiment: the input is dead after this layer (bar one residual read): evict it first
          tma2_load_4d_hint(&tmX, slab_full_leader0 + 8 * nbuf_i, smem_slab + nbuf_i * p.slab_bytes, u_cc * kBlockK, -1, h0 - 1, img,
                            kTmaEvictFirst);
        } else {
          tma2_load_4d(&tmX, slab_full_leader0 + 8 * nbuf_i, smem_slab + nbuf_i * p.slab_bytes, u_cc * kBlockK, -1, h0 - 1, img);
        }
      }
      if ((p.debug & 128) && elect_one()) {
        // pull the unit that will reuse this buffer (nbuf units ahead) into L2 now: its real load then hits L2
        const int f_unit = next + p.nbuf;
        const int f_it = f_unit / CHUNKS, f_cc = f_unit - f_it * CHUNKS;
        int f_tile = (first_pair + f_it * pair_step) * 2 + crank;
        if (f_it < n_iters && f_tile < num_tiles) {
          const int f_img = f_tile / tiles_per_img;
          const int f_h0 = (f_tile - f_img * tiles_per_img) * p.R;
          tma_prefetch_4d(&tmX, f_cc * kBlockK, -1, f_h0 - 1, f_img);
        }
      }
      __syncwarp();
      ++next;
      if (++u_cc == CHUNKS) { u_cc = 0; ++u_it; }
      if (++nbuf_i == p.nbuf) { nbuf_i = 0; nphase ^= 1; }
    };
    auto run_ahead = [&]() {
      while (next < n_units && mbar_test(&slab_empty[nbuf_i], nphase ^ 1)) issue_unit();
    };
    int bstage = 0;
    uint32_t bphase = 0;
    if (resident) {
      // weights do not depend on the previous layer: fetch them before waiting for it
      for (int cc = 0; cc < CHUNKS; ++cc)
        for (int tap = 0; tap < 9; ++tap) {
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(&b_full[bstage], 2 * kBBytes);
            tma2_load_2d(&tmB, b_full_leader0 + 8 * bstage, smem_b + bstage * kBBytes, (tap * CHUNKS + cc) * kBlockK,
                         crank * (BLOCK_N / 2) + ((p.debug & 64) ? (blockIdx.x >> 1) * BLOCK_N : 0));
          }
          __syncwarp();
          ++bstage;
        }
    }
    if (p.progress == nullptr || p.wait_target < 0) pdl_wait();
    for (int it = 0; it < n_iters; ++it) {
      while (next < (it + 1) * CHUNKS) {  // this tile's units must be in flight
        if (p.debug & 256) { while (!mbar_test(&slab_empty[nbuf_i], nphase ^ 1)) {} } else mbar_wait(&slab_empty[nbuf_i], nphase ^ 1);
        issue_unit();
      }
      run_ahead();
      if (!resident) {
#pragma unroll 1
        for (int cc = 0; cc < CHUNKS; ++cc) {
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_empty[bstage], bphase ^ 1);
            if (elect_one()) {
              if (leader) mbar_arrive_expect_tx(&b_full[bstage], 2 * kBBytes);
              tma2_load_2d(&tmB, b_full_leader0 + 8 * bstage, smem_b + bstage * kBBytes, (tap * CHUNKS + cc) * kBlockK,
                           crank * (BLOCK_N / 2));
            }
            __syncwarp();
            if (++bstage == p.b_stages) {
              bstage = 0;
              bphase ^= 1;
            }
            run_ahead();
          }
        }
      } else if (next < n_units) {
        // resident weights: nothing else to do but keep the slab ring full
        if (p.debug & 256) { while (!mbar_test(&slab_empty[nbuf_i], nphase ^ 1)) {} } else mbar_wait(&slab_empty[nbuf_i], nphase ^ 1);
        issue_unit();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // All waits run in the CONVERGED warp; the MMAs of a whole slab unit (resident weights: 36 back-to-back
    // UTCHMMAs) or of one K block (weight ring) are issued inside ONE elect.sync region with no wait in it.
    // ptxas only emits bare UTCHMMAs in such a branch-free single-thread region: any data-dependent branch
    // (an mbarrier spin) inside it, an `if (lane == 0)` region, or an elect per instruction makes it wrap
    // every UTCHMMA in an ELECT/BRA.U.ANY waterfall whose dependent predicate chain costs ~140 cycles per MMA
    // (profiles/r01f) - 2-4x the 32-64 cycles an N <= 128 MMA takes.
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kBlockM, BLOCK_N);
      const uint64_t desc0 = umma_desc_sw128(0);
      const uint32_t desc_hi = static_cast<uint32_t>(desc0 >> 32);
      const uint32_t slab_lo0 = ((smem_u32(smem_slab) & 0x3FFFFu) >> 4) | static_cast<uint32_t>(desc0);
      const uint32_t b_lo0 = ((smem_u32(smem_b) & 0x3FFFFu) >> 4) | static_cast<uint32_t>(desc0);
      const uint32_t slab_step = static_cast<uint32_t>(p.slab_bytes) >> 4;
      const uint32_t wp8 = static_cast<uint32_t>(Wp) * 8;  // one padded image row = Wp * 128 B
      int bstage = 0;
      uint32_t bphase = 0;
      int buf = 0;
      uint32_t sphase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (resident) {
        for (int i = 0; i < kNumKb; ++i) mbar_wait(&b_full[i], 0);  // weights land once
      }
      if (lane == 0) SLAB_TRACE(2);
      for (int it = 0; it < n_iters; ++it) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        if (lane == 0 && it < 28) SLAB_TRACE(16 + it * 4);
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
#pragma unroll
        for (int cc = 0; cc < CHUNKS; ++cc) {
          mbar_wait(&slab_full[buf], sphase);
          tc_fence_after();
          if (lane == 0 && it < 4) SLAB_TRACE(3 + it);
          if (lane == 0 && it < 28) SLAB_TRACE(16 + it * 4 + 1 + cc);
          const uint32_t a_lo = slab_lo0 + buf * slab_step;
          if (resident) {
            if (elect_one()) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const uint32_t b_lo = b_lo0 + (cc * 9 + tap) * (kBBytes >> 4);  // stages were filled in (cc, tap) order
                const uint32_t a_tap = a_lo + (tap / 3) * wp8 + (tap % 3) * 8;
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma2_bf16_ss_lo(tmem_d, a_tap + 2 * k, b_lo + 2 * k, desc_hi, idesc, (cc > 0 || tap > 0 || k > 0) ? 1u : 0u);
              }
              umma2_commit_pair(&slab_empty[buf]);
              if (cc == CHUNKS - 1) umma2_commit_pair(&tmem_full_bar[acc]);
            }
            __syncwarp();
          } else {
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&b_full[bstage], bphase);
              tc_fence_after();
              const uint32_t b_lo = b_lo0 + bstage * (kBBytes >> 4);
              const uint32_t a_tap = a_lo + (tap / 3) * wp8 + (tap % 3) * 8;
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma2_bf16_ss_lo(tmem_d, a_tap + 2 * k, b_lo + 2 * k, desc_hi, idesc, (cc > 0 || tap > 0 || k > 0) ? 1u : 0u);
                umma2_commit_pair(&b_empty[bstage]);
                if (tap == 8) {
                  umma2_commit_pair(&slab_empty[buf]);
                  if (cc == CHUNKS - 1) umma2_commit_pair(&tmem_full_bar[acc]);
                }
              }
              __syncwarp();
              if (++bstage == p.b_stages) {
                bstage = 0;
                bphase ^= 1;
              }
            }
          }
          if (++buf == p.nbuf) {
            buf = 0;
            sphase ^= 1;
          }
        }
        if (++acc == kAcc) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      if (lane == 0) SLAB_TRACE(7);
    }
  } else {
    // ===================== epilogue warps 2..9 =====================
    // Output path.  Direct: every thread stores its pixel's 64-byte pieces straight to global memory - a warp store
    // then touches 32 different 128-byte lines, and ncu showed the LSU data pipe at 78 % of its wavefront peak in the
    // 112x112 layer (tensor pipe 46 %, HBM 2.7 TB/s: profiles/r02_summary.md).  Staged (p.stage_out, the Cout = 64
    // layers): the tile's 112 x 64-channel block goes to a swizzled shared-memory buffer (conflict-free 16-byte
    // stores) and ONE TMA store per tile writes its 14 KB, contiguous in NHWC; two buffers, so the store of tile n
    // drains while tile n + 1 is computed.  Same values, same rounding: no bit changes.
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const bool staged = p.stage_out != 0;
    const int epi_tid = threadIdx.x - 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    int stage_buf = 0;
    // Round 2, second pass over this epilogue.  ncu source view of the 112x112 Cout = 64 layer: the MMA warp waits for a
    // free accumulator 14x longer than for a slab, the producer sits on a full ring - the layer is bound by the
    // per-tile latency chain of THIS code (all eight warps serve every tile), 1.17 us per tile against 0.86 us of
    // MMAs.  19 % of the epilogue samples were the per-tile index arithmetic (two integer divisions through
    // MUFU.RCP), 21 % waits on generic LD.E loads of the PReLU slopes from shared memory, 26 % the TMEM load together
    // with generic loads of the bias row.  Hence: everything that depends only on the thread is computed once, the
    // tile's position in its image is tracked incrementally (no division), the output offset is tile * R * W + t,
    // shared-memory tables are read with ld.shared, and for BLOCK_N = 64 (one 32-column chunk per warp) the slopes
    // live in registers for the whole kernel.
    const int i = quad * 32 + lane;             // accumulator row = padded position in the tile
    const int ri = i / Wp, wi = i - ri * Wp;
    const bool valid_pos = (ri < p.R) && (wi < p.W);
    const int t_pix = ri * p.W + wi;            // pixel of the tile (row of the staging buffer, offset in the output)
    const int cc = (wi == 0) ? 0 : ((wi == p.W - 1) ? 2 : 1);
    const int tile_step = 2 * pair_step;
    const int th_step = tile_step % tiles_per_img;
    int th = (first_pair * 2 + crank) % tiles_per_img;   // index of the tile inside its image
    float slope[(BLOCK_N == 64) ? 32 : 1];
    if (BLOCK_N == 64 && p.prelu != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) lds_f32x4(smem_u32(s_prelu + half * 32 + 4 * j), slope[4 * j], slope[4 * j + 1], slope[4 * j + 2], slope[4 * j + 3]);
    }
    // Shortcut by TMA (p.res_tma, the staged Cout = 64 layers): read thread by thread, a warp's 64-byte pieces of
    // 32 different pixels touched 32 lines per load instruction - the same pattern the direct stores had, and the
    // reason a conv2 (shortcut) layer took 101 us where its conv1 twin took 59.  Now the tile's 112 x 64 shortcut
    // block is fetched into the tile's own staging buffer two tiles ahead (three buffers), every thread adds its
    // pieces in place (same swizzled address for the read and the write) and the buffer goes out as before.
    constexpr bool res_tma = RES_TMA;       // host side: only together with stage_out
    auto issue_res = [&](int it_) {   // epi_tid 0 only; `it_` = index of the tile in this CTA's sequence
      const int tl = (first_pair + it_ * pair_step) * 2 + crank;
      if (first_pair + it_ * pair_step < total_pairs && tl < num_tiles) {
        const int b = it_ % 3;
        mbar_arrive_expect_tx(&res_full[b], 112 * 128);
        tma_load_2d(&tmRes, &res_full[b], smem_stage + b * kSlabStageBuf, 0, tl * (p.R * p.W));
      }
    };
    if (res_tma && epi_tid == 0) {
      issue_res(0);
      issue_res(1);
    }
    int it = 0;
    for (int pair = first_pair; pair < total_pairs; pair += pair_step, it += res_tma ? 1 : 0) {
      const int tile = pair * 2 + crank;
      const bool valid = (tile < num_tiles) && valid_pos;
      int bias_case = 0;
      const size_t m = static_cast<size_t>(tile) * (p.R * p.W) + t_pix;   // == (img * H + h) * W + wi: tiles cover whole rows, in order
      if (p.bias_cases == 9) {
        const int h = th * p.R + ri;
        const int rc = (h == 0) ? 0 : ((h == p.H - 1) ? 2 : 1);
        bias_case = valid ? rc * 3 + cc : 0;
      }
      th += th_step;
      if (th >= tiles_per_img) th -= tiles_per_img;
      const uint32_t bias_row = smem_u32(s_bias + bias_case * (BLOCK_N + kBiasPad));
      // The shortcut rows of this thread's pixel are fetched BEFORE the wait for the accumulator, so their latency
      // (HBM or L2, ~1 us) runs under the MMAs of the tile instead of after them.  Same-box A/B in the bench: embed
      // 5.61-5.64 -> 5.53-5.58 ms (+1.0-1.3 % faces/s); the per-launch ncu times of the residual layers did not move
      // (they carry 1.5x the bytes of their twins and sit at the same ~3 TB/s).  Only with whole-launch ordering (every
      // producer finished before griddepcontrol.wait returned); in dataflow mode the accumulator barrier is what
      // orders this thread behind the producers, and the loads stay behind it.
      constexpr int kChunksPerWarp = BLOCK_N / 64;
      uint4 rs[kChunksPerWarp][4];
      const bool has_res = p.residual != nullptr && valid && !res_tma;
      const bool res_early = has_res && (p.progress == nullptr || p.wait_target < 0);
      if (res_early) {
#pragma unroll
        for (int k = 0; k < kChunksPerWarp; ++k) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + m * p.N + (half + 2 * k) * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) rs[k][j] = __ldg(rp + j);
        }
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      if (res_tma && tile < num_tiles) mbar_wait(&res_full[stage_buf], static_cast<uint32_t>((it / 3) & 1));
      if (threadIdx.x == 64 && pair == first_pair) SLAB_TRACE(8);
      if (threadIdx.x == 64 && pair == first_pair + pair_step) SLAB_TRACE(9);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N;
#pragma unroll
      for (int k = 0; k < kChunksPerWarp; ++k) {
        const int c = half + 2 * k;
        if (p.debug & 8) break;
        uint32_t rr[32];
        tmem_ld_32x32(taddr + c * 32, rr);
        if (has_res && !res_early) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + m * p.N + c * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) rs[k][j] = __ldg(rp + j);
        }
        tmem_ld_wait();
        if (valid && !(p.debug & 4)) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 b;
            lds_f32x4(bias_row + (c * 32 + 4 * j) * 4, b.x, b.y, b.z, b.w);
            v[4 * j] = __uint_as_float(rr[4 * j]) + b.x;
            v[4 * j + 1] = __uint_as_float(rr[4 * j + 1]) + b.y;
            v[4 * j + 2] = __uint_as_float(rr[4 * j + 2]) + b.z;
            v[4 * j + 3] = __uint_as_float(rr[4 * j + 3]) + b.w;
          }
          if (p.prelu != nullptr) {
            if (BLOCK_N == 64) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * slope[(BLOCK_N == 64) ? j : 0];
            } else {
              const uint32_t s4 = smem_u32(s_prelu + c * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 sl;
                lds_f32x4(s4 + 16 * j, sl.x, sl.y, sl.z, sl.w);
                v[4 * j] = v[4 * j] > 0.f ? v[4 * j] : v[4 * j] * sl.x;
                v[4 * j + 1] = v[4 * j + 1] > 0.f ? v[4 * j + 1] : v[4 * j + 1] * sl.y;
                v[4 * j + 2] = v[4 * j + 2] > 0.f ? v[4 * j + 2] : v[4 * j + 2] * sl.z;
                v[4 * j + 3] = v[4 * j + 3] > 0.f ? v[4 * j + 3] : v[4 * j + 3] * sl.w;
              }
            }
          }
          if (res_tma) {
            const uint32_t rrow = smem_u32(smem_stage) + stage_buf * kSlabStageBuf + t_pix * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rs[k][j].x), "=r"(rs[k][j].y), "=r"(rs[k][j].z), "=r"(rs[k][j].w)
                           : "r"(rrow + (((c * 4 + j) ^ (t_pix & 7)) << 4)));
          }
          if (p.residual != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t w[4] = {rs[k][j].x, rs[k][j].y, rs[k][j].z, rs[k][j].w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                v[8 * j + 2 * t] += __uint_as_float(w[t] << 16);
                v[8 * j + 2 * t + 1] += __uint_as_float(w[t] & 0xFFFF0000u);
              }
            }
          }
          if (staged) {
            const int t = t_pix;                               // pixel of the tile, row of the staging buffer
            const uint32_t row = smem_u32(smem_stage) + stage_buf * kSlabStageBuf + t * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t ox = pack_bf16x2(v[8 * j], v[8 * j + 1]), oy = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              const uint32_t oz = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), ow = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((c * 4 + j) ^ (t & 7)) << 4)), "r"(ox), "r"(oy),
                           "r"(oz), "r"(ow) : "memory");
            }
          } else {
            uint4* dst = reinterpret_cast<uint4*>(p.out + m * p.N + c * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
              o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
              o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              dst[j] = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[acc]), 0));
      if (staged) {
        // the store of the PREVIOUS tile (other buffer) must have finished reading before the NEXT tile overwrites it:
        // the issuing thread waits for it here, everybody learns through the barrier
        if (epi_tid == 0) tma_store_wait_read<0>();
        fence_proxy_async();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (epi_tid == 0 && tile < num_tiles) {
          tma_store_2d(&tmOut, smem_stage + stage_buf * kSlabStageBuf, 0, tile * (p.R * p.W));
          tma_store_commit();
        }
        // the buffer of tile it + 2 is the one tile it - 1 was stored from: that store has finished reading (wait above)
        if (res_tma) {
          if (epi_tid == 0) issue_res(it + 2);
          stage_buf = (stage_buf == 2) ? 0 : stage_buf + 1;
        } else {
          stage_buf ^= 1;
        }
      }
      if (p.progress != nullptr) signal_rows(p.progress, valid, tile / tiles_per_img, BLOCK_N / 64, p.sig_fence != 0);
      if (++acc == kAcc) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  if (threadIdx.x == 64) SLAB_TRACE(10);
  if (p.stage_out && threadIdx.x == 64) tma_store_wait<0>();   // the staging buffers are read until the stores complete
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x == 0) SLAB_TRACE(11);
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem2_dealloc(tmem_base, kAcc * BLOCK_N);
  }
}

}  // namespace frb
