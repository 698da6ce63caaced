// gemm2_sm100.cuh — the implicit-GEMM convolution on CTA PAIRS (tcgen05 cta_group::2).
//
// Why: profiles/r01b show the 1-CTA kernel (gemm_sm100.cuh) bound by shared-memory bandwidth, not
// by L2 or the tensor pipe: per 128-cycle MMA (128x256x16) an SM reads 4 KB of A + 8 KB of B from
// shared memory while TMA writes the next stage into it.  A CTA pair computes a 256 x N tile with
// ONE instruction stream: each SM stages only its own 128 A rows and HALF of the weight tile
// (N/2 rows), the tensor cores exchange the B halves across the pair.  That halves B's L2->SM
// traffic, its shared-memory footprint (deeper pipeline: 6-9 stages instead of 4-8) and the
// shared-memory read pressure per MMA cycle.
//
// Cluster = (2,1,1).  CTA rank 0 ("leader") issues all tcgen05.mma; both CTAs run a TMA producer
// (own A rows + own B half, completing on the LEADER's full barrier) and four epilogue warps
// (own 128 accumulator rows, in their own TMEM).
#pragma once
#include "gemm_sm100.cuh"

namespace frb {

// ---- cluster-scope PTX used only here
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_smem_addr, uint32_t target_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_smem_addr), "r"(target_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-SM TMA forms: `bar_cluster_addr` is a shared::cluster address (the leader's barrier).
constexpr uint64_t kTmaMemDescDefault = 0x1000000000000000ull;
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* m, uint32_t bar_cluster_addr, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1),
      "l"(kTmaMemDescDefault)
      : "memory");
}
// multicast form: the box lands at the same offset in every CTA of `mask`; each destination's bytes complete on the
// barrier at the same offset in the EVEN CTA of that destination's pair (peer bit of `bar_addr` cleared - CUTLASS
// SM100_TMA_2SM_LOAD_MULTICAST passes its own barrier address & Sm100MmaPeerBitMask)
__device__ __forceinline__ void tma2_load_2d_mcast(const CUtensorMap* m, uint32_t bar_addr, void* dst, int c0, int c1,
                                                   uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
      " [%0], [%1, {%4, %5}], [%2], %3, %6;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "h"(mask), "r"(c0), "r"(c1),
      "l"(kTmaMemDescDefault)
      : "memory");
}
__device__ __forceinline__ void tma2_load_im2col_4d(const CUtensorMap* m, uint32_t bar_cluster_addr, void* dst, int c,
                                                    int w, int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8}, %9;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c), "r"(w), "r"(h),
      "r"(n), "h"(off_w), "h"(off_h), "l"(kTmaMemDescDefault)
      : "memory");
}
__device__ __forceinline__ void tmem2_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem2_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem2_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, descriptors given as (low word, shared high word): the low word carries the start address (>>4) and the
// LBO bit, so stepping along K or to another filter tap is one 32-bit add on the uniform datapath
__device__ __forceinline__ void umma2_bf16_ss_lo(uint32_t tmem_d, uint32_t adesc_lo, uint32_t bdesc_lo, uint32_t desc_hi,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "r"(adesc_lo), "r"(bdesc_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// four K steps (one 64-wide K block) in ONE asm statement, so ptxas moves the operands to uniform registers once
__device__ __forceinline__ void umma2_bf16_ss_lo_x4(uint32_t tmem_d, uint32_t adesc_lo, uint32_t bdesc_lo, uint32_t desc_hi,
                                                    uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 la, lb;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t"
      "add.u32 la, %1, 2;\n\t"
      "add.u32 lb, %2, 2;\n\t"
      "mov.b64 da, {la, %3};\n\t"
      "mov.b64 db, {lb, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, 1;\n\t"
      "add.u32 la, %1, 4;\n\t"
      "add.u32 lb, %2, 4;\n\t"
      "mov.b64 da, {la, %3};\n\t"
      "mov.b64 db, {lb, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, 1;\n\t"
      "add.u32 la, %1, 6;\n\t"
      "add.u32 lb, %2, 6;\n\t"
      "mov.b64 da, {la, %3};\n\t"
      "mov.b64 db, {lb, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, 1;\n\t"
      "}\n"
      ::"r"(tmem_d), "r"(adesc_lo), "r"(bdesc_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate_first)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_pair(uint64_t* bar) {  // arrive on `bar` in BOTH CTAs
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}

__device__ __forceinline__ void umma2_commit_mask(uint64_t* bar, uint16_t mask) {  // arrive on `bar` in every CTA of mask
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}

// Work items of one CTA pair (or quad) in launch order.  Full rounds first: tile = unit + i * units.  Then, when the
// last round is only partially filled (rem tiles for `units` clusters) and tail_split > 1, those tiles are cut along K:
// the rem * (split-1) partial producers come first, the rem owners last, so an owner only ever waits for CTAs that
// never wait themselves (no deadlock with every CTA resident).
struct TileItem {
  int tile;       // output tile (m unit * n_tiles + n tile)
  int kb0, kb1;   // K block range
  int part;       // -1 = whole K; 0 = owner of a split tile; >= 1 = partial producer
  int tail_tile;  // index among the split tiles
};
__device__ __forceinline__ bool tile_item(int it, int unit, int units, int total_tiles, int num_kb, int split, TileItem* w) {
  const int full_rounds = total_tiles / units;
  w->kb0 = 0; w->kb1 = num_kb; w->part = -1; w->tail_tile = 0;
  if (split <= 1 || it < full_rounds) {
    w->tile = unit + it * units;
    return w->tile < total_tiles;
  }
  const int rem = total_tiles - full_rounds * units;
  const int q = unit + (it - full_rounds) * units;
  const int nprod = rem * (split - 1);
  if (q >= rem * split) return false;
  if (q < nprod) {
    w->tail_tile = q / (split - 1);
    w->part = 1 + q % (split - 1);
  } else {
    w->tail_tile = q - nprod;
    w->part = 0;
  }
  w->tile = full_rounds * units + w->tail_tile;
  const int per = num_kb / split;
  w->kb0 = w->part * per;
  w->kb1 = w->kb0 + per;
  return true;
}

constexpr int kGemm2Threads = 320;  // producer warp + MMA warp + 8 epilogue warps
// The per-pixel bias table (9 border cases x Cout) is read with 16-byte loads by lanes that sit on different cases
// (a warp's 32 pixels span image rows and borders).  With a row pitch of Cout floats all cases fall on the same
// banks: ncu counted 3.4 extra wavefronts per load in the persistent run (34 M bank conflicts, the most sampled
// instruction of its epilogue).  Four floats of padding per row put neighbouring cases on neighbouring bank groups.
constexpr int kBiasPad = 4;

template <int BLOCK_N>
struct Gemm2Smem {
  static constexpr int kABytes = kBlockM * kBlockK * 2;            // own 128 A rows: 16 KB
  static constexpr int kBBytes = (BLOCK_N / 2) * kBlockK * 2;      // own half of the weight tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N == 256) ? 6 : (BLOCK_N == 128 ? 8 : 9);
  static constexpr int kAccStages = 2;
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kEpiBytes = 10 * BLOCK_N * 4 + 9 * kBiasPad * 4;  // bias table [<=9][BLOCK_N + pad] + PReLU [BLOCK_N], fp32
  static constexpr int kTotal = kStages * kStageBytes + 512 + kEpiBytes + 1024;
};

// CL = 2: one CTA pair per cluster.  CL = 4 ("quad"): two pairs that work on neighbouring 256-row tiles of the SAME
// weight tile in lock step; each CTA fetches a QUARTER of the weight tile and multicasts it to the CTA of the same
// pair rank in the other pair, which halves the weight L2->SM traffic again (the 14x14 layers are bound by the
// ~6300 B/clk the L2 slices deliver chip-wide: 231 MB of im2col operand + 231 MB of weights per layer at batch 256).
template <int BLOCK_N, int CL = 2>
__global__ void __launch_bounds__(kGemm2Threads, 1)
gemm2_sm100_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                   const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using S = Gemm2Smem<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + S::kStages * S::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStageBytes);
  uint64_t* full_bar = bars;                        // used in the leader only
  uint64_t* empty_bar = bars + S::kStages;          // per CTA, signalled by the leader's commit
  uint64_t* tmem_full_bar = bars + 2 * S::kStages;  // per CTA, signalled by the leader's commit
  uint64_t* tmem_empty_bar = tmem_full_bar + S::kAccStages;  // leader only: 8 epilogue warps arrive
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + S::kAccStages);
  float* s_bias = reinterpret_cast<float*>(smem + S::kStages * S::kStageBytes + 512);  // [cases][BLOCK_N]
  float* s_prelu = s_bias + 9 * (BLOCK_N + kBiasPad);                                               // [BLOCK_N]

  // shuffle-broadcast makes the warp index provably warp-uniform for ptxas (uniform branches / registers)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  static_assert(CL == 2 || CL == 4, "cluster of one or two CTA pairs");
  constexpr int PAIRS = CL / 2;
  const int crank_full = static_cast<int>(cluster_ctarank());
  const int crank = crank_full & 1;        // rank inside the CTA pair
  const int pair_id = crank_full >> 1;     // which pair of the cluster
  const int lead_rank = crank_full & ~1;   // cluster rank of this pair's leader
  const bool leader = (crank == 0);

  const int m_pairs = ((p.M + kBlockM - 1) / kBlockM + 1) / 2;  // 256-row super tiles
  const int m_units = (m_pairs + PAIRS - 1) / PAIRS;            // a cluster takes PAIRS neighbouring super tiles
  const int n_tiles = p.N / BLOCK_N;
  const int num_kb = p.num_kb_main + p.num_kb_sc;
  const int total_tiles = m_units * n_tiles;
  const int first_tile = blockIdx.x / CL;
  const int tile_step = gridDim.x / CL;
  const int split = (CL == 2) ? p.tail_split : 0;   // split-K of the last partial round (see TileItem)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (p.num_kb_sc > 0) prefetch_tmap(&tmA2);
    for (int i = 0; i < S::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], PAIRS);  // every pair's MMAs must have retired: peers multicast into this stage
    }
    for (int i = 0; i < S::kAccStages; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 16);  // 8 epilogue warps in each of the 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem2_alloc(tmem_ptr_smem, S::kTmemCols);
    tmem2_relinquish();
  }
  // Per-layer epilogue constants live in shared memory (they were a ~400-cycle global-load stall per
  // 32-column chunk).  With one N tile they are loaded once; Cout = 512 reloads per tile below.
  if (p.N == BLOCK_N) {
    for (int i = threadIdx.x; i < p.bias_cases * BLOCK_N; i += kGemm2Threads) s_bias[(i / BLOCK_N) * (BLOCK_N + kBiasPad) + (i % BLOCK_N)] = p.bias[i];
    if (p.prelu != nullptr)
      for (int i = threadIdx.x; i < BLOCK_N; i += kGemm2Threads) s_prelu[i] = p.prelu[i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  pdl_launch_dependents();
  // the prologue above overlapped the previous layer's tail.  Without progress counters the whole previous grid
  // must have finished; with them each tile waits only for the images it reads (producer warp, below).
  if (p.progress == nullptr || p.wait_target < 0) pdl_wait();

  // The producer and MMA warps run their loops CONVERGED (all 32 lanes) and only the instruction issue
  // itself is predicated on elect.sync: inside an `if (lane == 0)` region ptxas cannot prove operands
  // warp-uniform and wraps every TMA / UTCHMMA in an elect + R2UR.BROADCAST waterfall loop (~150 cycles of
  // single-thread issue per MMA, which bounded every Cout <= 128 layer; see profiles/r01c).
  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t full_leader0 = mapa_u32(smem_u32(&full_bar[0]), lead_rank);
    const uint32_t full_peerbit0 = smem_u32(&full_bar[0]) & 0xFEFFFFFFu;   // multicast form: "even CTA of each destination"
    const uint16_t b_mask = static_cast<uint16_t>((1u << crank) | (1u << (crank + 2)));
    const int pq = p.P * p.Q;
    TileItem w;
    for (int it = 0; tile_item(it, first_tile, tile_step, total_tiles, num_kb, split, &w); ++it) {
      const int tile = w.tile;
      const int n_tile = tile % n_tiles;
      const int m_tile = ((tile / n_tiles) * PAIRS + pair_id) * 2 + crank;
      const int m0 = (m_tile * kBlockM < p.M) ? m_tile * kBlockM : 0;  // padding tile: re-read tile 0, stores masked
      const int img = m0 / pq;
      const int rem = m0 - img * pq;
      const int pp = rem / p.Q;
      const int qq = rem - pp * p.Q;
      const int w_main = qq * p.stride - p.pad, h_main = pp * p.stride - p.pad;
      const int w_sc = qq * p.sc_stride, h_sc = pp * p.sc_stride;
      const int b_row = n_tile * BLOCK_N + crank * (BLOCK_N / 2) + pair_id * (BLOCK_N / CL);
      if (p.progress != nullptr && p.wait_target >= 0) {
        const int m_last = min(m0 + kBlockM, p.M) - 1;
        wait_images(p.progress, img, m_last / pq, p.wait_target);
      }
      // a split item starts in the middle of the filter (split tiles have no shortcut K blocks)
      int cc = w.kb0 % p.cin_chunks;
      int tap_r = (w.kb0 / p.cin_chunks) / 3, tap_s = (w.kb0 / p.cin_chunks) % 3;
      for (int kb = w.kb0; kb < w.kb1; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        {
          const uint32_t full_leader = full_leader0 + 8 * stage;
          void* sa = smem_a + stage * S::kABytes;
          void* sb = smem_b + stage * S::kBBytes;
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
            if (kb < p.num_kb_main)
              tma2_load_im2col_4d(&tmA, full_leader, sa, cc * kBlockK, w_main, h_main, img, static_cast<uint16_t>(tap_s),
                                  static_cast<uint16_t>(tap_r));
            else
              tma2_load_im2col_4d(&tmA2, full_leader, sa, cc * kBlockK, w_sc, h_sc, img, 0, 0);
            if (CL == 2)
              tma2_load_2d(&tmB, full_leader, sb, kb * kBlockK, b_row);
            else
              tma2_load_2d_mcast(&tmB, full_peerbit0 + 8 * stage, static_cast<uint8_t*>(sb) + pair_id * (S::kBBytes / 2),
                                 kb * kBlockK, b_row, b_mask);
          }
        }
        __syncwarp();
        // next K block: channel chunk fastest, then tap column, then tap row; then the shortcut chunks
        ++cc;
        if (kb + 1 == p.num_kb_main) {
          cc = 0;
        } else if (kb + 1 < p.num_kb_main && cc == p.cin_chunks) {
          cc = 0;
          if (++tap_s == 3) {
            tap_s = 0;
            ++tap_r;
          }
        }
        if (++stage == S::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kBlockM, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint64_t desc0 = umma_desc_sw128(0);
      const uint32_t desc_hi = static_cast<uint32_t>(desc0 >> 32);
      const uint32_t a_lo0 = ((smem_u32(smem_a) & 0x3FFFFu) >> 4) | static_cast<uint32_t>(desc0);
      const uint32_t b_lo0 = ((smem_u32(smem_b) & 0x3FFFFu) >> 4) | static_cast<uint32_t>(desc0);
      TileItem w;
      for (int it = 0; tile_item(it, first_tile, tile_step, total_tiles, num_kb, split, &w); ++it) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        const int kb_first = w.kb0, kb_last = w.kb1 - 1;
        for (int kb = kb_first; kb <= kb_last; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          {
            // descriptor low words are computed by the converged warp; ONE elect region issues the K block
            // (no branch inside it, so ptxas emits four bare UTCHMMAs + the commit; see conv_slab_sm100.cuh)
            const uint32_t a_lo = a_lo0 + stage * (S::kABytes >> 4);
            const uint32_t b_lo = b_lo0 + stage * (S::kBBytes >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k)
                umma2_bf16_ss_lo(tmem_d, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, (kb > kb_first || k > 0) ? 1u : 0u);
              umma2_commit_mask(&empty_bar[stage], static_cast<uint16_t>((1u << CL) - 1));   // one arrival in every CTA of the cluster
              if (kb == kb_last) umma2_commit_mask(&tmem_full_bar[acc], static_cast<uint16_t>(3u << lead_rank));   // accumulators ready in both CTAs of the pair
            }
          }
          __syncwarp();
          if (++stage == S::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++acc == S::kAccStages) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue warps 2..9 (both CTAs, own 128 rows) =====================
    // warp w reads TMEM lane quadrant (w & 3); the two warps of a quadrant take alternate 32-column chunks
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int epi_tid = threadIdx.x - 64;  // 0..255
    int acc = 0;
    uint32_t acc_phase = 0;
    TileItem w;
    for (int it = 0; tile_item(it, first_tile, tile_step, total_tiles, num_kb, split, &w); ++it) {
      const int tile = w.tile;
      const int n_tile = tile % n_tiles;
      const int m_tile = ((tile / n_tiles) * PAIRS + pair_id) * 2 + crank;
      const int row = quad * 32 + lane;
      const int m = m_tile * kBlockM + row;
      const bool valid = m < p.M;
      const int n0 = n_tile * BLOCK_N;
      int bias_case = 0, img = 0;
      size_t res_off = 0;
      if (valid) {
        const int pq = p.P * p.Q;
        img = m / pq;
        const int rem = m - img * pq;
        const int pp = rem / p.Q;
        const int qq = rem - pp * p.Q;
        if (p.bias_cases == 9) {
          const int rc = (pp == 0) ? 0 : ((pp == p.P - 1) ? 2 : 1);
          const int cc = (qq == 0) ? 0 : ((qq == p.Q - 1) ? 2 : 1);
          bias_case = rc * 3 + cc;
        }
        if (p.residual != nullptr)
          res_off = ((static_cast<size_t>(img) * p.RH + static_cast<size_t>(pp) * p.res_stride) * p.RW +
                     static_cast<size_t>(qq) * p.res_stride) * p.N;
      }
      if (p.N != BLOCK_N) {  // several N tiles (Cout = 512): refresh the constants of this tile
        asm volatile("bar.sync 1, 256;" ::: "memory");  // previous tile's readers are done
        for (int i = epi_tid; i < p.bias_cases * BLOCK_N; i += 256)
          s_bias[(i / BLOCK_N) * (BLOCK_N + kBiasPad) + (i % BLOCK_N)] = p.bias[(i / BLOCK_N) * p.N + n0 + (i % BLOCK_N)];
        if (p.prelu != nullptr)
          for (int i = epi_tid; i < BLOCK_N; i += 256) s_prelu[i] = p.prelu[n0 + i];
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      const float* bias_row = s_bias + bias_case * (BLOCK_N + kBiasPad);

      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N;
      // split tiles: this thread's row of the raw fp32 partial tiles, [tail tile][part-1][256 rows][BLOCK_N]
      float* part_row = nullptr;
      if (w.part >= 0)
        part_row = p.tail_partial + (static_cast<size_t>(w.tail_tile) * (split - 1) * 256 + crank * kBlockM + row) * BLOCK_N;
      if (w.part > 0) {
        // ---- partial producer: dump the accumulators, publish, next item
        float* dst_row = part_row + static_cast<size_t>(w.part - 1) * 256 * BLOCK_N;
#pragma unroll 1
        for (int c = half; c < BLOCK_N / 32; c += 2) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          uint4* dst = reinterpret_cast<uint4*>(dst_row + c * 32);
#pragma unroll
          if (!(p.tail_debug & 4))
            for (int j = 0; j < 8; ++j) __stcg(dst + j, make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]));
        }
        tc_fence_before();
        if (!(p.tail_debug & 8)) __threadfence();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[acc]), lead_rank));
          if (!(p.tail_debug & 8)) atomicAdd(p.tail_flag + w.tail_tile, 1);
        }
        if (++acc == S::kAccStages) {
          acc = 0;
          acc_phase ^= 1;
        }
        continue;
      }
      if (w.part == 0 && !(p.tail_debug & 1)) {
        // ---- owner: every partial of this tile must be in memory (8 warps x 2 CTAs per producer)
        const int target = (split - 1) * 16;
        uint32_t spins = 0;
        while (ld_acquire_gpu(p.tail_flag + w.tail_tile) < target) {
          __nanosleep(64);
          if (++spins > (1u << 24)) {
            printf("frb: split-K wait timeout block %d tile %d\n", blockIdx.x, tile);
            __trap();
          }
        }
      }
#pragma unroll 1
      for (int c = half; c < BLOCK_N / 32; c += 2) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        uint4 rs[4];
        if (p.residual != nullptr && valid) {  // residual loads fly while the TMEM load completes
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + res_off + n0 + c * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) rs[j] = __ldg(rp + j);
        }
        tmem_ld_wait();
        if (w.part == 0 && !(p.tail_debug & 2)) {
          for (int s2 = 0; s2 < split - 1; ++s2) {
            const uint4* src = reinterpret_cast<const uint4*>(part_row + static_cast<size_t>(s2) * 256 * BLOCK_N + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint4 q4 = __ldcg(src + j);
              r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + __uint_as_float(q4.x));
              r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + __uint_as_float(q4.y));
              r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + __uint_as_float(q4.z));
              r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + __uint_as_float(q4.w));
            }
          }
        }
        if (valid) {
          float v[32];
          const float4* bp = reinterpret_cast<const float4*>(bias_row + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = bp[j];
            v[4 * j] = __uint_as_float(r[4 * j]) + b.x;
            v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b.y;
            v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b.z;
            v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b.w;
          }
          if (p.prelu != nullptr) {
            const float4* s4 = reinterpret_cast<const float4*>(s_prelu + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 sl = s4[j];
              v[4 * j] = v[4 * j] > 0.f ? v[4 * j] : v[4 * j] * sl.x;
              v[4 * j + 1] = v[4 * j + 1] > 0.f ? v[4 * j + 1] : v[4 * j + 1] * sl.y;
              v[4 * j + 2] = v[4 * j + 2] > 0.f ? v[4 * j + 2] : v[4 * j + 2] * sl.z;
              v[4 * j + 3] = v[4 * j + 3] > 0.f ? v[4 * j + 3] : v[4 * j + 3] * sl.w;
            }
          }
          if (p.residual != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t w[4] = {rs[j].x, rs[j].y, rs[j].z, rs[j].w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                v[8 * j + 2 * t] += __uint_as_float(w[t] << 16);
                v[8 * j + 2 * t + 1] += __uint_as_float(w[t] & 0xFFFF0000u);
              }
            }
          }
          uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(m) * p.N + n0 + c * 32);
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[acc]), lead_rank));
      if (p.progress != nullptr) signal_rows(p.progress, valid, img, BLOCK_N / 64, p.sig_fence != 0);  // this warp stored BLOCK_N/64 chunks per row
      if (++acc == S::kAccStages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem2_dealloc(tmem_base, S::kTmemCols);
  }
}

}  // namespace frb
