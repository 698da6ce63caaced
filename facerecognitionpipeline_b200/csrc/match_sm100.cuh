// match_sm100.cuh — gallery match: bf16 probe x gallery cosine GEMM on tcgen05 with a fused
// per-row running top-k' in the epilogue, so the P x N score matrix never reaches HBM.
//
// Replaces the arithmetic of GalleryManager.search (reference gallery_manager.py:189-205:
// normalise query, np.dot(G, q), argsort()[::-1][:k]) for P probes at once.
//
// Work item = (probe tile of 128 rows, gallery slice of consecutive 256-row tiles).
// The probe tile (128 x 512 bf16 = 128 KB) stays resident in shared memory; gallery tiles
// stream through a 3-stage TMA ring (32 KB per K block).  Each epilogue thread owns one probe
// row (= one TMEM lane) and keeps that row's best kCand approximate scores of the slice in
// registers.  A second kernel (match_finalize) merges the slices, re-scores the survivors
// exactly from the fp32 gallery in f64, applies the canonical tie-break (score desc, index
// asc) and proves, per row, that the bf16 filter cannot have dropped a true top-k entry
// (rows failing the proof are re-done by the exact scan).
#pragma once
#include "gemm2_sm100.cuh"

namespace frb {

constexpr int kMatchDim = 512;          // embedding size
constexpr int kMatchKB = kMatchDim / 64;  // 8 K blocks
constexpr int kMatchBN = 256;           // gallery rows per MMA tile
#ifndef FRB_KCAND
#define FRB_KCAND 8
#endif
constexpr int kCand = FRB_KCAND;        // candidates kept per (probe, slice); a multiple of 4
constexpr int kMatchBStages = 3;
constexpr int kMatchThreads = 192;

// bf16 gallery copy, K-blocked: rows are grouped by kGalGroup = 128; K block kb (64 columns) of group g is ONE
// contiguous 16 KB box at box-row (g * 8 + kb) * 128 of a [groups * 8 * 128][64] tensor, so every TMA stage of the
// filter streams a contiguous piece of HBM (a row-major copy made each box 128 pieces of 128 B, 1 KB apart).
constexpr int kGalGroup = 128;
__host__ __device__ __forceinline__ int gallery_box_row(int group, int kb) { return (group * kMatchKB + kb) * kGalGroup; }

// L2 prefetch of a 2-D tiled box (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
               : "memory");
}

struct MatchParams {
  int prefetch_tiles;    // gallery tiles pulled into L2 ahead of the TMA ring (0 = off)
  int P;                 // probes
  long long N;           // gallery rows held by this rank
  int p_tiles;           // ceil(P / 128)
  int g_tiles;           // ceil(N / 256)
  int slices;            // gallery slices
  int tiles_per_slice;   // ceil(g_tiles / slices)
  float* cand_score;     // [P][slices][kCand]
  int* cand_idx;         // [P][slices][kCand]  (local gallery row, -1 = empty)
  unsigned* row_floor;   // [P] shared admission floors (ordered-integer image of a score, 0 = none yet), zeroed per match
};

// Order-preserving image of a float in an unsigned integer (so atomicMax works on scores of either sign); 0 is below
// every real score.
__device__ __forceinline__ unsigned score_key(float v) {
  const unsigned b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_score(unsigned k) {
  if (k == 0u) return -INFINITY;
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct MatchSmem {
  static constexpr int kABytes = 128 * 64 * 2;       // one K block of the probe tile (16 KB)
  static constexpr int kBBytes = kMatchBN * 64 * 2;  // one K block of a gallery tile (32 KB)
  static constexpr int kTotal = kMatchKB * kABytes + kMatchBStages * kBBytes + 256 + 1024;
};

// Insert (v, idx) into the descending list; v = -inf is a no-op (every slot "keeps").  Equal scores stay in arrival
// order (the newcomer goes below them), so the lower column wins a tie.  Written as independent selects per slot - a
// bubble of compare-and-swaps is a dependent chain of kCand steps, and the epilogue warps (one per scheduler) are
// latency-bound, not issue-bound (profiles/r02_summary.md: issue active 25 % while the tensor pipe waited for them).
__device__ __forceinline__ void cand_insert(float (&ls)[kCand], int (&li)[kCand], float v, int idx) {
  bool keep[kCand];
#pragma unroll
  for (int i = 0; i < kCand; ++i) keep[i] = ls[i] >= v;
  float nl[kCand];
  int ni[kCand];
  nl[0] = keep[0] ? ls[0] : v;
  ni[0] = keep[0] ? li[0] : idx;
#pragma unroll
  for (int i = 1; i < kCand; ++i) {
    nl[i] = keep[i] ? ls[i] : (keep[i - 1] ? v : ls[i - 1]);
    ni[i] = keep[i] ? li[i] : (keep[i - 1] ? idx : li[i - 1]);
  }
#pragma unroll
  for (int i = 0; i < kCand; ++i) {
    ls[i] = nl[i];
    li[i] = ni[i];
  }
}

// Fold 32 consecutive scores into the row's running top-kCand.  The common case (nothing beats the current
// kCand-th score) costs one max-reduction.  Otherwise candidates are taken one at a time in (score desc, column asc)
// order — argmax, insert, knock out, repeat — so a warp in which a single lane has a single candidate pays ~150
// instructions, not 32 predicated insertions (~1300): with 32 independent rows per warp SOME lane has a
// candidate in most chunks until ~10^5 scores have been seen, and that path bounded the whole match
// (profiles/r01c: tensor pipe 37 % active at P = 4096).  Round 2: the four epilogue warps of a CTA run one per
// scheduler, so this path is bound by instruction LATENCY, not issue (profiles/r02d: issue active 25 %, tensor pipe
// 50 % on a 125 k-row shard): the running arg-max (a dependent chain of 32 compare-selects) became a tournament tree
// and the bubble insertion independent selects - filter 0.482 -> 0.353 ms at 4096 x 125 k, 2.575 -> 2.384 ms at
// 4096 x 1 M (1.76 PFLOP/s).
// `floor`: admission floor shared by all slices of the probe row (see MatchParams::row_floor): scores not above it are
// dropped without touching the list.  It never exceeds the smallest kept score of some FULL slice list, so the bound
// match_finalize_kernel derives from the full lists (max over slices of their last entry) already covers every
// element dropped this way - the proof there is unchanged.
__device__ __forceinline__ void cand_scan32(float (&ls)[kCand], int (&li)[kCand], float (&v)[32], int base, float floor) {
  float m = v[0];
#pragma unroll
  for (int j = 1; j < 32; ++j) m = fmaxf(m, v[j]);
  if (!__any_sync(0xffffffffu, m > fmaxf(ls[kCand - 1], floor))) return;
  // Slow path, warp-uniform: every round takes each lane's best remaining score (tournament over the 32 values, the
  // lower column wins a tie), inserts it if it clears the lane's gate and knocks it out; lanes without a candidate
  // ride along with no-op inserts.  The tournament is a tree (depth 5) instead of a running arg-max (depth 32).
#pragma unroll 1
  while (true) {
    float a[16];
    int ai[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const bool r = v[2 * j + 1] > v[2 * j];
      a[j] = r ? v[2 * j + 1] : v[2 * j];
      ai[j] = r ? 2 * j + 1 : 2 * j;
    }
#pragma unroll
    for (int w = 8; w >= 1; w >>= 1) {
#pragma unroll
      for (int j = 0; j < w; ++j) {
        const bool r = a[2 * j + 1] > a[2 * j];
        a[j] = r ? a[2 * j + 1] : a[2 * j];
        ai[j] = r ? ai[2 * j + 1] : ai[2 * j];
      }
    }
    const bool cand = a[0] > fmaxf(ls[kCand - 1], floor);
    if (!__any_sync(0xffffffffu, cand)) break;
    cand_insert(ls, li, cand ? a[0] : -INFINITY, base + ai[0]);
    const int bj = cand ? ai[0] : -1;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (j == bj) ? -INFINITY : v[j];
  }
}

// Slices of one probe row warm each other up: a thread whose list is full publishes its smallest kept score
// (atomicMax on the ordered-integer image) and, once per gallery tile, takes the largest value published so far as
// its floor.  With 37-74 slices per probe tile running at once the floors reach the level of the row's few-hundredth
// best score after a tile or two, instead of every slice refilling a cold list through the slow insertion path
// (the limiter named in profiles/r01c: tensor pipe 37 % active at P = 4096 while the lists were cold).
__device__ __forceinline__ float floor_exchange(unsigned* row_floor, int row, bool row_ok, const float (&ls)[kCand],
                                                const int (&li)[kCand], unsigned& published) {
  if (!row_ok) return -INFINITY;
  if (li[kCand - 1] >= 0) {   // full list: its last entry bounds everything this slice dropped or will drop
    const unsigned mine = score_key(ls[kCand - 1]);
    if (mine > published) {
      atomicMax(row_floor + row, mine);
      published = mine;
    }
  }
  return key_score(__ldcg(row_floor + row));
}

__global__ void __launch_bounds__(kMatchThreads, 1)
match_filter_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmG,
                    const MatchParams p) {
  using S = MatchSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kMatchKB * S::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + kMatchBStages * S::kBBytes);
  uint64_t* a_full = bars;                    // 1
  uint64_t* a_empty = bars + 1;               // 1
  uint64_t* b_full = bars + 2;                // 3
  uint64_t* b_empty = b_full + kMatchBStages; // 3
  uint64_t* t_full = b_empty + kMatchBStages; // 2
  uint64_t* t_empty = t_full + 2;             // 2
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(t_empty + 2);

  // shuffle-broadcast makes the warp index provably warp-uniform for ptxas (uniform branches / registers)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int total_items = p.p_tiles * p.slices;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmP);
    prefetch_tmap(&tmG);
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int i = 0; i < kMatchBStages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);

  // Producer and MMA warps run CONVERGED; every wait is outside the elect.sync regions that issue TMA / UTCHMMA, so
  // ptxas emits bare back-to-back instructions (see conv_slab_sm100.cuh for the measurement behind this).
  pdl_launch_dependents();
  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0, a_phase = 0;
    pdl_wait();   // the probes come from the kernel launched just before (programmatic stream serialization)
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int pt = item % p.p_tiles;
      const int gs = item / p.p_tiles;
      const int t_begin = gs * p.tiles_per_slice;
      const int t_end = min(p.g_tiles, t_begin + p.tiles_per_slice);
      // probe tile: wait until the previous item's MMAs no longer read it
      mbar_wait(a_empty, a_phase ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(a_full, kMatchKB * S::kABytes);
#pragma unroll
        for (int kb = 0; kb < kMatchKB; ++kb)
          tma_load_2d(&tmP, a_full, smem_a + kb * S::kABytes, kb * 64, pt * 128);
      }
      __syncwarp();
      a_phase ^= 1;
      for (int t = t_begin; t < t_end; ++t) {
#pragma unroll 1
        for (int kb = 0; kb < kMatchKB; ++kb) {
          mbar_wait(&b_empty[stage], phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&b_full[stage], S::kBBytes);
            // K-blocked gallery copy (gallery_prepare_kernel): a 256-row tile is two 128-row groups, each K block of a
            // group one contiguous 16 KB box
            tma_load_2d(&tmG, &b_full[stage], smem_b + stage * S::kBBytes, 0, gallery_box_row(2 * t, kb));
            tma_load_2d(&tmG, &b_full[stage], smem_b + stage * S::kBBytes + S::kBBytes / 2, 0, gallery_box_row(2 * t + 1, kb));
          }
          __syncwarp();
          if (++stage == kMatchBStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, kMatchBN);
    const uint64_t desc0 = umma_desc_sw128(0);
    const uint32_t desc_hi = static_cast<uint32_t>(desc0 >> 32);
    const uint32_t a_lo0 = ((smem_u32(smem_a) & 0x3FFFFu) >> 4) | static_cast<uint32_t>(desc0);
    const uint32_t b_lo0 = ((smem_u32(smem_b) & 0x3FFFFu) >> 4) | static_cast<uint32_t>(desc0);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0, a_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int gs = item / p.p_tiles;
      const int t_begin = gs * p.tiles_per_slice;
      const int t_end = min(p.g_tiles, t_begin + p.tiles_per_slice);
      mbar_wait(a_full, a_phase);
      a_phase ^= 1;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&t_empty[acc], acc_phase ^ 1);
        const uint32_t tmem_d = tmem_base + acc * kMatchBN;
#pragma unroll 1
        for (int kb = 0; kb < kMatchKB; ++kb) {
          mbar_wait(&b_full[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + kb * (S::kABytes >> 4);
          const uint32_t b_lo = b_lo0 + stage * (S::kBBytes >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ss_lo(tmem_d, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&b_empty[stage]);
            if (kb == kMatchKB - 1) {
              umma_commit(&t_full[acc]);
              if (t == t_end - 1) umma_commit(a_empty);  // all MMAs reading this probe tile have retired
            }
          }
          __syncwarp();
          if (++stage == kMatchBStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    const int quad = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int pt = item % p.p_tiles;
      const int gs = item / p.p_tiles;
      const int t_begin = gs * p.tiles_per_slice;
      const int t_end = min(p.g_tiles, t_begin + p.tiles_per_slice);
      const int row = pt * 128 + quad * 32 + lane;
      float ls[kCand];
      int li[kCand];
#pragma unroll
      for (int i = 0; i < kCand; ++i) {
        ls[i] = -INFINITY;
        li[i] = -1;
      }
      unsigned published = 0u;
      for (int t = t_begin; t < t_end; ++t) {
        const float floor = floor_exchange(p.row_floor, row, row < p.P, ls, li, published);
        mbar_wait(&t_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kMatchBN;
        const long long col0 = static_cast<long long>(t) * kMatchBN;
        const int ncols = static_cast<int>(min(static_cast<long long>(kMatchBN), p.N - col0));
#pragma unroll 1
        for (int c = 0; c < kMatchBN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (c * 32 >= ncols) continue;
          float v[32];
          if (ncols == kMatchBN) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          } else {  // last gallery tile: columns past N are TMA zero fill, not scores
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = (c * 32 + j < ncols) ? __uint_as_float(r[j]) : -INFINITY;
          }
          cand_scan32(ls, li, v, static_cast<int>(col0) + c * 32, floor);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      (void)floor_exchange(p.row_floor, row, row < p.P, ls, li, published);   // the final list helps the slices still running
      if (row < p.P) {
        const size_t o = (static_cast<size_t>(row) * p.slices + gs) * kCand;
        float4* ds = reinterpret_cast<float4*>(p.cand_score + o);
        int4* di = reinterpret_cast<int4*>(p.cand_idx + o);
#pragma unroll
        for (int q = 0; q < kCand / 4; ++q) {
          ds[q] = make_float4(ls[4 * q], ls[4 * q + 1], ls[4 * q + 2], ls[4 * q + 3]);
          di[q] = make_int4(li[4 * q], li[4 * q + 1], li[4 * q + 2], li[4 * q + 3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2), used when there are at least two probe tiles (P > 128).
// A pair works on TWO probe tiles (256 probes) against the same gallery slice: each CTA keeps its own 128-probe
// tile resident and streams only HALF of every gallery tile (128 of 256 rows per K block); the tensor cores
// exchange the halves.  That halves the L2->SM gallery traffic per probe — the bound of the 1-CTA kernel
// (64 B/clk/SM needed against the ~42 B/clk/SM ingest ceiling) — and doubles the ring depth (6 x 16 KB).
constexpr int kMatch2BStages = 6;
constexpr int kMatch2Threads = 192;

struct Match2Smem {
  static constexpr int kABytes = 128 * 64 * 2;              // one K block of this CTA's probe tile
  static constexpr int kBBytes = (kMatchBN / 2) * 64 * 2;   // this CTA's half of a gallery tile K block (16 KB)
  static constexpr int kTotal = kMatchKB * kABytes + kMatch2BStages * kBBytes + 256 + 1024;
};

// p.p_tiles here = number of probe-tile PAIRS; tmG2: gallery map with a 128-row box.
__global__ void __launch_bounds__(kMatch2Threads, 1)
match_filter2_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmG2,
                     const MatchParams p) {
  using S = Match2Smem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kMatchKB * S::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + kMatch2BStages * S::kBBytes);
  uint64_t* a_full = bars;                      // leader only
  uint64_t* a_empty = bars + 1;                 // per CTA (commit multicast)
  uint64_t* b_full = bars + 2;                  // [6] leader only
  uint64_t* b_empty = b_full + kMatch2BStages;  // [6] per CTA
  uint64_t* t_full = b_empty + kMatch2BStages;  // [2] per CTA
  uint64_t* t_empty = t_full + 2;               // [2] leader only, 8 arrivals
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int crank = static_cast<int>(cluster_ctarank());
  const bool leader = (crank == 0);
  const int total_items = p.p_tiles * p.slices;
  const int first_item = blockIdx.x >> 1, item_step = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmP);
    prefetch_tmap(&tmG2);
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int i = 0; i < kMatch2BStages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem2_alloc(tmem_ptr_smem, 512);
    tmem2_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  pdl_launch_dependents();

  if (warp == 0) {
    // ---- TMA producer (both CTAs): own probe tile + own half of every gallery tile, completing on the leader's barriers
    const uint32_t a_full_leader = mapa_u32(smem_u32(a_full), 0);
    const uint32_t b_full_leader0 = mapa_u32(smem_u32(&b_full[0]), 0);
    int stage = 0;
    uint32_t phase = 0, a_phase = 0;
    pdl_wait();   // the probes come from the kernel launched just before (programmatic stream serialization)
    for (int item = first_item; item < total_items; item += item_step) {
      const int pt = (item % p.p_tiles) * 2 + crank;
      const int gs = item / p.p_tiles;
      const int t_begin = gs * p.tiles_per_slice;
      const int t_end = min(p.g_tiles, t_begin + p.tiles_per_slice);
      mbar_wait(a_empty, a_phase ^ 1);
      if (elect_one()) {
        if (leader) mbar_arrive_expect_tx(a_full, 2 * kMatchKB * S::kABytes);
#pragma unroll
        for (int kb = 0; kb < kMatchKB; ++kb) tma2_load_2d(&tmP, a_full_leader, smem_a + kb * S::kABytes, kb * 64, pt * 128);
      }
      __syncwarp();
      a_phase ^= 1;
      for (int t = t_begin; t < t_end; ++t) {
#pragma unroll 1
        for (int kb = 0; kb < kMatchKB; ++kb) {
          mbar_wait(&b_empty[stage], phase ^ 1);
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(&b_full[stage], 2 * S::kBBytes);
            tma2_load_2d(&tmG2, b_full_leader0 + 8 * stage, smem_b + stage * S::kBBytes, 0, gallery_box_row(2 * t + crank, kb));
            // the ring holds 96 KB per SM; with few probes the filter is HBM-bound, so more bytes are kept in flight by
            // pulling the same K block of a later tile into L2 now
            if (p.prefetch_tiles > 0 && t + p.prefetch_tiles < t_end)
              tma_prefetch_2d(&tmG2, 0, gallery_box_row(2 * (t + p.prefetch_tiles) + crank, kb));
          }
          __syncwarp();
          if (++stage == kMatch2BStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (leader): 256 probes x 256 gallery rows per tile
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, kMatchBN);
      const uint64_t desc0 = umma_desc_sw128(0);
      const uint32_t desc_hi = static_cast<uint32_t>(desc0 >> 32);
      const uint32_t a_lo0 = ((smem_u32(smem_a) & 0x3FFFFu) >> 4) | static_cast<uint32_t>(desc0);
      const uint32_t b_lo0 = ((smem_u32(smem_b) & 0x3FFFFu) >> 4) | static_cast<uint32_t>(desc0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0, a_phase = 0;
      for (int item = first_item; item < total_items; item += item_step) {
        const int gs = item / p.p_tiles;
        const int t_begin = gs * p.tiles_per_slice;
        const int t_end = min(p.g_tiles, t_begin + p.tiles_per_slice);
        mbar_wait(a_full, a_phase);
        a_phase ^= 1;
        for (int t = t_begin; t < t_end; ++t) {
          mbar_wait(&t_empty[acc], acc_phase ^ 1);
          const uint32_t tmem_d = tmem_base + acc * kMatchBN;
#pragma unroll 1
          for (int kb = 0; kb < kMatchKB; ++kb) {
            mbar_wait(&b_full[stage], phase);
            tc_fence_after();
            const uint32_t a_lo = a_lo0 + kb * (S::kABytes >> 4);
            const uint32_t b_lo = b_lo0 + stage * (S::kBBytes >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma2_bf16_ss_lo(tmem_d, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              umma2_commit_pair(&b_empty[stage]);
              if (kb == kMatchKB - 1) {
                umma2_commit_pair(&t_full[acc]);
                if (t == t_end - 1) umma2_commit_pair(a_empty);
              }
            }
            __syncwarp();
            if (++stage == kMatch2BStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      }
    }
  } else {
    // ---- epilogue (both CTAs): one probe row per thread, running top-kCand of the slice in registers
    // (eight epilogue warps - two per scheduler, each pair splitting a tile's columns and keeping a list each - were
    // measured after the insertion path had been rewritten for instruction-level parallelism: filter 0.353 vs 0.359 ms
    // at 4096 x 125 k, but twice the lists cost the finalize kernel 0.025 ms; not kept)
    const int quad = warp & 3;
    const uint32_t t_empty_leader0 = mapa_u32(smem_u32(&t_empty[0]), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = first_item; item < total_items; item += item_step) {
      const int pt = (item % p.p_tiles) * 2 + crank;
      const int gs = item / p.p_tiles;
      const int t_begin = gs * p.tiles_per_slice;
      const int t_end = min(p.g_tiles, t_begin + p.tiles_per_slice);
      const int row = pt * 128 + quad * 32 + lane;
      float ls[kCand];
      int li[kCand];
#pragma unroll
      for (int i = 0; i < kCand; ++i) {
        ls[i] = -INFINITY;
        li[i] = -1;
      }
      unsigned published = 0u;
      for (int t = t_begin; t < t_end; ++t) {
        const float floor = floor_exchange(p.row_floor, row, row < p.P, ls, li, published);
        mbar_wait(&t_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kMatchBN;
        const long long col0 = static_cast<long long>(t) * kMatchBN;
        const int ncols = static_cast<int>(min(static_cast<long long>(kMatchBN), p.N - col0));
#pragma unroll 1
        for (int c = 0; c < kMatchBN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (c * 32 >= ncols) continue;
          float v[32];
          if (ncols == kMatchBN) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (c * 32 + j < ncols) ? __uint_as_float(r[j]) : -INFINITY;
          }
          cand_scan32(ls, li, v, static_cast<int>(col0) + c * 32, floor);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(t_empty_leader0 + 8 * acc);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      (void)floor_exchange(p.row_floor, row, row < p.P, ls, li, published);   // the final list helps the slices still running
      if (row < p.P) {
        const size_t o = (static_cast<size_t>(row) * p.slices + gs) * kCand;
        float4* ds = reinterpret_cast<float4*>(p.cand_score + o);
        int4* di = reinterpret_cast<int4*>(p.cand_idx + o);
#pragma unroll
        for (int q = 0; q < kCand / 4; ++q) {
          ds[q] = make_float4(ls[4 * q], ls[4 * q + 1], ls[4 * q + 2], ls[4 * q + 3]);
          di[q] = make_int4(li[4 * q], li[4 * q + 1], li[4 * q + 2], li[4 * q + 3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem2_dealloc(tmem_base, 512);
  }
}

}  // namespace frb
