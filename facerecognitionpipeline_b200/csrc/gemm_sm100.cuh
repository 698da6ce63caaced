// gemm_sm100.cuh — persistent warp-specialised tcgen05 GEMM core for sm_100a.
//
// One kernel template serves the three dense contractions of the embed path:
//   * IR residual-unit convolutions as implicit GEMM (A operand gathered by im2col-mode TMA
//     straight from the NHWC activation tensor; zero padding = TMA out-of-bounds fill),
//     with the unit's 1x1 strided shortcut convolution folded in as extra K blocks,
//   * the BN-Flatten-FC tail as a split-K tiled GEMM writing fp32 partials.
//
// Roles (192 threads): warp 0 lane 0 = TMA producer, warp 1 lane 0 = tcgen05.mma issuer
// (warp 1 also owns the TMEM allocation), warps 2..5 = epilogue (TMEM -> registers ->
// bias / PReLU / residual -> bf16 NHWC in HBM).  Accumulators are double-buffered in TMEM so
// the epilogue of tile i overlaps the mainloop of tile i+1.
#pragma once
#include "ptx.cuh"

namespace frb {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int kGemmThreads = 192;

enum AMode : int { A_TILED = 0, A_IM2COL = 1 };

struct GemmParams {
  int M;            // GEMM rows: output pixels (B*P*Q) or batch rows
  int N;            // GEMM cols: Cout
  int num_kb_main;  // K blocks taken from the main A source
  int num_kb_sc;    // K blocks taken from the shortcut A source (0 = none)
  int num_splits;   // split-K factor (tiled mode / FC only; 1 otherwise)
  // im2col geometry of the main source
  int P, Q;         // output spatial size
  int stride;       // conv stride (also the traversal stride baked into the tensor map)
  int pad;          // 1 for 3x3, 0 for 1x1
  int cin_chunks;   // Cin / 64
  int sc_stride;    // stride of the fused 1x1 shortcut conv
  int sc_chunks;    // Cin_sc / 64
  // epilogue
  const float* bias;              // [bias_cases][N]
  int bias_cases;                 // 1, or 9 = border-position table (pre-activation BN fold)
  const float* prelu;             // [N] or nullptr
  const __nv_bfloat16* residual;  // identity shortcut source or nullptr; NHWC (., RH, RW, N)
  int res_stride, RH, RW;
  __nv_bfloat16* out;             // [M, N] bf16 (nullptr when out_f32 is used)
  float* out_f32;                 // [num_splits][M][N] fp32 raw partials (FC tail)
  // inter-layer dataflow (ptx.cuh: wait_images / signal_rows); nullptr = whole-grid dependency (pdl_wait)
  int* progress;
  int wait_target;
  int sig_fence;  // experiments only: 0 drops the release fence before the progress update (UNSAFE)
  // split-K for the LAST, partially filled round of tiles (gemm2_sm100_kernel, CL = 2): the `rem` tiles left after the
  // full rounds are cut into tail_split K ranges; part 0 ("owner") adds the other parts' raw fp32 accumulators
  // (tail_partial, [rem][tail_split-1][256][BLOCK_N]) before its epilogue, once tail_flag[tile] shows they are stored
  int tail_split;        // 0/1 = off
  float* tail_partial;
  int* tail_flag;        // [rem], zeroed before the launch
  int tail_debug;        // timing experiments only (wrong results): 1 no wait, 2 no partial loads, 4 no dump, 8 no fence/flag
  int full_wait;  // persistent runs in flow mode: this layer's output buffer changes its per-image layout (stage
                  // transition), so image-local dependencies do not cover the write-after-read side: wait for ALL images
  int quad;       // host side only: launch gemm2_sm100_kernel<., 4> (two CTA pairs sharing the weight tile by multicast)
};

template <int BLOCK_N>
struct GemmSmem {
  static constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KB
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;   // 8/16/32 KB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int kAccStages = 2;
  static constexpr int kTmemCols = 2 * BLOCK_N;  // 128 / 256 / 512: all powers of two >= 32
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kStages * kStageBytes + kBarBytes + 1024;  // +1024 alignment slack
};

// CLUSTER > 1: the CTAs of a cluster work on CLUSTER consecutive M tiles of the same N tile and
// share the weight tile: each CTA fetches 1/CLUSTER of it and TMA-multicasts it to all of them,
// which divides the L2->SM weight traffic (the measured bound of these convs) by CLUSTER.
template <int BLOCK_N, int A_MODE, int CLUSTER>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                  const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using S = GemmSmem<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + S::kStages * S::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + S::kStages;
  uint64_t* tmem_full_bar = bars + 2 * S::kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + S::kAccStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + S::kAccStages);

  // shuffle-broadcast makes the warp index provably warp-uniform for ptxas (uniform branches / registers)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  // work units ("tiles") are enumerated per cluster; CTA `crank` of a cluster takes M tile
  // m_super * CLUSTER + crank (tiles past the end are all-zero loads with masked stores)
  const int m_tiles = ((p.M + kBlockM - 1) / kBlockM + CLUSTER - 1) / CLUSTER;  // M super-tiles
  const int n_tiles = p.N / BLOCK_N;
  const int num_kb = p.num_kb_main + p.num_kb_sc;
  const int kb_per_split = (num_kb + p.num_splits - 1) / p.num_splits;
  const int total_tiles = m_tiles * n_tiles * p.num_splits;
  const int crank = (CLUSTER > 1) ? static_cast<int>(cluster_ctarank()) : 0;
  const int first_tile = blockIdx.x / CLUSTER;
  const int tile_step = gridDim.x / CLUSTER;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CLUSTER) - 1);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (p.num_kb_sc > 0) prefetch_tmap(&tmA2);
    for (int i = 0; i < S::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], CLUSTER);  // one tcgen05.commit arrive from every CTA of the cluster
    }
    for (int i = 0; i < S::kAccStages; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, S::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();  // peers' barriers must be initialised before any multicast
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int split = tile % p.num_splits;
        const int mn = tile / p.num_splits;
        const int n_tile = mn % n_tiles;
        const int m_tile = (mn / n_tiles) * CLUSTER + crank;
        // a padding tile past the end (odd tile count in a cluster) re-reads tile 0; its stores are masked
        const int m0 = (m_tile * kBlockM < p.M) ? m_tile * kBlockM : 0;
        int img = 0, pp = 0, qq = 0;
        if (A_MODE == A_IM2COL) {
          const int pq = p.P * p.Q;
          img = m0 / pq;
          const int rem = m0 - img * pq;
          pp = rem / p.Q;
          qq = rem - pp * p.Q;
        }
        const int kb_begin = split * kb_per_split;
        const int kb_end = min(num_kb, kb_begin + kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
          void* sa = smem_a + stage * S::kABytes;
          void* sb = smem_b + stage * S::kBBytes;
          if (A_MODE == A_IM2COL) {
            if (kb < p.num_kb_main) {
              const int tap = kb / p.cin_chunks;
              const int cc = kb - tap * p.cin_chunks;
              const int r = (p.pad ? tap / 3 : 0), s = (p.pad ? tap - 3 * (tap / 3) : 0);
              tma_load_im2col_4d(&tmA, &full_bar[stage], sa, cc * kBlockK, qq * p.stride - p.pad,
                                 pp * p.stride - p.pad, img, static_cast<uint16_t>(s),
                                 static_cast<uint16_t>(r));
            } else {
              const int cc = kb - p.num_kb_main;
              tma_load_im2col_4d(&tmA2, &full_bar[stage], sa, cc * kBlockK, qq * p.sc_stride,
                                 pp * p.sc_stride, img, 0, 0);
            }
          } else {
            tma_load_2d(&tmA, &full_bar[stage], sa, kb * kBlockK, m0);
          }
          if (CLUSTER > 1) {
            constexpr int kRows = BLOCK_N / CLUSTER;
            tma_load_2d_mcast(&tmB, &full_bar[stage], static_cast<uint8_t*>(sb) + crank * kRows * 128,
                              kb * kBlockK, n_tile * BLOCK_N + crank * kRows, kMask);
          } else {
            tma_load_2d(&tmB, &full_bar[stage], sb, kb * kBlockK, n_tile * BLOCK_N);
          }
          if (++stage == S::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int split = tile % p.num_splits;
        const int kb_begin = split * kb_per_split;
        const int kb_end = min(num_kb, kb_begin + kb_per_split);
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + stage * S::kABytes));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + stage * S::kBBytes));
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // +32 bytes per UMMA_K=16 step inside the 128-byte swizzle row (>>4 => +2)
            umma_bf16_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          // smem slot reusable once these MMAs retire (in every CTA the slot is multicast into)
          if (CLUSTER > 1) umma_commit_mcast(&empty_bar[stage], kMask);
          else umma_commit(&empty_bar[stage]);
          if (++stage == S::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == S::kAccStages) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue warps 2..5 =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
      const int split = tile % p.num_splits;
      const int mn = tile / p.num_splits;
      const int n_tile = mn % n_tiles;
      const int m_tile = (mn / n_tiles) * CLUSTER + crank;
      const int row = quad * 32 + lane;
      const int m = m_tile * kBlockM + row;
      const bool valid = m < p.M;
      const int n0 = n_tile * BLOCK_N;

      // per-row epilogue coordinates
      int bias_case = 0;
      size_t res_off = 0;
      if (A_MODE == A_IM2COL && valid) {
        const int pq = p.P * p.Q;
        const int img = m / pq;
        const int rem = m - img * pq;
        const int pp = rem / p.Q;
        const int qq = rem - pp * p.Q;
        if (p.bias_cases == 9) {
          const int rc = (pp == 0) ? 0 : ((pp == p.P - 1) ? 2 : 1);
          const int cc = (qq == 0) ? 0 : ((qq == p.Q - 1) ? 2 : 1);
          bias_case = rc * 3 + cc;
        }
        if (p.residual != nullptr) {
          res_off = ((static_cast<size_t>(img) * p.RH + static_cast<size_t>(pp) * p.res_stride) * p.RW +
                     static_cast<size_t>(qq) * p.res_stride) * p.N;
        }
      }
      const float* bias_row = p.bias ? p.bias + static_cast<size_t>(bias_case) * p.N + n0 : nullptr;

      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        if (valid) {
          if (p.out_f32 != nullptr) {
            float4* dst = reinterpret_cast<float4*>(
                p.out_f32 + (static_cast<size_t>(split) * p.M + m) * p.N + n0 + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                   __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          } else {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (bias_row != nullptr) {
              const float4* b4 = reinterpret_cast<const float4*>(bias_row + c * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(b4 + j);
                v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
              }
            }
            if (p.prelu != nullptr) {
              const float4* s4 = reinterpret_cast<const float4*>(p.prelu + n0 + c * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 s = __ldg(s4 + j);
                v[4 * j] = v[4 * j] > 0.f ? v[4 * j] : v[4 * j] * s.x;
                v[4 * j + 1] = v[4 * j + 1] > 0.f ? v[4 * j + 1] : v[4 * j + 1] * s.y;
                v[4 * j + 2] = v[4 * j + 2] > 0.f ? v[4 * j + 2] : v[4 * j + 2] * s.z;
                v[4 * j + 3] = v[4 * j + 3] > 0.f ? v[4 * j + 3] : v[4 * j + 3] * s.w;
              }
            }
            if (p.residual != nullptr) {
              const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + res_off + n0 + c * 32);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 x = __ldg(r4 + j);
                const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  // bf16 -> fp32 is a 16-bit shift
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
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == S::kAccStages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();  // no CTA may exit while a peer can still signal its barriers
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, S::kTmemCols);
  }
}

}  // namespace frb
