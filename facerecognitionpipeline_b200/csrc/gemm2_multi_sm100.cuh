// gemm2_multi_sm100.cuh — a RUN of consecutive implicit-GEMM conv layers in ONE persistent launch.
//
// Why (profiles/r01c): a 14x14 layer of IR-101 at batch 256 is 2.65 rounds of ~13.5 us on 74 CTA pairs, but costs
// 45-54 us as a kernel of its own: ~9-14 us per layer go to the boundary — last epilogue, grid completion, CTA
// launch, prologue (barrier init, TMEM allocation, constants), first TMA round trip — and programmatic dependent
// launch hides little of it because a CTA owns a whole SM's shared memory.  Stage 3 + stage 4 of IR-101 are 66
// consecutive launches of gemm2_sm100_kernel<256>.  Here the CTAs stay resident across the whole run: per layer
// they read that layer's tensor maps and parameters from a descriptor array in global memory, and between layers
// they meet at a grid-wide barrier (one counter per layer, one release-add per CTA, the TMA producer warp polls).  The boundary shrinks
// to the slowest CTA's last epilogue + the barrier + one TMA round trip; TMEM, the mbarrier ring and its phases
// simply carry over.
//
// Same tile arithmetic and the same per-tile instruction streams as gemm2_sm100_kernel<BLOCK_N, 2> (results are
// bit-identical to launching the layers one by one).  Requires every CTA of the grid to be resident at once:
// grid <= number of SMs with one CTA per SM, which is how the launcher sizes it.
#pragma once
#include "gemm2_sm100.cuh"

namespace frb {

struct alignas(128) Gemm2Layer {  // one layer of a run (global memory; tensor maps need 64-byte alignment)
  CUtensorMap tmA, tmA2, tmB;
  GemmParams p;
};

// Tensor maps that live in GLOBAL memory (written by the host with cudaMemcpy) must be acquired through the tensormap
// proxy before the TMA unit uses them (CUDA programming guide, "Using TMA ... tensor map in global memory").
__device__ __forceinline__ void tensormap_acquire(const CUtensorMap* m) {
  asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kGemm2Threads, 1)
gemm2_multi_sm100_kernel(const Gemm2Layer* __restrict__ layers, int num_layers, int* __restrict__ grid_bar) {
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

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int crank = static_cast<int>(cluster_ctarank());
  const bool leader = (crank == 0);
  const int first_tile = blockIdx.x >> 1;
  const int tile_step = gridDim.x >> 1;
  const int num_ctas = gridDim.x;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < S::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
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
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  pdl_launch_dependents();
  pdl_wait();  // the run's first layer reads what the previous kernel wrote

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    int rot = 0;  // flow mode: tiles are dealt to the pairs round-robin ACROSS layers (layer l+1 continues where layer l
                  // stopped), so the pairs that took the partial last round of one layer are not the same in the next
    const uint32_t full_leader0 = mapa_u32(smem_u32(&full_bar[0]), 0);
    for (int l = 0; l < num_layers; ++l) {
      const Gemm2Layer* L = layers + l;
      const GemmParams p = L->p;  // by value: the loops below must not re-read global memory
      const int n_tiles = p.N / BLOCK_N;
      const int num_kb = p.num_kb_main + p.num_kb_sc;
      const int total_tiles = (((p.M + kBlockM - 1) / kBlockM + 1) / 2) * n_tiles;
      const int pq = p.P * p.Q;
      tensormap_acquire(&L->tmA);
      tensormap_acquire(&L->tmA2);
      tensormap_acquire(&L->tmB);
      if (lane == 0) {
        prefetch_tmap(&L->tmA);
        prefetch_tmap(&L->tmB);
        if (p.num_kb_sc > 0) prefetch_tmap(&L->tmA2);
      }
      const bool flow = p.progress != nullptr;   // per-image dependencies instead of the grid barrier (see below)
      if (l > 0 && !flow) {
        // grid barrier: every CTA has stored (and fenced) all its tiles of layer l-1
        // one counter PER LAYER: a CTA without tiles in some layers runs ahead and arrives early for them, which must
        // not count towards an earlier layer's barrier
        uint32_t spins = 0;
        while (ld_acquire_gpu(grid_bar + (l - 1)) < num_ctas) {
          if (++spins > (1u << 26)) {
            printf("frb: layer barrier timeout block %d layer %d\n", blockIdx.x, l);
            __trap();
          }
        }
        if (p.tail_debug & 1) __nanosleep(20000);   // diagnostics: is the barrier racing with something?
        fence_proxy_async_global();  // order the TMA reads below after the acquire
      }
      const int tile0 = flow ? (first_tile + tile_step - rot) % tile_step : first_tile;
      if (flow && !(p.tail_debug & 4)) rot = (rot + total_tiles) % tile_step;
      if (flow && l > 0 && p.full_wait && tile0 < total_tiles) {
        // The image-local argument below needs image i to occupy the same bytes of a buffer every time the buffer is
        // written.  Where the per-image size of the output buffer changes (stage transitions: 28x28x256 -> 14x14x256,
        // 14x14x256 -> 14x14x512 -> 7x7x512), image i of the new layout overlaps OTHER images of the old one, so this
        // layer may only start writing when every earlier layer has finished every image.
        wait_images(p.progress, 0, p.M / pq - 1, p.wait_target);
      }
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int n_tile = tile % n_tiles;
        const int m_tile = (tile / n_tiles) * 2 + crank;
        const int m0 = (m_tile * kBlockM < p.M) ? m_tile * kBlockM : 0;  // padding tile: re-read tile 0, stores masked
        const int img = m0 / pq;
        const int rem = m0 - img * pq;
        const int pp = rem / p.Q;
        const int qq = rem - pp * p.Q;
        const int w_main = qq * p.stride - p.pad, h_main = pp * p.stride - p.pad;
        const int w_sc = qq * p.sc_stride, h_sc = pp * p.sc_stride;
        const int b_row = n_tile * BLOCK_N + crank * (BLOCK_N / 2);
        if (flow && l > 0) {
          // Dataflow between the layers of a run: this CTA's 128 rows need only the images they lie in, completely
          // written by every earlier layer of the run (a 3x3 window never leaves its image).  Image-complete also means
          // every earlier reader of those images' rows is done, so the write-after-read side is covered too.  A pair
          // that runs out of tiles in layer l simply continues with layer l+1: no drain, no partial last round.
          const int m_last = min(m0 + kBlockM, p.M) - 1;
          wait_images(p.progress, img, m_last / pq, p.wait_target);
          if (p.tail_debug & 16) {
            __syncwarp();
            (void)ld_acquire_gpu(p.progress + img);
            fence_proxy_async_global();
            __nanosleep(2000);
          }
        }
        int tap_r = 0, tap_s = 0, cc = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          {
            const uint32_t full_leader = full_leader0 + 8 * stage;
            void* sa = smem_a + stage * S::kABytes;
            void* sb = smem_b + stage * S::kBBytes;
            if (elect_one()) {
              if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
              if (kb < p.num_kb_main)
                tma2_load_im2col_4d(&L->tmA, full_leader, sa, cc * kBlockK, w_main, h_main, img, static_cast<uint16_t>(tap_s),
                                    static_cast<uint16_t>(tap_r));
              else
                tma2_load_im2col_4d(&L->tmA2, full_leader, sa, cc * kBlockK, w_sc, h_sc, img, 0, 0);
              tma2_load_2d(&L->tmB, full_leader, sb, kb * kBlockK, b_row);
            }
          }
          __syncwarp();
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
      int rot = 0;
      for (int l = 0; l < num_layers; ++l) {
        const GemmParams p = layers[l].p;
        const int num_kb = p.num_kb_main + p.num_kb_sc;
        const int total_tiles = (((p.M + kBlockM - 1) / kBlockM + 1) / 2) * (p.N / BLOCK_N);
        const bool flow = p.progress != nullptr;
        const int tile0 = flow ? (first_tile + tile_step - rot) % tile_step : first_tile;
        if (flow && !(p.tail_debug & 4)) rot = (rot + total_tiles) % tile_step;
        for (int tile = tile0; tile < total_tiles; tile += tile_step) {
          mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            {
              const uint32_t a_lo = a_lo0 + stage * (S::kABytes >> 4);
              const uint32_t b_lo = b_lo0 + stage * (S::kBBytes >> 4);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma2_bf16_ss_lo(tmem_d, a_lo + 2 * k, b_lo + 2 * k, desc_hi, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                umma2_commit_pair(&empty_bar[stage]);
                if (kb == num_kb - 1) umma2_commit_pair(&tmem_full_bar[acc]);
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
    }
  } else {
    // ===================== epilogue warps 2..9 (both CTAs, own 128 rows) =====================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int epi_tid = threadIdx.x - 64;  // 0..255
    int acc = 0;
    uint32_t acc_phase = 0;
    int rot = 0;
    for (int l = 0; l < num_layers; ++l) {
      const GemmParams p = layers[l].p;
      const int n_tiles = p.N / BLOCK_N;
      const int total_tiles = (((p.M + kBlockM - 1) / kBlockM + 1) / 2) * n_tiles;
      const int tile0 = (p.progress != nullptr) ? (first_tile + tile_step - rot) % tile_step : first_tile;
      if (p.progress != nullptr && !(p.tail_debug & 4)) rot = (rot + total_tiles) % tile_step;
      // this layer's epilogue constants (the previous layer's readers are past the barrier at its end)
      if (p.progress != nullptr && l > 0) asm volatile("bar.sync 1, 256;" ::: "memory");  // flow mode has no end-of-layer barrier
      if (p.N == BLOCK_N) {
        for (int i = epi_tid; i < p.bias_cases * BLOCK_N; i += 256) s_bias[(i / BLOCK_N) * (BLOCK_N + kBiasPad) + (i % BLOCK_N)] = p.bias[i];
        if (p.prelu != nullptr)
          for (int i = epi_tid; i < BLOCK_N; i += 256) s_prelu[i] = p.prelu[i];
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int n_tile = tile % n_tiles;
        const int m_tile = (tile / n_tiles) * 2 + crank;
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
#pragma unroll 1
        for (int c = half; c < BLOCK_N / 32; c += 2) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          uint4 rs[4];
          if (p.residual != nullptr && valid && !(p.tail_debug & 2)) {
            // plain (coherent) loads: the residual was written by an earlier layer of THIS launch
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + res_off + n0 + c * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) rs[j] = __ldcg(rp + j);
          }
          tmem_ld_wait();
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
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[acc]), 0));
        if (p.progress != nullptr && l + 1 < num_layers) {
          if (p.tail_debug & 8) __threadfence();
          signal_rows(p.progress, valid, img, BLOCK_N / 64, true);
        }
        if (++acc == S::kAccStages) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
      // ---- end of layer: publish this CTA's stores, then one arrival at the grid barrier
      if (l + 1 < num_layers && p.progress == nullptr) {
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (epi_tid == 0) {
          __threadfence();
          atomicAdd(grid_bar + l, 1);
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
    tmem2_dealloc(tmem_base, S::kTmemCols);
  }
}

}  // namespace frb
