// conv_slab_multi_sm100.cuh — a RUN of identical activation-slab conv layers in ONE persistent launch.
//
// Stage 2 of IR-101 is 24 consecutive launches of conv_slab_sm100_kernel<128, 2> (28x28, 128 -> 128, 24 % of a
// batch-256 step, tensor pipe 64 % active): every launch ends with a drain and a partial 13th round of tiles (896
// pair-tiles on 74 pairs).  As in gemm2_multi_sm100.cuh the CTA pairs stay resident for the whole run: a tile of
// layer l+1 starts when its image is complete in layer l (per-image progress counters), tiles are dealt round-robin
// across layers, TMEM / the slab ring / barrier phases carry over.  What is specific to this kernel: the weights are
// RESIDENT in shared memory (36-147 KB per CTA: no room for a second copy), so they are swapped at every layer
// boundary.  Round 1 released them with ONE commit after the layer's last MMA and then refilled all 18 K blocks with
// the tensor pipe idle (~4-5 us per boundary, every CTA at once: 0.35 ms slower per batch-256 embed than one launch
// per layer, profiles/r01d).  Now the swap is STREAMED through the pair's last tile of the layer: in that tile the
// MMA warp commits a per-K-block barrier (b_free[i]) right after the four MMAs that read weight block i for the last
// time, and the producer refills block i with the next layer's weights as soon as that barrier completes - while the
// remaining taps of the tile still run.  The first tile of the next layer waits per unit (9 K blocks) for its weights,
// which were requested one to two tile-times earlier.
//
// All layers of a run share geometry (B, H, W, R, Cin, Cout) and buffer layout, so image-complete dependencies cover
// read-after-write and write-after-read (see gemm2_multi_sm100.cuh).  Per-tile instruction streams are those of
// conv_slab_sm100_kernel's resident path: results are bit-identical to one launch per layer.
#pragma once
#include "conv_slab_sm100.cuh"
#include "gemm2_multi_sm100.cuh"

namespace frb {

struct alignas(128) SlabLayer {  // one layer of a run (global memory)
  CUtensorMap tmX, tmB;
  SlabParams p;
};

template <int BLOCK_N, int CHUNKS>
__global__ void __launch_bounds__(kGemm2Threads, 1)
conv_slab_multi_sm100_kernel(const SlabLayer* __restrict__ layers, int num_layers) {
  constexpr int kBBytes = (BLOCK_N / 2) * kBlockK * 2;
  constexpr int kAcc = (512 / BLOCK_N) < 4 ? (512 / BLOCK_N) : 4;  // TMEM accumulator stages
  constexpr int kNumKb = 9 * CHUNKS;
  const SlabParams g = layers[0].p;  // geometry and shared-memory carve-up are the same for every layer of the run
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_slab = smem;                                  // [nbuf][slab_bytes]
  uint8_t* smem_b = smem + g.nbuf * g.slab_bytes;             // [kNumKb][kBBytes] resident weights
  uint8_t* tail = smem_b + kNumKb * kBBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint64_t* slab_full = bars;                             // [6]  leader only
  uint64_t* slab_empty = bars + kSlabMaxBuf;              // [6]  per CTA
  uint64_t* b_full = bars + 2 * kSlabMaxBuf;              // [18] leader only
  uint64_t* b_free = b_full + kSlabMaxBStages;            // [18] per CTA: the layer's MMAs no longer read weight block i
  uint64_t* tmem_full_bar = b_full + 2 * kSlabMaxBStages; // [4] per CTA   (same offsets as conv_slab_sm100_kernel)
  uint64_t* tmem_empty_bar = tmem_full_bar + 4;           // [4] leader only, 16 arrivals
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 4);
  float* s_bias = reinterpret_cast<float*>(tail + 1024);      // [9][BLOCK_N]
  float* s_prelu = s_bias + 9 * (BLOCK_N + kBiasPad);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int crank = static_cast<int>(cluster_ctarank());
  const bool leader = (crank == 0);
  const int Wp = g.W + 2;
  const int tiles_per_img = g.H / g.R;
  const int num_tiles = g.B * tiles_per_img;
  const int total_pairs = (num_tiles + 1) / 2;
  const int my_pair = blockIdx.x >> 1;
  const int pair_step = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
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
      mbar_init(&b_free[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem2_alloc(tmem_ptr_smem, kAcc * BLOCK_N);
    tmem2_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  pdl_launch_dependents();
  if (warp != 0) pdl_wait();

  // tiles are dealt to the pairs round-robin ACROSS layers: layer l starts at pair `rot`, rot advances by the tile count
  auto first_pair_of = [&](int rot) { return (my_pair + pair_step - rot) % pair_step; };
  auto iters_of = [&](int first) { return first < total_pairs ? (total_pairs - first + pair_step - 1) / pair_step : 0; };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; converged warp, elected issue) =====================
    const uint32_t slab_full_leader0 = mapa_u32(smem_u32(&slab_full[0]), 0);
    const uint32_t b_full_leader0 = mapa_u32(smem_u32(&b_full[0]), 0);
    int nbuf_i = 0;
    uint32_t nphase = 0;
    int rot = 0;
    for (int l = 0; l < num_layers; ++l) {
      const SlabLayer* L = layers + l;
      const int* progress = L->p.progress;
      const int wait_target = L->p.wait_target;
      tensormap_acquire(&L->tmX);
      tensormap_acquire(&L->tmB);
      if (lane == 0) {
        prefetch_tmap(&L->tmX);
        prefetch_tmap(&L->tmB);
      }
      const int first = first_pair_of(rot);
      const int n_units = iters_of(first) * CHUNKS;
      rot = (rot + total_pairs) % pair_step;
      int next = 0, u_it = 0, u_cc = 0;
      auto issue_unit = [&]() {  // caller has made sure slab buffer nbuf_i is free
        int tile = (first + u_it * pair_step) * 2 + crank;
        if (tile >= num_tiles) tile = 0;  // padding tile of an odd count: stores are masked
        const int img = tile / tiles_per_img;
        const int h0 = (tile - img * tiles_per_img) * g.R;
        if (l > 0 && u_cc == 0) wait_images(progress, img, img, wait_target);   // the image is complete in layer l-1
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(&slab_full[nbuf_i], 2 * g.box_bytes);
          tma2_load_4d(&L->tmX, slab_full_leader0 + 8 * nbuf_i, smem_slab + nbuf_i * g.slab_bytes, u_cc * kBlockK, -1, h0 - 1, img);
        }
        __syncwarp();
        ++next;
        if (++u_cc == CHUNKS) { u_cc = 0; ++u_it; }
        if (++nbuf_i == g.nbuf) { nbuf_i = 0; nphase ^= 1; }
      };
      if (l == 0) {
        // the first layer's weights do not depend on the previous kernel: fetch them before waiting for it
        for (int i = 0; i < kNumKb; ++i) {
          const int cc = i / 9, tap = i - cc * 9;
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(&b_full[i], 2 * kBBytes);
            tma2_load_2d(&L->tmB, b_full_leader0 + 8 * i, smem_b + i * kBBytes, (tap * CHUNKS + cc) * kBlockK, crank * (BLOCK_N / 2));
          }
          __syncwarp();
        }
        pdl_wait();
      } else {
        // weight block i is refilled as soon as the previous layer's last tile has used it (b_free[i]); in between,
        // this layer's first slab units (independent of the weights) go out whenever a slab buffer is free
        for (int i = 0; i < kNumKb; ++i) {
          const int cc = i / 9, tap = i - cc * 9;
          while (next < n_units && mbar_test(&slab_empty[nbuf_i], nphase ^ 1)) issue_unit();
          mbar_wait(&b_free[i], static_cast<uint32_t>((l - 1) & 1));
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(&b_full[i], 2 * kBBytes);
            tma2_load_2d(&L->tmB, b_full_leader0 + 8 * i, smem_b + i * kBBytes, (tap * CHUNKS + cc) * kBlockK, crank * (BLOCK_N / 2));
          }
          __syncwarp();
        }
      }
      while (next < n_units) {
        mbar_wait(&slab_empty[nbuf_i], nphase ^ 1);
        issue_unit();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kBlockM, BLOCK_N);
      const uint64_t desc0 = umma_desc_sw128(0);
      const uint32_t desc_hi = static_cast<uint32_t>(desc0 >> 32);
      const uint32_t slab_lo0 = ((smem_u32(smem_slab) & 0x3FFFFu) >> 4) | static_cast<uint32_t>(desc0);
      const uint32_t b_lo0 = ((smem_u32(smem_b) & 0x3FFFFu) >> 4) | static_cast<uint32_t>(desc0);
      const uint32_t slab_step = static_cast<uint32_t>(g.slab_bytes) >> 4;
      const uint32_t wp8 = static_cast<uint32_t>(Wp) * 8;  // one padded image row = Wp * 128 B
      int buf = 0;
      uint32_t sphase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int rot = 0;
      for (int l = 0; l < num_layers; ++l) {
        const int n_iters = iters_of(first_pair_of(rot));
        rot = (rot + total_pairs) % pair_step;
        if (n_iters == 0) {
          // no tile of this layer for this pair (small batches): its weights still pass through shared memory, so
          // wait until they have landed, then hand every block back (a commit with nothing outstanding arrives at once)
          for (int i = 0; i < kNumKb; ++i) mbar_wait(&b_full[i], static_cast<uint32_t>(l & 1));
          if (elect_one()) {
#pragma unroll
            for (int i = 0; i < kNumKb; ++i) umma2_commit_pair(&b_free[i]);
          }
          __syncwarp();
        }
        for (int it = 0; it < n_iters; ++it) {
          const bool last = (it == n_iters - 1);   // the pair's last tile of this layer: releases the weights block by block
          mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
          const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
#pragma unroll
          for (int cc = 0; cc < CHUNKS; ++cc) {
            mbar_wait(&slab_full[buf], sphase);
            if (it == 0)   // first tile: this unit's nine weight blocks of the new layer
              for (int tap = 0; tap < 9; ++tap) mbar_wait(&b_full[cc * 9 + tap], static_cast<uint32_t>(l & 1));
            tc_fence_after();
            const uint32_t a_lo = slab_lo0 + buf * slab_step;
            if (!last) {
              if (elect_one()) {
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                  const uint32_t b_lo = b_lo0 + (cc * 9 + tap) * (kBBytes >> 4);
                  const uint32_t a_tap = a_lo + (tap / 3) * wp8 + (tap % 3) * 8;
#pragma unroll
                  for (int k = 0; k < kBlockK / 16; ++k)
                    umma2_bf16_ss_lo(tmem_d, a_tap + 2 * k, b_lo + 2 * k, desc_hi, idesc, (cc > 0 || tap > 0 || k > 0) ? 1u : 0u);
                }
                umma2_commit_pair(&slab_empty[buf]);
                if (cc == CHUNKS - 1) umma2_commit_pair(&tmem_full_bar[acc]);
              }
            } else {
              if (elect_one()) {
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                  const uint32_t b_lo = b_lo0 + (cc * 9 + tap) * (kBBytes >> 4);
                  const uint32_t a_tap = a_lo + (tap / 3) * wp8 + (tap % 3) * 8;
#pragma unroll
                  for (int k = 0; k < kBlockK / 16; ++k)
                    umma2_bf16_ss_lo(tmem_d, a_tap + 2 * k, b_lo + 2 * k, desc_hi, idesc, (cc > 0 || tap > 0 || k > 0) ? 1u : 0u);
                  umma2_commit_pair(&b_free[cc * 9 + tap]);   // weight block (cc, tap) has been read for the last time
                }
                umma2_commit_pair(&slab_empty[buf]);
                if (cc == CHUNKS - 1) umma2_commit_pair(&tmem_full_bar[acc]);
              }
            }
            __syncwarp();
            if (++buf == g.nbuf) {
              buf = 0;
              sphase ^= 1;
            }
          }
          if (++acc == kAcc) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps 2..9 =====================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int epi_tid = threadIdx.x - 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    int rot = 0;
    for (int l = 0; l < num_layers; ++l) {
      const SlabParams p = layers[l].p;
      const int first = first_pair_of(rot);
      rot = (rot + total_pairs) % pair_step;
      // this layer's epilogue constants (all eight warps are done with the previous layer's)
      if (l > 0) asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int i = epi_tid; i < p.bias_cases * BLOCK_N; i += 256) s_bias[(i / BLOCK_N) * (BLOCK_N + kBiasPad) + (i % BLOCK_N)] = p.bias[i];
      if (p.prelu != nullptr)
        for (int i = epi_tid; i < BLOCK_N; i += 256) s_prelu[i] = p.prelu[i];
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int pair = first; pair < total_pairs; pair += pair_step) {
        const int tile = pair * 2 + crank;
        const int i = quad * 32 + lane;           // accumulator row = padded position in the tile
        const int ri = i / Wp, wi = i - ri * Wp;
        const bool valid = (tile < num_tiles) && (ri < p.R) && (wi < p.W);
        int bias_case = 0, img = 0;
        size_t m = 0;
        if (valid) {
          img = tile / tiles_per_img;
          const int h = (tile - img * tiles_per_img) * p.R + ri;
          m = (static_cast<size_t>(img) * p.H + h) * p.W + wi;
          if (p.bias_cases == 9) {
            const int rc = (h == 0) ? 0 : ((h == p.H - 1) ? 2 : 1);
            const int cc = (wi == 0) ? 0 : ((wi == p.W - 1) ? 2 : 1);
            bias_case = rc * 3 + cc;
          }
        }
        const float* bias_row = s_bias + bias_case * (BLOCK_N + kBiasPad);
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
        for (int c = half; c < BLOCK_N / 32; c += 2) {
          uint32_t rr[32];
          tmem_ld_32x32(taddr + c * 32, rr);
          uint4 rs[4];
          if (p.residual != nullptr && valid) {
            // coherent loads: the residual was written by an earlier layer of THIS launch
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + m * p.N + c * 32);
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
              v[4 * j] = __uint_as_float(rr[4 * j]) + b.x;
              v[4 * j + 1] = __uint_as_float(rr[4 * j + 1]) + b.y;
              v[4 * j + 2] = __uint_as_float(rr[4 * j + 2]) + b.z;
              v[4 * j + 3] = __uint_as_float(rr[4 * j + 3]) + b.w;
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
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[acc]), 0));
        if (l + 1 < num_layers) signal_rows(p.progress, valid, img, BLOCK_N / 64, true);
        if (++acc == kAcc) {
          acc = 0;
          acc_phase ^= 1;
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
    tmem2_dealloc(tmem_base, kAcc * BLOCK_N);
  }
}

}  // namespace frb
