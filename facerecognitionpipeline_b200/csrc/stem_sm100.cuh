// stem_sm100.cuh — input layer Conv3x3(3->64, stride 1, pad 1) + BN + PReLU on the tensor cores.
//
// Reference: the first module of the external AdaFace `net.Backbone.input_layer` / iresnet `conv1-bn1-prelu`
// invoked from face_embedder.py:119,157.  K = 27 is far too thin for an im2col TMA (the
// 3-channel pixel is 6 bytes), so the A operand is gathered by ordinary threads: one CTA owns a strip of
// image rows; per row, thread x packs pixel x's 27 inputs (+5 zeros) into a 64-byte K-major row of a
// 128B-swizzled shared-memory tile, one elected thread issues two 128x64x16 tcgen05.mma, and the same
// threads run the epilogue (TMEM -> +bias -> PReLU -> bf16 -> swizzled smem tile -> one TMA store of
// the whole 112-pixel x 64-channel row, 14 KB contiguous in NHWC).  128 B go out per 6 B that come in, but the kernel
// is NOT bound by that write stream (3.3 TB/s where a write-only stream reaches 6.2 on this part, tools/hbm_rw_peaks.py):
// it is a chain of dependent steps per image row (gather - barrier - MMA - wait - epilogue - barrier - store), so what
// bounds it is how many rows an SM has in flight and how many shared-memory wavefronts each costs.
//
// Round 2 history (all bit-identical; profiles/r02_summary.md): two restructurings that did NOT help - two image rows
// per MMA step (block-diagonal [128][64] weights, N = 128, 3 CTAs per SM): 175 us against 155 us; two threads per pixel
// (256 threads, K halves in the gather, channel halves in the epilogue): 153 us.  What did: the gather's 27 two-byte
// GENERIC loads per pixel and the epilogue's generic loads of bias / slopes replaced by aligned ld.shared (75 % -> 62 %
// of the shared-memory wavefront peak, 155 -> 118 us), and the output tile reusing the A tile (50.6 -> 34 KB of shared
// memory, 4 -> 6 CTAs per SM, 118 -> 105 us; without the alignment slack 7 CTAs per SM, 100 us).
#pragma once
#include "ptx.cuh"

namespace frb {

constexpr int kStemRows = 8;          // image rows per CTA
constexpr int kStemThreads = 128;
constexpr int kStemInStride = 352;    // bf16 elements per staged input row: 8 lead (5 unused + 1 zero pixel) + 336 + 8
// Shared memory bounds the occupancy of this kernel (a chain of dependent steps per image row, so the rows in flight
// per SM are what hides its latencies): with a separate 16 KB output tile a CTA took 50.6 KB = 4 CTAs per SM.  The
// output tile now reuses the A tile (the MMAs of a row have finished reading it when the epilogue runs; the next
// row's gather waits until the TMA store has finished reading it): 34 KB = 6 CTAs per SM.  Without a kilobyte of
// alignment slack it is 32 192 B + 1 KB reserved = 7 CTAs per SM (7 x 64 TMEM columns and 7 x 128 x 64 registers fit as
// well): the dynamic window of a kernel without static shared memory starts 1024-aligned (CTA allocations are
// 1 KB-granular with 1 KB reserved in front); the kernel checks that and traps otherwise.
constexpr int kStemSmemBytes = 16384 /*A, then the output tile*/ + 8192 /*B*/ + (kStemRows + 2) * kStemInStride * 2 + 512 + 64;

// w: [64][32] bf16 K-major (k = (r*3+s)*3 + c, 27..31 zero), bias/prelu: [64] fp32.
// in: [B][H][W][3] bf16 with W == 112 (one TMEM lane per pixel of a row), out via tmOut: 2-D [B*H*W][64] bf16,
// box 64 x W, SWIZZLE_128B.
__global__ void __launch_bounds__(kStemThreads)
stem_tc_kernel(const __grid_constant__ CUtensorMap tmOut, const __nv_bfloat16* __restrict__ in,
               const __nv_bfloat16* __restrict__ w, const float* __restrict__ bias,
               const float* __restrict__ prelu, int H, int W, int* __restrict__ progress) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();   // the 128-byte swizzle of the A / B / output tiles needs it
  uint8_t* sA = smem;
  uint8_t* sB = smem + 16384;
  uint8_t* sOut = sA;
  __nv_bfloat16* sIn = reinterpret_cast<__nv_bfloat16*>(smem + 16384 + 8192);
  float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sIn) + (kStemRows + 2) * kStemInStride * 2);
  float* s_prelu = s_bias + 64;
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_prelu + 64);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int strips = H / kStemRows;
  const int img = blockIdx.x / strips;
  const int y0 = (blockIdx.x - img * strips) * kStemRows;

  if (tid == 0) {
    prefetch_tmap(&tmOut);
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr_smem, 64);
    tmem_relinquish();
  }
  // weights -> swizzled K-major B tile (row n = output channel, four 16-byte chunks of K)
  for (int i = tid; i < 64 * 4; i += kStemThreads) {
    const int n = i >> 2, c = i & 3;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(w) + i);
    *reinterpret_cast<uint4*>(sB + n * 128 + ((c ^ (n & 7)) << 4)) = v;
  }
  if (tid < 64) {
    s_bias[tid] = bias[tid];
    s_prelu[tid] = prelu[tid];
  }
  // A rows 112..127 are never gathered: zero them once so the unused accumulator lanes stay finite
  for (int i = tid; i < 16 * 8; i += kStemThreads) *reinterpret_cast<uint4*>(sA + 112 * 128 + i * 16) = make_uint4(0, 0, 0, 0);
  pdl_launch_dependents();
  pdl_wait();  // the input tensor is written by the preprocessing / warp kernel launched just before
  // stage the strip's input rows (y0-1 .. y0+kStemRows) with zero borders
  {
    const int row_u4 = (W * 3 * 2) / 16;  // 42 for W = 112
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < (kStemRows + 2) * (row_u4 + 2); i += kStemThreads) {
      const int rr = i / (row_u4 + 2), j = i - rr * (row_u4 + 2);  // j = 0: lead pad, 1..row_u4: pixels, row_u4+1: tail pad
      const int gy = y0 - 1 + rr;
      uint4 v = z;
      if (j >= 1 && j <= row_u4 && gy >= 0 && gy < H)
        v = __ldg(reinterpret_cast<const uint4*>(in + (static_cast<size_t>(img) * H + gy) * W * 3) + (j - 1));
      *reinterpret_cast<uint4*>(sIn + rr * kStemInStride + j * 8) = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
  constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
  uint32_t phase = 0;

  for (int ry = 0; ry < kStemRows; ++ry) {
    if (ry > 0) {   // the previous row's store has finished reading the tile the gather is about to overwrite
      if (tid == 0) tma_store_wait_read<0>();
      __syncthreads();
    }
    // ---- gather: pixel x = tid, K index (r*3+s)*3+c = 9 contiguous staged values (18 bytes) per filter row r.
    // ncu (profiles/r02_summary.md): this kernel is bound by the shared-memory data pipe (75 % of its wavefront peak),
    // most of it the 27 two-byte GENERIC loads per pixel this gather used to issue (7.5 wavefronts each).  Now: five
    // aligned ld.shared.u32 per filter row and funnel shifts - pixel x starts at byte 10 + 6x of the staged row, i.e. on
    // a word boundary for odd x and one halfword past it for even x.
    if (tid < W) {
      const uint32_t hp = (tid & 1) ? 0u : 16u;                         // halfword phase of this pixel inside its first word
      const uint32_t row0 = smem_u32(sIn) + static_cast<uint32_t>(ry) * (kStemInStride * 2) + ((10u + 6u * tid) & ~3u);
      uint32_t w[3][6];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int j = 0; j < 5; ++j)
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[r][j]) : "r"(row0 + r * (kStemInStride * 2) + 4 * j));
        w[r][5] = 0u;
      }
      // S(r, j) = values (2j, 2j+1) of filter row r; T(j) = values (2j+1, 2j+2) of filter row 1 (it starts at odd k = 9)
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) pk[j] = __funnelshift_r(w[0][j], w[0][j + 1], hp);                  // k 0..7
      const uint32_t e08 = __funnelshift_r(w[0][4], 0u, hp) & 0xffffu;                                 // k 8
      const uint32_t e10 = __funnelshift_r(w[1][0], 0u, hp) & 0xffffu;                                 // k 9
      pk[4] = e08 | (e10 << 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) pk[5 + j] = __funnelshift_rc(w[1][j], w[1][j + 1], 16u + hp);       // k 10..17
#pragma unroll
      for (int j = 0; j < 4; ++j) pk[9 + j] = __funnelshift_r(w[2][j], w[2][j + 1], hp);              // k 18..25
      pk[13] = __funnelshift_r(w[2][4], 0u, hp) & 0xffffu;                                            // k 26, 27 = 0
      pk[14] = 0u;
      pk[15] = 0u;
      const uint32_t a_row = smem_u32(sA) + tid * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a_row + ((c ^ (tid & 7)) << 4)), "r"(pk[4 * c]), "r"(pk[4 * c + 1]),
                     "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3]) : "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      umma_bf16_ss(tmem, umma_desc_sw128(a_addr), umma_desc_sw128(b_addr), idesc, 0u);
      umma_bf16_ss(tmem, umma_desc_sw128(a_addr) + 2, umma_desc_sw128(b_addr) + 2, idesc, 1u);
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: thread = pixel (TMEM lane), 64 channels in two 32-column loads
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      if (tid < W) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v[8], bz[8], pz[8];
          // explicit ld.shared.v4 (a broadcast: one wavefront each); the generic loads the compiler emitted for
          // s_bias[ch] / s_prelu[ch] were the other half of this kernel's shared-memory wavefronts
          const uint32_t bo = smem_u32(s_bias) + (c * 32 + j * 8) * 4, po = smem_u32(s_prelu) + (c * 32 + j * 8) * 4;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bz[0]), "=f"(bz[1]), "=f"(bz[2]), "=f"(bz[3]) : "r"(bo));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bz[4]), "=f"(bz[5]), "=f"(bz[6]), "=f"(bz[7]) : "r"(bo + 16));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pz[0]), "=f"(pz[1]), "=f"(pz[2]), "=f"(pz[3]) : "r"(po));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pz[4]), "=f"(pz[5]), "=f"(pz[6]), "=f"(pz[7]) : "r"(po + 16));
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float a = __uint_as_float(r[j * 8 + t]) + bz[t];
            v[t] = a > 0.f ? a : a * pz[t];
          }
          uint4 o;
          o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
          o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
          const int chunk = c * 4 + j;
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(sOut) + tid * 128 + ((chunk ^ (tid & 7)) << 4)), "r"(o.x),
                       "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tma_store_2d(&tmOut, sOut, 0, (img * H + y0 + ry) * W);
      tma_store_commit();
    }
  }
  if (tid == 0) {
    tma_store_wait<0>();  // all rows of this strip are written
    if (progress != nullptr) {  // inter-layer dataflow: kStemRows x W rows x (64 / 32) chunk units of image `img` are done
      fence_proxy_async_global();
      __threadfence();
      atomicAdd(progress + img, kStemRows * W * 2);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

}  // namespace frb
