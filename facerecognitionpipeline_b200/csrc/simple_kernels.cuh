// simple_kernels.cuh — the bandwidth-bound / small kernels of the embed-and-match path:
//   * warp_normalize : 5-landmark similarity warp (bit-exact restatement of cv2.warpAffine
//                      INTER_LINEAR/BORDER_CONSTANT fixed-point arithmetic) fused with the
//                      RGB->BGR + (x/255-0.5)/0.5 normalisation, emitting NHWC bf16
//   * preprocess_u8  : FaceEmbedder.preprocess for already-aligned crops (112, or 224 -> 112)
//   * stem_conv      : input_layer Conv3x3(3->64)+BN+PReLU (K = 27: CUDA-core direct conv)
//   * conv_ref       : slow direct convolution used only by the tests as an on-device checker
//   * fc_finalize    : split-K reduction + folded bias + L2 normalisation of the FC tail
//   * match_finalize / exact scan : exact f64 re-scoring, canonical top-k, threshold
#pragma once
#include "ptx.cuh"

namespace frb {

// ------------------------------------------------------------------ warp + normalise
// Per-face job. M is the FORWARD 2x3 matrix exactly as cv2.estimateAffinePartial2D returns it
// (reference face_recognition.py:64); the kernel inverts it the way cv::warpAffine does.
struct WarpJob {
  unsigned long long src_off;  // byte offset of the source image from the base pointer
  int H, W, pitch;             // source image geometry (pitch in bytes), RGB u8
  int _pad;
  double M[6];
};

// cv::warpAffine (imgwarp.cpp) constants
constexpr int kAbBits = 10, kAbScale = 1 << kAbBits, kInterBits = 5, kInterTab = 1 << kInterBits;

__device__ __forceinline__ int cv_round_sat(double v) {
  // cv::saturate_cast<int>(double) == cvRound == lrint (round half to even), saturating
  if (v >= 2147483647.0) return 2147483647;
  if (v <= -2147483648.0) return (-2147483647 - 1);
  return __double2int_rn(v);
}

// wtab: [32*32][4] uint16 bilinear weights (sum exactly 32768), built on the host the way
// cv::initInterTab2D builds BilinearTab_i.  lut: 256 bf16 bit patterns of the normalised value.
// One thread per output pixel (adjacent lanes read adjacent source pixels: the best coalescing this gather allows).
// The inverse matrix is a per-face constant: one thread per block computes it in the exact operation order of
// cv::invertAffineTransform and shares it through shared memory.  The 2x2 tap block is fetched as two 6-byte runs
// (pixels sx, sx+1 of rows sy, sy+1) with aligned 32-bit loads + funnel shifts instead of twelve byte loads, which
// halves the L1 tag traffic that bounds this kernel; taps that touch the border take the byte path (zero fill).
__device__ __forceinline__ void warp_load6(const uint8_t* p, uint32_t* lo, uint32_t* hi) {
  // bytes p[0..5] -> lo = p[0..3], hi = p[4..5] (upper half undefined); reads only aligned words that contain them
  const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(addr & ~static_cast<uintptr_t>(3));
  const uint32_t sh = static_cast<uint32_t>(addr & 3) * 8;
  const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1);
  const uint32_t w2 = (sh == 24) ? __ldg(w + 2) : 0u;
  *lo = __funnelshift_r(w0, w1, sh);
  *hi = __funnelshift_r(w1, w2, sh);
}
constexpr int kWarpThreads = 256, kWarpPixPerBlock = 2048;   // 8 pixels per thread, 256 apart (lanes stay adjacent)
constexpr int kWarpMaxS = 256;                               // output size limit of the table-driven kernels
constexpr int kWarpMaxRows = kWarpPixPerBlock / 64 + 2;      // rows a block can touch (S >= 64)
template <bool kWriteU8, bool kWriteBf16>
__global__ void __launch_bounds__(kWarpThreads)
warp_normalize_kernel(const uint8_t* __restrict__ src_base, const WarpJob* __restrict__ jobs, int S,
                      const unsigned short* __restrict__ wtab_g, const unsigned short* __restrict__ lut_g,
                      uint8_t* __restrict__ out_u8, __nv_bfloat16* __restrict__ out_bf16) {
  __shared__ double s_m[6];          // m00, m01, m10, m11, b1, b2 of the inverse map
  __shared__ WarpJob s_job;
  // the bilinear weight table (8 KB) and the normalisation LUT live in shared memory: every global load goes through
  // the L1 tag stage, which is what bounds this gather
  __shared__ __align__(8) unsigned short wtab[kInterTab * kInterTab * 4];
  __shared__ unsigned short lut[256];
  for (int i = threadIdx.x; i < kInterTab * kInterTab; i += kWarpThreads)
    reinterpret_cast<uint2*>(wtab)[i] = __ldg(reinterpret_cast<const uint2*>(wtab_g) + i);
  if (threadIdx.x < 256) lut[threadIdx.x] = lut_g[threadIdx.x];
  const int face = blockIdx.y;
  if (threadIdx.x == 0) {
    const WarpJob jb = jobs[face];
    s_job = jb;
    // invert (double, unfused: mirrors the scalar C++ in cv::warpAffine / cv::invertAffineTransform)
    double D = __dsub_rn(__dmul_rn(jb.M[0], jb.M[4]), __dmul_rn(jb.M[1], jb.M[3]));
    D = (D != 0.0) ? __ddiv_rn(1.0, D) : 0.0;
    const double A11 = __dmul_rn(jb.M[4], D), A22 = __dmul_rn(jb.M[0], D);
    const double m00 = A11;
    const double m01 = __dmul_rn(jb.M[1], -D);
    const double m10 = __dmul_rn(jb.M[3], -D);
    const double m11 = A22;
    s_m[0] = m00; s_m[1] = m01; s_m[2] = m10; s_m[3] = m11;
    s_m[4] = __dsub_rn(__dmul_rn(-m00, jb.M[2]), __dmul_rn(m01, jb.M[5]));
    s_m[5] = __dsub_rn(__dmul_rn(-m10, jb.M[2]), __dmul_rn(m11, jb.M[5]));
  }
  __syncthreads();
  const int srcH = s_job.H, srcW = s_job.W, pitch = s_job.pitch;
  const uint8_t* img = src_base + s_job.src_off;
  const int pix_begin = blockIdx.x * kWarpPixPerBlock;
  const int pix_end = min(S * S, pix_begin + kWarpPixPerBlock);
  // per-column and per-row fixed-point terms, exactly the adelta / bdelta / X0 / Y0 tables of cv::warpAffine: the f64
  // arithmetic leaves the per-pixel path (it was ~40 % of the instructions)
  __shared__ int s_adelta[kWarpMaxS], s_bdelta[kWarpMaxS], s_X0[kWarpMaxRows], s_Y0[kWarpMaxRows];
  const int y_first = pix_begin / S;
  {
    const double m00 = s_m[0], m01 = s_m[1], m10 = s_m[2], m11 = s_m[3], b1 = s_m[4], b2 = s_m[5];
    const int round_delta = kAbScale / kInterTab / 2;
    for (int x = threadIdx.x; x < S; x += kWarpThreads) {
      s_adelta[x] = cv_round_sat(__dmul_rn(__dmul_rn(m00, (double)x), (double)kAbScale));
      s_bdelta[x] = cv_round_sat(__dmul_rn(__dmul_rn(m10, (double)x), (double)kAbScale));
    }
    const int n_rows = (pix_end - 1) / S - y_first + 1;
    for (int r = threadIdx.x; r < n_rows; r += kWarpThreads) {
      const int y = y_first + r;
      s_X0[r] = cv_round_sat(__dmul_rn(__dadd_rn(__dmul_rn(m01, (double)y), b1), (double)kAbScale)) + round_delta;
      s_Y0[r] = cv_round_sat(__dmul_rn(__dadd_rn(__dmul_rn(m11, (double)y), b2), (double)kAbScale)) + round_delta;
    }
  }
  __syncthreads();
  for (int pix = pix_begin + threadIdx.x; pix < pix_end; pix += kWarpThreads) {
  const int y = pix / S, x = pix - y * S;
  const int X = (s_X0[y - y_first] + s_adelta[x]) >> (kAbBits - kInterBits);
  const int Y = (s_Y0[y - y_first] + s_bdelta[x]) >> (kAbBits - kInterBits);
  const int sx = X >> kInterBits, sy = Y >> kInterBits;
  const int ax = X & (kInterTab - 1), ay = Y & (kInterTab - 1);
  const ushort4 w4 = reinterpret_cast<const ushort4*>(wtab)[ay * kInterTab + ax];
  // weights are 0..32768 inclusive (the (0,0) phase is exactly 32768): keep them unsigned
  const int4 w = make_int4(w4.x, w4.y, w4.z, w4.w);
  int acc0 = 0, acc1 = 0, acc2 = 0;
  if (sx >= 0 && sx + 3 <= srcW && sy >= 0 && sy + 1 < srcH) {
    // interior: every byte of the aligned words lies inside the image (3 * sx + 9 <= 3 * W)
    const uint8_t* q0 = img + static_cast<size_t>(sy) * pitch + 3 * sx;
    uint32_t lo0, hi0, lo1, hi1;
    warp_load6(q0, &lo0, &hi0);
    warp_load6(q0 + pitch, &lo1, &hi1);
    acc0 = w.x * (lo0 & 0xff) + w.y * (lo0 >> 24) + w.z * (lo1 & 0xff) + w.w * (lo1 >> 24);
    acc1 = w.x * ((lo0 >> 8) & 0xff) + w.y * (hi0 & 0xff) + w.z * ((lo1 >> 8) & 0xff) + w.w * (hi1 & 0xff);
    acc2 = w.x * ((lo0 >> 16) & 0xff) + w.y * ((hi0 >> 8) & 0xff) + w.z * ((lo1 >> 16) & 0xff) + w.w * ((hi1 >> 8) & 0xff);
  } else {
    const bool x0ok = (sx >= 0 && sx < srcW), x1ok = (sx + 1 >= 0 && sx + 1 < srcW);
    if (sy >= 0 && sy < srcH) {
      const uint8_t* rowp = img + static_cast<size_t>(sy) * pitch;
      if (x0ok) {
        const uint8_t* q = rowp + 3 * sx;
        acc0 += w.x * q[0]; acc1 += w.x * q[1]; acc2 += w.x * q[2];
      }
      if (x1ok) {
        const uint8_t* q = rowp + 3 * (sx + 1);
        acc0 += w.y * q[0]; acc1 += w.y * q[1]; acc2 += w.y * q[2];
      }
    }
    if (sy + 1 >= 0 && sy + 1 < srcH) {
      const uint8_t* rowp = img + static_cast<size_t>(sy + 1) * pitch;
      if (x0ok) {
        const uint8_t* q = rowp + 3 * sx;
        acc0 += w.z * q[0]; acc1 += w.z * q[1]; acc2 += w.z * q[2];
      }
      if (x1ok) {
        const uint8_t* q = rowp + 3 * (sx + 1);
        acc0 += w.w * q[0]; acc1 += w.w * q[1]; acc2 += w.w * q[2];
      }
    }
  }
  // FixedPtCast<int, uchar, 15>: (v + 2^14) >> 15, saturated
  const int r = min(255, max(0, (acc0 + (1 << 14)) >> 15));
  const int g = min(255, max(0, (acc1 + (1 << 14)) >> 15));
  const int b = min(255, max(0, (acc2 + (1 << 14)) >> 15));
  const size_t o = (static_cast<size_t>(face) * S * S + pix) * 3;
  if (kWriteU8) {
    out_u8[o] = (uint8_t)r; out_u8[o + 1] = (uint8_t)g; out_u8[o + 2] = (uint8_t)b;
  }
  if (kWriteBf16) {  // model is fed BGR (reference face_embedder.py:99)
    unsigned short* ob = reinterpret_cast<unsigned short*>(out_bf16) + o;
    ob[0] = lut[b]; ob[1] = lut[g]; ob[2] = lut[r];
  }
  }  // pixel loop
}

// ---- staged variant: the whole source region of a face first goes to shared memory with coalesced 16-byte loads
// (one block per face, up to ~200 KB), then the gather reads shared memory.  The global-memory gather above is bound
// by L1 wavefronts (10.8 sectors per request, 31 % of HBM speed); this one reads every source byte once, in order.
// MEASURED slower than the gather above (1.14 vs 0.75 ms for 8192 faces): with ~200 KB per block only one block fits
// an SM, so its load phase and its gather phase never overlap.  Kept behind FRB_WARP_STAGED=1.
// Used when every face's source box fits (frb_warp_normalize checks on the host); identical arithmetic, so the
// output is the same bytes.  box = inclusive source pixel range [x0, x1] x [y0, y1] that contains every interior tap.
constexpr int kWarpStagedThreads = 1024;
constexpr int kWarpStagedMaxBytes = 208 * 1024;
__host__ __device__ __forceinline__ int warp_staged_pitch(int x0, int x1) { return (((x1 - x0 + 1) * 3 + 30) / 16) * 16; }

__device__ __forceinline__ void warp_load6_smem(const uint8_t* p, uint32_t* lo, uint32_t* hi) {
  const uint32_t a = smem_u32(p);
  const uint32_t sh = (a & 3u) * 8u;
  uint32_t w0, w1, w2;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(a & ~3u));
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"((a & ~3u) + 4u));
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w2) : "r"((a & ~3u) + 8u));
  *lo = __funnelshift_r(w0, w1, sh);
  *hi = __funnelshift_r(w1, w2, sh);
}

template <bool kWriteU8, bool kWriteBf16>
__global__ void __launch_bounds__(kWarpStagedThreads)
warp_normalize_staged_kernel(const uint8_t* __restrict__ src_base, const WarpJob* __restrict__ jobs,
                             const int4* __restrict__ boxes, int S, const unsigned short* __restrict__ wtab_g,
                             const unsigned short* __restrict__ lut_g, uint8_t* __restrict__ out_u8,
                             __nv_bfloat16* __restrict__ out_bf16) {
  extern __shared__ __align__(16) uint8_t s_src[];
  __shared__ double s_m[6];
  __shared__ WarpJob s_job;
  __shared__ int4 s_box;
  __shared__ __align__(8) unsigned short wtab[kInterTab * kInterTab * 4];
  __shared__ unsigned short lut[256];
  const int face = blockIdx.x;
  for (int i = threadIdx.x; i < kInterTab * kInterTab; i += kWarpStagedThreads)
    reinterpret_cast<uint2*>(wtab)[i] = __ldg(reinterpret_cast<const uint2*>(wtab_g) + i);
  if (threadIdx.x < 256) lut[threadIdx.x] = lut_g[threadIdx.x];
  if (threadIdx.x == 0) {
    const WarpJob jb = jobs[face];
    s_job = jb;
    s_box = boxes[face];
    double D = __dsub_rn(__dmul_rn(jb.M[0], jb.M[4]), __dmul_rn(jb.M[1], jb.M[3]));
    D = (D != 0.0) ? __ddiv_rn(1.0, D) : 0.0;
    const double A11 = __dmul_rn(jb.M[4], D), A22 = __dmul_rn(jb.M[0], D);
    const double m00 = A11;
    const double m01 = __dmul_rn(jb.M[1], -D);
    const double m10 = __dmul_rn(jb.M[3], -D);
    const double m11 = A22;
    s_m[0] = m00; s_m[1] = m01; s_m[2] = m10; s_m[3] = m11;
    s_m[4] = __dsub_rn(__dmul_rn(-m00, jb.M[2]), __dmul_rn(m01, jb.M[5]));
    s_m[5] = __dsub_rn(__dmul_rn(-m10, jb.M[2]), __dmul_rn(m11, jb.M[5]));
  }
  __syncthreads();
  __shared__ int s_adelta[kWarpMaxS], s_bdelta[kWarpMaxS], s_X0[kWarpMaxS], s_Y0[kWarpMaxS];
  {
    const double m00 = s_m[0], m01 = s_m[1], m10 = s_m[2], m11 = s_m[3], b1 = s_m[4], b2 = s_m[5];
    const int round_delta = kAbScale / kInterTab / 2;
    for (int i = threadIdx.x; i < S; i += kWarpStagedThreads) {
      s_adelta[i] = cv_round_sat(__dmul_rn(__dmul_rn(m00, (double)i), (double)kAbScale));
      s_bdelta[i] = cv_round_sat(__dmul_rn(__dmul_rn(m10, (double)i), (double)kAbScale));
      s_X0[i] = cv_round_sat(__dmul_rn(__dadd_rn(__dmul_rn(m01, (double)i), b1), (double)kAbScale)) + round_delta;
      s_Y0[i] = cv_round_sat(__dmul_rn(__dadd_rn(__dmul_rn(m11, (double)i), b2), (double)kAbScale)) + round_delta;
    }
  }
  const int srcH = s_job.H, srcW = s_job.W, pitch = s_job.pitch;
  const uint8_t* img = src_base + s_job.src_off;
  const uint8_t* img_end = img + static_cast<size_t>(srcH) * pitch;
  const int bx0 = s_box.x, by0 = s_box.y, bx1 = s_box.z, by1 = s_box.w;
  const int rows = by1 >= by0 && bx1 >= bx0 ? by1 - by0 + 1 : 0;
  const int rowbytes = (bx1 - bx0 + 1) * 3;
  const int spitch = warp_staged_pitch(bx0, bx1);
  const int cpr = spitch / 16;
  // ---- stage: 16-byte chunks keep their global alignment phase (row r starts at s_src[r * spitch + phase_r])
  // Four chunks per thread are requested before the first one is stored: the loop is latency-bound otherwise (one
  // DRAM round trip per iteration made this kernel slower than the global-memory gather).
  constexpr int kUnroll = 4;
  const int total_chunks = rows * cpr;
  for (int c0 = threadIdx.x; c0 < total_chunks; c0 += kWarpStagedThreads * kUnroll) {
    uint4 v[kUnroll];
    int dst[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int c = c0 + u * kWarpStagedThreads;
      dst[u] = -1;
      v[u] = make_uint4(0, 0, 0, 0);
      if (c < total_chunks) {
        const int r = c / cpr, k = c - r * cpr;
        const uint8_t* g = img + static_cast<size_t>(by0 + r) * pitch + bx0 * 3;
        const int ph = static_cast<int>(reinterpret_cast<uintptr_t>(g) & 15);
        if (k * 16 < ph + rowbytes) {
          const uint8_t* ga = g - ph + k * 16;
          dst[u] = r * spitch + k * 16;
          if (ga >= img && ga + 16 <= img_end) {
            v[u] = __ldg(reinterpret_cast<const uint4*>(ga));
          } else {  // chunk straddles the first / last bytes of the image: never read outside it
            unsigned char t[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) t[i] = (ga + i >= img && ga + i < img_end) ? ga[i] : 0;
            v[u].x = t[0] | (t[1] << 8) | (t[2] << 16) | (static_cast<uint32_t>(t[3]) << 24);
            v[u].y = t[4] | (t[5] << 8) | (t[6] << 16) | (static_cast<uint32_t>(t[7]) << 24);
            v[u].z = t[8] | (t[9] << 8) | (t[10] << 16) | (static_cast<uint32_t>(t[11]) << 24);
            v[u].w = t[12] | (t[13] << 8) | (t[14] << 16) | (static_cast<uint32_t>(t[15]) << 24);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (dst[u] >= 0) *reinterpret_cast<uint4*>(s_src + dst[u]) = v[u];
  }
  __syncthreads();
  // ---- gather (fixed-point terms from the per-column / per-row tables, as cv::warpAffine)
  for (int pix = threadIdx.x; pix < S * S; pix += kWarpStagedThreads) {
    const int y = pix / S, x = pix - y * S;
    const int X = (s_X0[y] + s_adelta[x]) >> (kAbBits - kInterBits);
    const int Y = (s_Y0[y] + s_bdelta[x]) >> (kAbBits - kInterBits);
    const int sx = X >> kInterBits, sy = Y >> kInterBits;
    const int ax = X & (kInterTab - 1), ay = Y & (kInterTab - 1);
    const ushort4 w4 = reinterpret_cast<const ushort4*>(wtab)[ay * kInterTab + ax];
    const int4 w = make_int4(w4.x, w4.y, w4.z, w4.w);
    int acc0 = 0, acc1 = 0, acc2 = 0;
    if (sx >= bx0 && sx + 1 <= bx1 && sy >= by0 && sy + 1 <= by1) {
      // both rows and both columns are staged (the box is clamped to the image, so all four taps are in bounds)
      const int r = sy - by0;
      const uintptr_t g0 = reinterpret_cast<uintptr_t>(img) + static_cast<size_t>(sy) * pitch + bx0 * 3;
      const int ph0 = static_cast<int>(g0 & 15), ph1 = static_cast<int>((g0 + pitch) & 15);
      uint32_t lo0, hi0, lo1, hi1;
      warp_load6_smem(s_src + static_cast<size_t>(r) * spitch + ph0 + (sx - bx0) * 3, &lo0, &hi0);
      warp_load6_smem(s_src + static_cast<size_t>(r + 1) * spitch + ph1 + (sx - bx0) * 3, &lo1, &hi1);
      acc0 = w.x * (lo0 & 0xff) + w.y * (lo0 >> 24) + w.z * (lo1 & 0xff) + w.w * (lo1 >> 24);
      acc1 = w.x * ((lo0 >> 8) & 0xff) + w.y * (hi0 & 0xff) + w.z * ((lo1 >> 8) & 0xff) + w.w * (hi1 & 0xff);
      acc2 = w.x * ((lo0 >> 16) & 0xff) + w.y * ((hi0 >> 8) & 0xff) + w.z * ((lo1 >> 16) & 0xff) + w.w * ((hi1 >> 8) & 0xff);
    } else {
      const bool x0ok = (sx >= 0 && sx < srcW), x1ok = (sx + 1 >= 0 && sx + 1 < srcW);
      if (sy >= 0 && sy < srcH) {
        const uint8_t* rowp = img + static_cast<size_t>(sy) * pitch;
        if (x0ok) { const uint8_t* q = rowp + 3 * sx; acc0 += w.x * q[0]; acc1 += w.x * q[1]; acc2 += w.x * q[2]; }
        if (x1ok) { const uint8_t* q = rowp + 3 * (sx + 1); acc0 += w.y * q[0]; acc1 += w.y * q[1]; acc2 += w.y * q[2]; }
      }
      if (sy + 1 >= 0 && sy + 1 < srcH) {
        const uint8_t* rowp = img + static_cast<size_t>(sy + 1) * pitch;
        if (x0ok) { const uint8_t* q = rowp + 3 * sx; acc0 += w.z * q[0]; acc1 += w.z * q[1]; acc2 += w.z * q[2]; }
        if (x1ok) { const uint8_t* q = rowp + 3 * (sx + 1); acc0 += w.w * q[0]; acc1 += w.w * q[1]; acc2 += w.w * q[2]; }
      }
    }
    const int r8 = min(255, max(0, (acc0 + (1 << 14)) >> 15));
    const int g8 = min(255, max(0, (acc1 + (1 << 14)) >> 15));
    const int b8 = min(255, max(0, (acc2 + (1 << 14)) >> 15));
    const size_t o = (static_cast<size_t>(face) * S * S + pix) * 3;
    if (kWriteU8) {
      out_u8[o] = (uint8_t)r8; out_u8[o + 1] = (uint8_t)g8; out_u8[o + 2] = (uint8_t)b8;
    }
    if (kWriteBf16) {
      unsigned short* ob = reinterpret_cast<unsigned short*>(out_bf16) + o;
      ob[0] = lut[b8]; ob[1] = lut[g8]; ob[2] = lut[r8];
    }
  }
}

// ---- band-staged variant (round 2; FRB_WARP_BAND=1, off: measured 3x slower than the direct gather): one block = one band of whole output rows (~2048 pixels) of one face.
// The band's source footprint is a parallelogram; for every source row the block records the exact pixel span its
// taps touch (computed from the SAME fixed-point coordinates the gather uses), copies those spans - and nothing else -
// into shared memory with 16-byte loads that keep their global alignment phase, and gathers the 2x2 taps from there.
// ~30 KB per block instead of the ~200 KB of the whole-face variant above, so 4-5 blocks share an SM and the copy of
// one overlaps the gather of another.  Global memory sees every source byte of a band once, in 16-byte pieces, instead
// of the 10.8 sectors per request of the direct gather.  Pixels whose taps leave the image (or whose rows did not fit
// the staging buffer) take the same global paths as warp_normalize_kernel: identical arithmetic, identical bytes out.
constexpr int kWarpBandThreads = 256;
constexpr int kWarpBandMaxRows = 224;             // source rows a band may stage
constexpr int kWarpBandStageBytes = 36 * 1024;    // dynamic shared memory for the staged spans (+16 slack inside)
__host__ __device__ __forceinline__ int warp_band_rows(int S) { return kWarpPixPerBlock / S > 0 ? kWarpPixPerBlock / S : 1; }

template <bool kWriteU8, bool kWriteBf16>
__global__ void __launch_bounds__(kWarpBandThreads)
warp_normalize_band_kernel(const uint8_t* __restrict__ src_base, const WarpJob* __restrict__ jobs, int S,
                           const unsigned short* __restrict__ wtab_g, const unsigned short* __restrict__ lut_g,
                           uint8_t* __restrict__ out_u8, __nv_bfloat16* __restrict__ out_bf16) {
  extern __shared__ __align__(16) uint8_t s_stage[];
  __shared__ double s_m[6];
  __shared__ WarpJob s_job;
  __shared__ __align__(8) unsigned short wtab[kInterTab * kInterTab * 4];
  __shared__ unsigned short lut[256];
  __shared__ int s_adelta[kWarpMaxS], s_bdelta[kWarpMaxS], s_X0[kWarpMaxRows], s_Y0[kWarpMaxRows];
  __shared__ int s_xs[kWarpBandMaxRows], s_xe[kWarpBandMaxRows];   // staged pixel span of source row Ysrc0 + i (xe < xs: none)
  __shared__ int s_delta[kWarpBandMaxRows];                        // s_stage offset of (that row's pixel 0)
  __shared__ int s_cstart[kWarpBandMaxRows + 1];                   // prefix sums of the rows' 16-byte chunk counts
  __shared__ int s_ysrc0, s_nrows;
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < kInterTab * kInterTab; i += kWarpBandThreads)
    reinterpret_cast<uint2*>(wtab)[i] = __ldg(reinterpret_cast<const uint2*>(wtab_g) + i);
  if (tid < 256) lut[tid] = lut_g[tid];
  const int face = blockIdx.y;
  if (tid == 0) {
    const WarpJob jb = jobs[face];
    s_job = jb;
    double D = __dsub_rn(__dmul_rn(jb.M[0], jb.M[4]), __dmul_rn(jb.M[1], jb.M[3]));
    D = (D != 0.0) ? __ddiv_rn(1.0, D) : 0.0;
    const double A11 = __dmul_rn(jb.M[4], D), A22 = __dmul_rn(jb.M[0], D);
    const double m00 = A11;
    const double m01 = __dmul_rn(jb.M[1], -D);
    const double m10 = __dmul_rn(jb.M[3], -D);
    const double m11 = A22;
    s_m[0] = m00; s_m[1] = m01; s_m[2] = m10; s_m[3] = m11;
    s_m[4] = __dsub_rn(__dmul_rn(-m00, jb.M[2]), __dmul_rn(m01, jb.M[5]));
    s_m[5] = __dsub_rn(__dmul_rn(-m10, jb.M[2]), __dmul_rn(m11, jb.M[5]));
  }
  for (int i = tid; i < kWarpBandMaxRows; i += kWarpBandThreads) {
    s_xs[i] = 0x7fffffff;
    s_xe[i] = -1;
  }
  __syncthreads();
  const int srcH = s_job.H, srcW = s_job.W, pitch = s_job.pitch;
  const uint8_t* img = src_base + s_job.src_off;
  const uint8_t* img_end = img + static_cast<size_t>(srcH) * pitch;
  const int rows_per_band = warp_band_rows(S);
  const int y_first = blockIdx.x * rows_per_band;
  const int y_last = min(S, y_first + rows_per_band) - 1;
  const int pix_begin = y_first * S, pix_end = (y_last + 1) * S;
  {
    const double m00 = s_m[0], m01 = s_m[1], m10 = s_m[2], m11 = s_m[3], b1 = s_m[4], b2 = s_m[5];
    const int round_delta = kAbScale / kInterTab / 2;
    for (int x = tid; x < S; x += kWarpBandThreads) {
      s_adelta[x] = cv_round_sat(__dmul_rn(__dmul_rn(m00, (double)x), (double)kAbScale));
      s_bdelta[x] = cv_round_sat(__dmul_rn(__dmul_rn(m10, (double)x), (double)kAbScale));
    }
    for (int r = tid; r <= y_last - y_first; r += kWarpBandThreads) {
      const int y = y_first + r;
      s_X0[r] = cv_round_sat(__dmul_rn(__dadd_rn(__dmul_rn(m01, (double)y), b1), (double)kAbScale)) + round_delta;
      s_Y0[r] = cv_round_sat(__dmul_rn(__dadd_rn(__dmul_rn(m11, (double)y), b2), (double)kAbScale)) + round_delta;
    }
  }
  __syncthreads();
  if (tid == 0) {
    // source rows of the band: sy is monotone in x and in y, so its extremes sit at the band's corner pixels
    int lo = 0x7fffffff, hi = -0x7fffffff;
    for (int c = 0; c < 4; ++c) {
      const int x = (c & 1) ? S - 1 : 0, r = (c & 2) ? y_last - y_first : 0;
      const int sy = (s_Y0[r] + s_bdelta[x]) >> kAbBits;
      lo = min(lo, sy);
      hi = max(hi, sy + 1);
    }
    lo = max(lo, 0);
    hi = min(hi, srcH - 1);
    s_ysrc0 = lo;
    s_nrows = (hi >= lo) ? min(hi - lo + 1, kWarpBandMaxRows) : 0;
  }
  __syncthreads();
  const int ysrc0 = s_ysrc0, nrows = s_nrows;
  // ---- pass 1: the pixel span every source row needs (interior taps only; the rest never reads shared memory)
  for (int pix0 = pix_begin + (tid & ~31); pix0 < pix_end; pix0 += kWarpBandThreads) {
    const int pix = pix0 + lane;
    int sx = 0, sy = -0x40000000;
    bool interior = false;
    if (pix < pix_end) {
      const int y = pix / S, x = pix - y * S;
      sx = (s_X0[y - y_first] + s_adelta[x]) >> kAbBits;
      sy = (s_Y0[y - y_first] + s_bdelta[x]) >> kAbBits;
      interior = sx >= 0 && sx + 3 <= srcW && sy >= 0 && sy + 1 < srcH && sy - ysrc0 >= 0 && sy + 1 - ysrc0 < nrows;
    }
    const unsigned act = __ballot_sync(0xffffffffu, interior);
    if (interior) {
      const unsigned grp = __match_any_sync(act, sy);      // lanes of this warp that read the same two source rows
      const int lo = __reduce_min_sync(grp, sx), hi = __reduce_max_sync(grp, sx) + 1;
      if (lane == __ffs(grp) - 1) {
        const int i = sy - ysrc0;
        atomicMin(&s_xs[i], lo); atomicMax(&s_xe[i], hi);
        atomicMin(&s_xs[i + 1], lo); atomicMax(&s_xe[i + 1], hi);
      }
    }
  }
  __syncthreads();
  // ---- pass 2: lay the spans out as 16-byte chunks that keep their global alignment phase
  if (tid < 32) {
    int carry = 0;
    for (int base = 0; base < nrows; base += 32) {
      const int i = base + lane;
      int nch = 0;
      if (i < nrows && s_xe[i] >= s_xs[i]) {
        const uintptr_t g0 = reinterpret_cast<uintptr_t>(img) + static_cast<size_t>(ysrc0 + i) * pitch + 3 * s_xs[i];
        const uintptr_t g1 = reinterpret_cast<uintptr_t>(img) + static_cast<size_t>(ysrc0 + i) * pitch + 3 * (s_xe[i] + 1);
        nch = static_cast<int>((g1 - (g0 & ~static_cast<uintptr_t>(15)) + 15) >> 4);
      }
      int incl = nch;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      int start = carry + incl - nch;
      if (i < nrows) {
        if ((start + nch) * 16 + 16 > kWarpBandStageBytes) {   // does not fit: this row is not staged (global path)
          s_xs[i] = 0x7fffffff;
          s_xe[i] = -1;
          nch = 0;
        }
      }
      // rows dropped for capacity keep the layout of the rows before them intact: recompute the inclusive scan
      incl = nch;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      start = carry + incl - nch;
      if (i < nrows) {
        s_cstart[i] = start;
        if (nch > 0) {
          const uintptr_t g0 = reinterpret_cast<uintptr_t>(img) + static_cast<size_t>(ysrc0 + i) * pitch + 3 * s_xs[i];
          s_delta[i] = start * 16 + static_cast<int>(g0 & 15) - 3 * s_xs[i];
        }
      }
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_cstart[nrows] = carry;
  }
  __syncthreads();
  // ---- pass 3: copy (four 16-byte requests in flight per thread before the first is stored)
  {
    const int total = s_cstart[nrows];
    constexpr int kUnroll = 4;
    for (int c0 = tid; c0 < total; c0 += kWarpBandThreads * kUnroll) {
      uint4 v[kUnroll];
      int dst[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int c = c0 + u * kWarpBandThreads;
        dst[u] = -1;
        v[u] = make_uint4(0, 0, 0, 0);
        if (c < total) {
          int lo = 0, hi = nrows - 1;             // the row whose chunk range contains c
          while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_cstart[mid] <= c) lo = mid; else hi = mid - 1;
          }
          const int i = lo, k = c - s_cstart[i];
          const uintptr_t g0 = reinterpret_cast<uintptr_t>(img) + static_cast<size_t>(ysrc0 + i) * pitch + 3 * s_xs[i];
          const uint8_t* ga = reinterpret_cast<const uint8_t*>(g0 & ~static_cast<uintptr_t>(15)) + k * 16;
          dst[u] = c * 16;
          if (ga >= img && ga + 16 <= img_end) {
            v[u] = __ldg(reinterpret_cast<const uint4*>(ga));
          } else {  // chunk straddles the first / last bytes of the image: never read outside it
            unsigned char t[16];
#pragma unroll
            for (int b = 0; b < 16; ++b) t[b] = (ga + b >= img && ga + b < img_end) ? ga[b] : 0;
            v[u].x = t[0] | (t[1] << 8) | (t[2] << 16) | (static_cast<uint32_t>(t[3]) << 24);
            v[u].y = t[4] | (t[5] << 8) | (t[6] << 16) | (static_cast<uint32_t>(t[7]) << 24);
            v[u].z = t[8] | (t[9] << 8) | (t[10] << 16) | (static_cast<uint32_t>(t[11]) << 24);
            v[u].w = t[12] | (t[13] << 8) | (t[14] << 16) | (static_cast<uint32_t>(t[15]) << 24);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        if (dst[u] >= 0) *reinterpret_cast<uint4*>(s_stage + dst[u]) = v[u];
    }
  }
  __syncthreads();
  // ---- pass 4: gather
  for (int pix = pix_begin + tid; pix < pix_end; pix += kWarpBandThreads) {
    const int y = pix / S, x = pix - y * S;
    const int X = (s_X0[y - y_first] + s_adelta[x]) >> (kAbBits - kInterBits);
    const int Y = (s_Y0[y - y_first] + s_bdelta[x]) >> (kAbBits - kInterBits);
    const int sx = X >> kInterBits, sy = Y >> kInterBits;
    const int ax = X & (kInterTab - 1), ay = Y & (kInterTab - 1);
    const ushort4 w4 = reinterpret_cast<const ushort4*>(wtab)[ay * kInterTab + ax];
    const int4 w = make_int4(w4.x, w4.y, w4.z, w4.w);
    int acc0 = 0, acc1 = 0, acc2 = 0;
    if (sx >= 0 && sx + 3 <= srcW && sy >= 0 && sy + 1 < srcH) {
      uint32_t lo0, hi0, lo1, hi1;
      const int i = sy - ysrc0;
      if (i >= 0 && i + 1 < nrows && sx >= s_xs[i] && sx + 1 <= s_xe[i] && sx >= s_xs[i + 1] && sx + 1 <= s_xe[i + 1]) {
        warp_load6_smem(s_stage + s_delta[i] + 3 * sx, &lo0, &hi0);
        warp_load6_smem(s_stage + s_delta[i + 1] + 3 * sx, &lo1, &hi1);
      } else {   // rows that did not fit the staging buffer
        const uint8_t* q0 = img + static_cast<size_t>(sy) * pitch + 3 * sx;
        warp_load6(q0, &lo0, &hi0);
        warp_load6(q0 + pitch, &lo1, &hi1);
      }
      acc0 = w.x * (lo0 & 0xff) + w.y * (lo0 >> 24) + w.z * (lo1 & 0xff) + w.w * (lo1 >> 24);
      acc1 = w.x * ((lo0 >> 8) & 0xff) + w.y * (hi0 & 0xff) + w.z * ((lo1 >> 8) & 0xff) + w.w * (hi1 & 0xff);
      acc2 = w.x * ((lo0 >> 16) & 0xff) + w.y * ((hi0 >> 8) & 0xff) + w.z * ((lo1 >> 16) & 0xff) + w.w * ((hi1 >> 8) & 0xff);
    } else {
      const bool x0ok = (sx >= 0 && sx < srcW), x1ok = (sx + 1 >= 0 && sx + 1 < srcW);
      if (sy >= 0 && sy < srcH) {
        const uint8_t* rowp = img + static_cast<size_t>(sy) * pitch;
        if (x0ok) { const uint8_t* q = rowp + 3 * sx; acc0 += w.x * q[0]; acc1 += w.x * q[1]; acc2 += w.x * q[2]; }
        if (x1ok) { const uint8_t* q = rowp + 3 * (sx + 1); acc0 += w.y * q[0]; acc1 += w.y * q[1]; acc2 += w.y * q[2]; }
      }
      if (sy + 1 >= 0 && sy + 1 < srcH) {
        const uint8_t* rowp = img + static_cast<size_t>(sy + 1) * pitch;
        if (x0ok) { const uint8_t* q = rowp + 3 * sx; acc0 += w.z * q[0]; acc1 += w.z * q[1]; acc2 += w.z * q[2]; }
        if (x1ok) { const uint8_t* q = rowp + 3 * (sx + 1); acc0 += w.w * q[0]; acc1 += w.w * q[1]; acc2 += w.w * q[2]; }
      }
    }
    const int r8 = min(255, max(0, (acc0 + (1 << 14)) >> 15));
    const int g8 = min(255, max(0, (acc1 + (1 << 14)) >> 15));
    const int b8 = min(255, max(0, (acc2 + (1 << 14)) >> 15));
    const size_t o = (static_cast<size_t>(face) * S * S + pix) * 3;
    if (kWriteU8) {
      out_u8[o] = (uint8_t)r8; out_u8[o + 1] = (uint8_t)g8; out_u8[o + 2] = (uint8_t)b8;
    }
    if (kWriteBf16) {
      unsigned short* ob = reinterpret_cast<unsigned short*>(out_bf16) + o;
      ob[0] = lut[b8]; ob[1] = lut[g8]; ob[2] = lut[r8];
    }
  }
}

// ------------------------------------------------------------------ preprocess (aligned crops)
// in: [B][S][S][3] RGB u8, S in {112, 224}.  224 -> 112 is cv2.resize INTER_LINEAR at exactly
// 2x, which equals the 2x2 box mean rounded half up.  flip: also emit the horizontally flipped
// crop as image B+b (flip-fusion, BASELINE config 5).
__global__ void __launch_bounds__(256)
preprocess_u8_kernel(const uint8_t* __restrict__ in, int B, int S, const unsigned short* __restrict__ lut,
                     __nv_bfloat16* __restrict__ out, int flip) {
  const size_t total = static_cast<size_t>(B) * 112 * 112;
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int b = static_cast<int>(i / (112 * 112));
  const int pix = static_cast<int>(i - static_cast<size_t>(b) * 112 * 112);
  const int y = pix / 112, x = pix - y * 112;
  int r, g, bl;
  if (S == 112) {
    const uint8_t* q = in + (static_cast<size_t>(b) * 112 * 112 + pix) * 3;
    r = q[0]; g = q[1]; bl = q[2];
  } else {
    const uint8_t* q0 = in + ((static_cast<size_t>(b) * 224 + 2 * y) * 224 + 2 * x) * 3;
    const uint8_t* q1 = q0 + 224 * 3;
    r = (q0[0] + q0[3] + q1[0] + q1[3] + 2) >> 2;
    g = (q0[1] + q0[4] + q1[1] + q1[4] + 2) >> 2;
    bl = (q0[2] + q0[5] + q1[2] + q1[5] + 2) >> 2;
  }
  unsigned short* o = reinterpret_cast<unsigned short*>(out);
  const size_t d = (static_cast<size_t>(b) * 112 * 112 + pix) * 3;
  o[d] = lut[bl]; o[d + 1] = lut[g]; o[d + 2] = lut[r];
  if (flip) {
    const size_t f = ((static_cast<size_t>(B + b) * 112 + y) * 112 + (111 - x)) * 3;
    o[f] = lut[bl]; o[f + 1] = lut[g]; o[f + 2] = lut[r];
  }
}

// ------------------------------------------------------------------ reference conv (tests only)
// Same contract as the tcgen05 implicit-GEMM conv (GemmParams semantics), one thread per output
// element, fp32 accumulate in K order.  Used by tests as an on-device checker for big shapes.
struct ConvRefParams {
  const __nv_bfloat16* in;   // [B][H][W][Cin]
  const __nv_bfloat16* sc;   // [B][SH][SW][Csc] or nullptr
  const __nv_bfloat16* wt;   // [Cout][taps*Cin + Csc]
  const float* bias; int bias_cases;
  const float* prelu;
  const __nv_bfloat16* residual; int res_stride, RH, RW;
  __nv_bfloat16* out;        // [B][P][Q][Cout]
  int B, H, W, Cin, Cout, P, Q, stride, ksize, pad;
  int SH, SW, Csc, sc_stride;
};

__global__ void conv_ref_kernel(const ConvRefParams p) {
  const size_t total = static_cast<size_t>(p.B) * p.P * p.Q * p.Cout;
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int co = static_cast<int>(i % p.Cout);
  const size_t m = i / p.Cout;
  const int qq = static_cast<int>(m % p.Q);
  const int pp = static_cast<int>((m / p.Q) % p.P);
  const int img = static_cast<int>(m / (static_cast<size_t>(p.P) * p.Q));
  const int Ktot = p.ksize * p.ksize * p.Cin + p.Csc;
  const __nv_bfloat16* wrow = p.wt + static_cast<size_t>(co) * Ktot;
  float acc = 0.f;
  for (int r = 0; r < p.ksize; ++r)
    for (int s = 0; s < p.ksize; ++s) {
      const int iy = pp * p.stride - p.pad + r, ix = qq * p.stride - p.pad + s;
      if (iy < 0 || iy >= p.H || ix < 0 || ix >= p.W) continue;
      const __nv_bfloat16* x = p.in + ((static_cast<size_t>(img) * p.H + iy) * p.W + ix) * p.Cin;
      const __nv_bfloat16* ww = wrow + (r * p.ksize + s) * p.Cin;
      for (int c = 0; c < p.Cin; ++c) acc = fmaf(__bfloat162float(x[c]), __bfloat162float(ww[c]), acc);
    }
  if (p.sc != nullptr) {
    const __nv_bfloat16* x = p.sc + ((static_cast<size_t>(img) * p.SH + pp * p.sc_stride) * p.SW + qq * p.sc_stride) * p.Csc;
    const __nv_bfloat16* ww = wrow + p.ksize * p.ksize * p.Cin;
    for (int c = 0; c < p.Csc; ++c) acc = fmaf(__bfloat162float(x[c]), __bfloat162float(ww[c]), acc);
  }
  int bc = 0;
  if (p.bias_cases == 9) {
    const int rc = (pp == 0) ? 0 : ((pp == p.P - 1) ? 2 : 1);
    const int cc = (qq == 0) ? 0 : ((qq == p.Q - 1) ? 2 : 1);
    bc = rc * 3 + cc;
  }
  if (p.bias) acc += p.bias[static_cast<size_t>(bc) * p.Cout + co];
  if (p.prelu) acc = acc > 0.f ? acc : acc * p.prelu[co];
  if (p.residual)
    acc += __bfloat162float(p.residual[((static_cast<size_t>(img) * p.RH + pp * p.res_stride) * p.RW + qq * p.res_stride) * p.Cout + co]);
  p.out[i] = __float2bfloat16_rn(acc);
}

// ------------------------------------------------------------------ FC tail finalize
// partial: [splits][B][512] fp32.  x = sum(partials) + bias; AdaFace: norm = ||x||2, emb = x/norm
// (upstream net.py Backbone.forward).  l2 = 0 keeps the raw BN1d output (iresnet / ArcFace layout).
// renorm = 1 applies FaceEmbedder's extra e/(||e||+1e-8) (reference face_embedder.py:133,178-180).
__global__ void __launch_bounds__(128)
fc_finalize_kernel(const float* __restrict__ partial, int splits, int B, const float* __restrict__ bias,
                   int l2, int renorm, float* __restrict__ out_emb, float* __restrict__ out_norm,
                   __nv_bfloat16* __restrict__ out_bf16) {
  const int b = blockIdx.x;
  const int t = threadIdx.x;
  __shared__ float red[4];
  float x[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = t + 128 * j;
    float a = 0.f;
    for (int s = 0; s < splits; ++s) a += partial[(static_cast<size_t>(s) * B + b) * 512 + n];
    x[j] = a + bias[n];
  }
  auto block_norm = [&]() -> float {
    float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    __syncthreads();
    if ((t & 31) == 0) red[t >> 5] = ss;
    __syncthreads();
    return sqrtf(red[0] + red[1] + red[2] + red[3]);
  };
  float nrm = block_norm();
  if (l2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = x[j] / nrm;
  }
  if (out_norm && t == 0) out_norm[b] = nrm;
  if (renorm) {
    const float n2 = block_norm() + 1e-8f;
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = x[j] / n2;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = t + 128 * j;
    if (out_emb) out_emb[static_cast<size_t>(b) * 512 + n] = x[j];
    if (out_bf16) out_bf16[static_cast<size_t>(b) * 512 + n] = __float2bfloat16_rn(x[j]);
  }
}

// flip fusion (BASELINE config 5): rows b and B+b are the embeddings of a crop and its hflip,
// each already L2-normalised; template = mean of the two, renormalised with +1e-8
// (GalleryManager._aggregate_embeddings with n = 2: filter skipped, gallery_manager.py:297-317).
__global__ void __launch_bounds__(128)
flip_fuse_kernel(const float* __restrict__ emb2, int B, float* __restrict__ out, __nv_bfloat16* __restrict__ out_bf16) {
  const int b = blockIdx.x, t = threadIdx.x;
  __shared__ float red[4];
  float x[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = t + 128 * j;
    x[j] = (emb2[static_cast<size_t>(b) * 512 + n] + emb2[static_cast<size_t>(B + b) * 512 + n]) * 0.5f;
  }
  float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((t & 31) == 0) red[t >> 5] = ss;
  __syncthreads();
  const float n2 = sqrtf(red[0] + red[1] + red[2] + red[3]) + 1e-8f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = t + 128 * j;
    const float v = x[j] / n2;
    out[static_cast<size_t>(b) * 512 + n] = v;
    if (out_bf16) out_bf16[static_cast<size_t>(b) * 512 + n] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------ enrollment template aggregation
// GalleryManager._aggregate_embeddings + _filter_quality_embeddings (reference gallery_manager.py:104-122,297-317),
// batched over identities: one block per identity, its n embeddings are rows [seg[s], seg[s+1]) of `emb`.
//   n == 1 : template = the row itself (NOT re-normalised, :298-299)
//   n  > 2 : quality filter - gram matrix, diagonal zeroed, row mean over ALL n columns (divides by n, :110-112), keep
//            rows with mean >= min_sim, fall back to the two best rows when fewer than two survive (:117-119)
//   then mean (method 0) / median (1) / weighted mean (2: weights = gram row means of the KEPT rows incl. the
//   diagonal, normalised to sum 1) over the kept rows in their original order, and v / (||v|| + 1e-8).
// n <= kAggMaxRows.  fp32 throughout like the reference; BLAS summation order is not reproduced (1e-6 agreement).
constexpr int kAggMaxRows = 64;

__global__ void __launch_bounds__(128)
aggregate_templates_kernel(const float* __restrict__ emb, const long long* __restrict__ seg, int method, float min_sim,
                           float* __restrict__ out, int* __restrict__ out_kept) {
  __shared__ float gram[kAggMaxRows][kAggMaxRows + 1];
  __shared__ float row_score[kAggMaxRows];
  __shared__ float weight[kAggMaxRows];
  __shared__ int kept_idx[kAggMaxRows];
  __shared__ int n_kept_s;
  __shared__ float red[4];
  const int s = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const long long r0 = seg[s];
  const int n = static_cast<int>(seg[s + 1] - r0);
  const float* E = emb + r0 * 512;
  float* o = out + static_cast<size_t>(s) * 512;
  if (n <= 0) {
    for (int d = t; d < 512; d += 128) o[d] = 0.f;
    if (t == 0 && out_kept) out_kept[s] = 0;
    return;
  }
  if (n == 1) {
    for (int d = t; d < 512; d += 128) o[d] = E[d];
    if (t == 0 && out_kept) out_kept[s] = 1;
    return;
  }
  // gram matrix: one warp per (i, j >= i) pair
  for (int pidx = warp; pidx < n * n; pidx += 4) {
    const int i = pidx / n, j = pidx - i * n;
    if (j < i) continue;
    float acc = 0.f;
    for (int d = lane; d < 512; d += 32) acc = fmaf(E[i * 512 + d], E[j * 512 + d], acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      gram[i][j] = acc;
      gram[j][i] = acc;
    }
  }
  __syncthreads();
  if (t == 0) {
    int nk = 0;
    if (n <= 2) {
      for (int i = 0; i < n; ++i) kept_idx[nk++] = i;
    } else {
      for (int i = 0; i < n; ++i) {
        float sum = 0.f;
        for (int j = 0; j < n; ++j)
          if (j != i) sum += gram[i][j];
        row_score[i] = __fdiv_rn(sum, static_cast<float>(n));
        if (row_score[i] >= min_sim) kept_idx[nk++] = i;
      }
      if (nk < 2) {  // the two best rows (np.argsort(avg)[-2:]: second best first)
        int b1 = 0;
        for (int i = 1; i < n; ++i)
          if (row_score[i] >= row_score[b1]) b1 = i;
        int b2 = (b1 == 0) ? 1 : 0;
        for (int i = 0; i < n; ++i)
          if (i != b1 && row_score[i] >= row_score[b2]) b2 = i;
        kept_idx[0] = b2;
        kept_idx[1] = b1;
        nk = 2;
      }
    }
    n_kept_s = nk;
    if (method == 2) {  // weights from the gram of the kept rows (diagonal included), normalised to sum 1
      float wsum = 0.f;
      for (int a = 0; a < nk; ++a) {
        float sum = 0.f;
        for (int b = 0; b < nk; ++b) sum += gram[kept_idx[a]][kept_idx[b]];
        weight[a] = __fdiv_rn(sum, static_cast<float>(nk));
        wsum += weight[a];
      }
      for (int a = 0; a < nk; ++a) weight[a] = __fdiv_rn(weight[a], wsum);
    }
    if (out_kept) out_kept[s] = nk;
  }
  __syncthreads();
  const int nk = n_kept_s;
  float v[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int d = t + 128 * q;
    if (method == 1) {  // median over the kept rows: insertion sort of <= 64 values
      float buf[kAggMaxRows];
      for (int a = 0; a < nk; ++a) {
        const float x = E[kept_idx[a] * 512 + d];
        int b = a;
        while (b > 0 && buf[b - 1] > x) {
          buf[b] = buf[b - 1];
          --b;
        }
        buf[b] = x;
      }
      v[q] = (nk & 1) ? buf[nk >> 1] : 0.5f * (buf[(nk >> 1) - 1] + buf[nk >> 1]);
    } else if (method == 2) {
      float acc = 0.f;
      for (int a = 0; a < nk; ++a) acc += E[kept_idx[a] * 512 + d] * weight[a];
      v[q] = acc;
    } else {
      float acc = 0.f;
      for (int a = 0; a < nk; ++a) acc += E[kept_idx[a] * 512 + d];
      v[q] = __fdiv_rn(acc, static_cast<float>(nk));
    }
  }
  float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  const float nrm = __fsqrt_rn(red[0] + red[1] + red[2] + red[3]) + 1e-8f;
#pragma unroll
  for (int q = 0; q < 4; ++q) o[t + 128 * q] = __fdiv_rn(v[q], nrm);
}

// ------------------------------------------------------------------ probe / gallery prep
// q = q / (||q|| + 1e-8)  (GalleryManager.search, gallery_manager.py:195), fp32, + bf16 copy
// row_floor / counters (either may be null): the match's per-call device state is zeroed here - block b its row's
// admission floor, block 0 the two counters - instead of by memset nodes in front of the kernel.
__global__ void __launch_bounds__(128)
probe_prepare_kernel(const float* __restrict__ in, int normalize, float* __restrict__ out_f32,
                     __nv_bfloat16* __restrict__ out_bf16, unsigned* __restrict__ row_floor, int* __restrict__ counters) {
  const int b = blockIdx.x, t = threadIdx.x;
  __shared__ float red[4];
  pdl_launch_dependents();   // the filter kernel may set itself up (barriers, TMEM, tensor maps) while this runs
  if (t == 0) {
    if (row_floor) row_floor[b] = 0u;
    if (counters && b == 0) {
      counters[0] = 0;
      counters[1] = 0;
    }
  }
  float x[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) x[j] = in[static_cast<size_t>(b) * 512 + t + 128 * j];
  if (normalize) {
    float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((t & 31) == 0) red[t >> 5] = ss;
    __syncthreads();
    const float n2 = sqrtf(red[0] + red[1] + red[2] + red[3]) + 1e-8f;
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = x[j] / n2;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const size_t o = static_cast<size_t>(b) * 512 + t + 128 * j;
    out_f32[o] = x[j];
    if (out_bf16) out_bf16[o] = __float2bfloat16_rn(x[j]);
  }
}

// fp32 gallery rows -> K-blocked bf16 copy (match_sm100.cuh: gallery_box_row) + the two gallery constants of the
// filter's error bound: max_norm[0] = max row norm ||g||, max_norm[1] = max rounding distance ||g - bf16(g)|| (see
// match_finalize_kernel).  The grid covers whole 128-row groups: rows >= N are written as zeros.
__global__ void __launch_bounds__(256)
gallery_prepare_kernel(const float* __restrict__ g, long long N, __nv_bfloat16* __restrict__ gb,
                       float* __restrict__ max_norm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + warp;
  const long long group = row >> 7;
  const int r = static_cast<int>(row & 127);
  float ss = 0.f, dd = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = 4 * (lane + 32 * j);          // this lane's 4 columns
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < N) v = reinterpret_cast<const float4*>(g + row * 512)[lane + 32 * j];
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    // x - bf16(x) is exact in fp32 (the two are within a factor of two of each other)
    const float dx = v.x - __uint_as_float(o.x << 16), dy = v.y - __uint_as_float(o.x & 0xFFFF0000u);
    const float dz = v.z - __uint_as_float(o.y << 16), dw = v.w - __uint_as_float(o.y & 0xFFFF0000u);
    dd += dx * dx + dy * dy + dz * dz + dw * dw;
    const size_t box_row = static_cast<size_t>(group * 8 + (col >> 6)) * 128 + r;
    *reinterpret_cast<uint2*>(gb + box_row * 64 + (col & 63)) = o;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    dd += __shfl_xor_sync(0xffffffffu, dd, o);
  }
  if (lane == 0 && row < N) {   // both >= 0: the integer image orders like the float
    atomicMax(reinterpret_cast<int*>(max_norm), __float_as_int(sqrtf(ss)));
    atomicMax(reinterpret_cast<int*>(max_norm) + 1, __float_as_int(sqrtf(dd)));
  }
}

// ------------------------------------------------------------------ exact scoring helpers
// f64 dot of two fp32 rows of 512 by one warp: lane owns elements lane + 32*j, fixed order,
// then a fixed xor-shuffle tree -> deterministic, identical for identical rows.
__device__ __forceinline__ double warp_dot512_f64(const float* __restrict__ a, const float* __restrict__ b, int lane) {
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s = fma(static_cast<double>(a[lane + 32 * j]), static_cast<double>(b[lane + 32 * j]), s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// Same dot product with the probe's 16 values of this lane already converted to f64 (same order, same bits): the
// float -> double conversions run on a slow pipe (ncu: they, not the loads or the DFMAs, bounded the exact kernels), and
// a warp scores many gallery rows against the same probe.
__device__ __forceinline__ void probe_lane_f64(const float* __restrict__ probe, int lane, double (&pd)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) pd[j] = static_cast<double>(probe[lane + 32 * j]);
}
__device__ __forceinline__ double warp_dot512_f64_pre(const float* __restrict__ a, const double (&pd)[16], int lane) {
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s = fma(static_cast<double>(a[lane + 32 * j]), pd[j], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// canonical order: score descending, then index ascending; idx < 0 sorts last
__device__ __forceinline__ bool cand_before(double sa, long long ia, double sb, long long ib) {
  if (ia < 0) return false;
  if (ib < 0) return true;
  if (sa != sb) return sa > sb;
  return ia < ib;
}

// ---- identity-sharded galleries: per-rank top-k rows travel to the peers' merge buffers over NVLink (peer memory)
// One record of a per-rank top-k list as it is exchanged and merged: exact f64 score + global gallery id.
struct alignas(16) TopkRec {
  double score;
  long long idx;
};
constexpr int kMaxPeers = 8;
constexpr int kFlagStride = 32;   // flag words of different ranks sit 128 bytes apart

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Where the rows a kernel finishes go besides its local outputs.  world == 0: nowhere (single-GPU match).
// Every producer kernel of a sharded match (finalize, exact fix-up, plain push) adds the rows it completed to
// `done_rows`; the block that completes row number P raises this rank's flag word on every peer (epoch value), after
// which the peers' merge kernels may read the slot.  slot[g] / flag[g] are addresses inside peer g's exchange buffer
// (cudaIpc mapping; g == rank is the local buffer).
struct PeerPush {
  int world;
  int rank;
  unsigned epoch;
  int P;                       // rows of this match (all probes of all ranks)
  int* done_rows;              // local counter, zeroed at the start of the match
  TopkRec* slot[kMaxPeers];    // [P][k] records of THIS rank inside peer g's buffer (current epoch parity)
  unsigned* flag[kMaxPeers];   // this rank's result flag inside peer g's buffer
};

// n rows of this block are complete (their remote stores fenced by the storing threads, then __syncthreads)
__device__ __forceinline__ void peer_rows_done(const PeerPush& pp, int n) {
  __threadfence();
  const int prev = atomicAdd(pp.done_rows, n);
  if (prev + n == pp.P) {
    __threadfence_system();
    for (int g = 0; g < pp.world; ++g) st_release_sys(pp.flag[g], pp.epoch);
  }
}

// threads t < k of a block hold record t of `row`; all threads of the block must call this
__device__ __forceinline__ void peer_push_row(const PeerPush& pp, int row, int k, int t, double score, long long gid) {
  if (t < k) {
    TopkRec r;
    r.score = score;
    r.idx = gid;
    for (int g = 0; g < pp.world; ++g) pp.slot[g][static_cast<size_t>(row) * k + t] = r;
    __threadfence_system();
  }
  __syncthreads();
  if (t == 0) peer_rows_done(pp, 1);
}

constexpr int kRescore = 64;       // survivors re-scored exactly per probe
constexpr int kMaxCandPad = 2048;  // >= slices * kCand, power of two

struct FinalizeParams {
  const float* cand_score; const int* cand_idx;  // [P][slices][kCand]
  int slices;
  const float* probes;   // [P][512] fp32 (already normalised as search() does)
  const float* gallery;  // [N][512] fp32
  long long N;
  long long first_global_id;
  int k;
  int rescore;               // survivors re-scored exactly: 8 k clamped to [24, kRescore] (all of them only for k > 8)
  float thr;
  const float* max_norm;     // [2] gallery max row norm, max row rounding distance ||g - bf16(g)|| (device)
  double* out_score;         // [P][k] f64
  long long* out_idx;        // [P][k] global ids, -1 = none
  float* out_score_f32;      // [P][k]
  unsigned char* out_accept; // [P]  top-1 score >= thr
  int* flagged;              // [P] 1 = proof failed -> exact scan required
  int* flag_rows;            // compacted list of the flagged rows (order unspecified) ...
  int* flag_count;           // ... and its length (device counter, zeroed at the start of the match)
  PeerPush push;             // sharded match: proven rows go straight to the peers
};

// one block (128 threads) per probe
__global__ void __launch_bounds__(128)
match_finalize_kernel(const FinalizeParams p) {
  __shared__ float s_sc[kMaxCandPad];
  __shared__ int s_ix[kMaxCandPad];
  __shared__ float s_probe[512];
  __shared__ double s_ex[kRescore];
  __shared__ long long s_exi[kRescore];
  __shared__ float s_excl;
  __shared__ float s_e2[4], s_b2[4];
  __shared__ int s_flag;
  const int row = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  pdl_launch_dependents();
  pdl_wait();   // launched with programmatic stream serialization: the filter's candidate lists are complete from here
  const int C = p.slices * kCand;
  int Cp = 64;
  while (Cp < C) Cp <<= 1;
  for (int i = t; i < Cp; i += 128) {
    if (i < C) {
      s_sc[i] = p.cand_score[static_cast<size_t>(row) * C + i];
      s_ix[i] = p.cand_idx[static_cast<size_t>(row) * C + i];
    } else {
      s_sc[i] = -INFINITY;
      s_ix[i] = -1;
    }
  }
  // the probe, and the two probe terms of the error bound: ||q - bf16(q)||^2 and ||bf16(q)||^2 (the filter scored
  // with exactly this rounding of exactly these values: probe_prepare_kernel / probe_push_kernel)
  float e2 = 0.f, b2 = 0.f;
  for (int i = t; i < 512; i += 128) {
    const float q = p.probes[static_cast<size_t>(row) * 512 + i];
    const float qb = __bfloat162float(__float2bfloat16_rn(q));
    s_probe[i] = q;
    e2 += (q - qb) * (q - qb);
    b2 += qb * qb;
  }
  if (t == 0) s_excl = -INFINITY;
  __syncthreads();
  // bound on every element a slice dropped: that slice's smallest kept score (if its list is full)
  {
    float mx = -INFINITY;
    for (int s = t; s < p.slices; s += 128) {
      const int last = s * kCand + kCand - 1;
      if (s_ix[last] >= 0) mx = fmaxf(mx, s_sc[last]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      e2 += __shfl_xor_sync(0xffffffffu, e2, o);
      b2 += __shfl_xor_sync(0xffffffffu, b2, o);
    }
    __shared__ float s_mx[4];
    if (lane == 0) {
      s_mx[warp] = mx;
      s_e2[warp] = e2;
      s_b2[warp] = b2;
    }
    __syncthreads();
    if (t == 0) s_excl = fmaxf(fmaxf(s_mx[0], s_mx[1]), fmaxf(s_mx[2], s_mx[3]));
    __syncthreads();
  }
  // Only the best kRescore candidates (and the next one, for the bound) matter.  With many slices, first cut the
  // list down to its ~128 best by bisection on the ordered-integer image of the scores (32 counting passes), then
  // sort 256 entries instead of up to 2048 (the full sort was 75 us per launch at P = 256: profiles/r01c).  Massive
  // ties at the cut (more than 256 entries at or above it) keep the full sort.
  if (Cp > 256) {
    __shared__ int s_w[4];
    __shared__ int s_fill;
    unsigned keys[kMaxCandPad / 128];
    const int per = Cp / 128;
#pragma unroll
    for (int u = 0; u < kMaxCandPad / 128; ++u) {
      keys[u] = 0u;
      if (u < per) {
        const int i = t + u * 128;
        if (s_ix[i] >= 0) {
          const unsigned b = __float_as_uint(s_sc[i]);
          keys[u] = (b & 0x80000000u) ? ~b : (b | 0x80000000u);   // order-preserving; 0 = empty slot
          if (keys[u] == 0u) keys[u] = 1u;
        }
      }
    }
    auto count_ge = [&](unsigned thr_key) {
      int c = 0;
#pragma unroll
      for (int u = 0; u < kMaxCandPad / 128; ++u) c += (keys[u] >= thr_key) ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (lane == 0) s_w[warp] = c;
      __syncthreads();
      const int total = s_w[0] + s_w[1] + s_w[2] + s_w[3];
      __syncthreads();
      return total;
    };
    const int valid = count_ge(1u);
    const int want = valid < 128 ? valid : 128;
    unsigned lo = 1u, hi = 0xFFFFFFFFu;      // largest key threshold that still keeps `want` candidates
    // (with the filter's shared admission floors most slice lists hold only a few entries: when at most 256 are valid
    // they are all kept and the 32 counting passes are skipped)
    if (valid > 256)
      while (lo < hi) {
        const unsigned mid = lo + (hi - lo + 1u) / 2u;
        if (count_ge(mid) >= want) lo = mid; else hi = mid - 1u;
      }
    const int kept = valid > 256 ? count_ge(lo) : valid;
    if (want > 0 && kept <= 256) {
      __shared__ float s_sc2[256];
      __shared__ int s_ix2[256];
      if (t == 0) s_fill = 0;
      for (int i = t; i < 256; i += 128) {
        s_sc2[i] = -INFINITY;
        s_ix2[i] = -1;
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < kMaxCandPad / 128; ++u)
        if (u < per && keys[u] >= lo) {
          const int slot = atomicAdd(&s_fill, 1);
          s_sc2[slot] = s_sc[t + u * 128];
          s_ix2[slot] = s_ix[t + u * 128];
        }
      __syncthreads();
      for (int i = t; i < 256; i += 128) {
        s_sc[i] = s_sc2[i];
        s_ix[i] = s_ix2[i];
      }
      Cp = 256;
      __syncthreads();
    }
  }
  // bitonic sort of the approximate candidates, canonical order
  for (int k2 = 2; k2 <= Cp; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = t; i < Cp; i += 128) {
        const int l = i ^ j;
        if (l > i) {
          const bool up = ((i & k2) == 0);
          const bool i_first = cand_before(s_sc[i], s_ix[i], s_sc[l], s_ix[l]);
          if (up ? !i_first : i_first) {
            const float ts = s_sc[i]; s_sc[i] = s_sc[l]; s_sc[l] = ts;
            const int ti = s_ix[i]; s_ix[i] = s_ix[l]; s_ix[l] = ti;
          }
        }
      }
      __syncthreads();
    }
  }
  // Error bound of the filter's scores (derivation at the proof below); every thread evaluates the same expression.
  const float qerr = sqrtf(s_e2[0] + s_e2[1] + s_e2[2] + s_e2[3]);
  const float qbn = sqrtf(s_b2[0] + s_b2[1] + s_b2[2] + s_b2[3]);
  const float gmax = p.max_norm[0], gerr = p.max_norm[1];
  const float eps = (qerr * gmax + qbn * gerr + 0.000244140625f * qbn * (gmax + gerr)) * 1.0001f + 1e-6f;
  // exact re-score of the best R survivors (R <= p.rescore <= kRescore; slots beyond R stay empty).  Not all of
  // p.rescore need it: k candidates have an approximate score >= a_k (the k-th best), hence an exact score
  // >= a_k - eps, so a candidate below a_k - 2 eps has an exact score below a_k - eps and is not among the k best.
  // On random 125 k-row shards that leaves 10-15 of 40 rows to fetch - the re-score is a gather of 2 KB gallery rows
  // and was HBM-bound at 4096 probes (335 MB per match).
  int R = p.rescore;
  {
    const int kk = static_cast<int>(min(static_cast<long long>(p.k), p.N));
    const bool have_k = kk >= 1 && s_ix[kk - 1] >= 0;          // block-uniform
    const float cut = have_k ? s_sc[kk - 1] - 2.f * eps - 1e-6f : -INFINITY;
    const int need = __syncthreads_count(t < R && s_ix[t] >= 0 && s_sc[t] >= cut);   // the list is sorted: a prefix
    if (have_k) R = max(kk, need);
  }
  for (int c = warp; c < kRescore; c += 4) {
    const int gi = c < R ? s_ix[c] : -1;
    double sc = 0.0;
    if (gi >= 0) sc = warp_dot512_f64(p.gallery + static_cast<size_t>(gi) * 512, s_probe, lane);
    if (lane == 0) {
      s_ex[c] = sc;
      s_exi[c] = gi;
    }
  }
  __syncthreads();
  // sort the kRescore exact scores (bitonic, 64 elements, 128 threads)
  for (int k2 = 2; k2 <= kRescore; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      if (t < kRescore) {
        const int i = t, l = i ^ j;
        if (l > i) {
          const bool up = ((i & k2) == 0);
          const bool i_first = cand_before(s_ex[i], s_exi[i], s_ex[l], s_exi[l]);
          if (up ? !i_first : i_first) {
            const double ts = s_ex[i]; s_ex[i] = s_ex[l]; s_ex[l] = ts;
            const long long ti = s_exi[i]; s_exi[i] = s_exi[l]; s_exi[l] = ti;
          }
        }
      }
      __syncthreads();
    }
  }
  if (t < p.k) {
    const long long gi = s_exi[t];
    const size_t o = static_cast<size_t>(row) * p.k + t;
    p.out_idx[o] = gi >= 0 ? gi + p.first_global_id : -1;
    p.out_score[o] = gi >= 0 ? s_ex[t] : -INFINITY;
    p.out_score_f32[o] = gi >= 0 ? static_cast<float>(s_ex[t]) : -INFINITY;
  }
  if (t == 0) {
    p.out_accept[row] = (s_exi[0] >= 0 && static_cast<float>(s_ex[0]) >= p.thr) ? 1 : 0;
    // Proof that the bf16 filter dropped nothing from the true top-k:
    // anything not re-scored has approx score <= bound, hence exact score <= bound + eps, where
    //   exact - approx = (q - qb).g + qb.(g - gb) + (sum of the bf16 products - its fp32 accumulation in TMEM)
    //   |.|           <= ||q - qb|| max||g|| + ||qb|| max||g - gb|| + 2^-12 ||qb|| max||gb||        (Cauchy-Schwarz)
    // with qb = bf16(q), gb = bf16(g).  ||q - qb|| is this probe's own rounding distance and max||g - gb|| the
    // gallery's (gallery_prepare_kernel): ~0.0017 each for unit rows, so eps ~ 0.0037 where the worst-case
    // element-wise bound 2^-7 ||q|| ||g|| used until round 2 gave 0.0081 - and on a 125 k-row shard the margin
    // between the 5th exact score and the best dropped score is only 0.002-0.02 (tools/flag_diag.py): with the loose
    // eps one row in a thousand went to the exact fix-up, which a sharded match then waits for on every rank.
    const int kk = static_cast<int>(min(static_cast<long long>(p.k), p.N));
    int flag = 0;
    if (kk > 0) {
      float bound = s_excl;
      if (Cp > R && s_ix[R] >= 0) bound = fmaxf(bound, s_sc[R]);
      if (bound > -INFINITY) {
        if (s_exi[kk - 1] < 0 || s_ex[kk - 1] <= static_cast<double>(bound) + eps) flag = 1;
      }
    }
    p.flagged[row] = flag;
    s_flag = flag;
    if (flag) p.flag_rows[atomicAdd(p.flag_count, 1)] = row;   // the exact fix-up kernels take it from here
  }
  if (p.push.world > 0) {
    __syncthreads();
    if (!s_flag) {   // block-uniform
      const long long gi = t < p.k ? s_exi[t] : -1;
      peer_push_row(p.push, row, p.k, t, gi >= 0 ? s_ex[t] : -INFINITY, gi >= 0 ? gi + p.first_global_id : -1);
    }
  }
}

// Exact scan: dense f64 scores of F listed probes against the whole local gallery.
// grid.x covers gallery rows (8 warps -> 8 rows per block iteration), grid.y = probe in list.
__global__ void __launch_bounds__(256)
match_exact_scores_kernel(const float* __restrict__ gallery, long long N, const float* __restrict__ probes,
                          const int* __restrict__ rows, double* __restrict__ scores) {
  __shared__ float s_probe[512];
  const int f = blockIdx.y;
  const int prow = rows ? rows[f] : f;
  for (int i = threadIdx.x; i < 512; i += 256) s_probe[i] = probes[static_cast<size_t>(prow) * 512 + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double pd[16];
  probe_lane_f64(s_probe, lane, pd);
  for (long long g = static_cast<long long>(blockIdx.x) * 8 + warp; g < N; g += static_cast<long long>(gridDim.x) * 8) {
    const double s = warp_dot512_f64_pre(gallery + g * 512, pd, lane);
    if (lane == 0) scores[static_cast<size_t>(f) * N + g] = s;
  }
}

// top-k of a dense f64 score row by k rounds of block arg-max (canonical order); one block per row.
__global__ void __launch_bounds__(256)
match_exact_topk_kernel(const double* __restrict__ scores, long long N, const int* __restrict__ rows, int k,
                        float thr, long long first_global_id, double* __restrict__ out_score,
                        long long* __restrict__ out_idx, float* __restrict__ out_score_f32,
                        unsigned char* __restrict__ out_accept) {
  __shared__ double s_best[256];
  __shared__ long long s_besti[256];
  __shared__ double s_prev;
  __shared__ long long s_previ;
  const int f = blockIdx.x, t = threadIdx.x;
  const int prow = rows ? rows[f] : f;
  const double* sc = scores + static_cast<size_t>(f) * N;
  if (t == 0) {
    s_prev = INFINITY;
    s_previ = -1;
  }
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    const double pv = s_prev;
    const long long pi = s_previ;
    double best = -INFINITY;
    long long besti = -1;
    for (long long g = t; g < N; g += 256) {
      const double v = sc[g];
      // strictly after (pv, pi) in canonical order
      const bool after = (pi < 0) ? true : ((v < pv) || (v == pv && g > pi));
      if (after && cand_before(v, g, best, besti)) {
        best = v;
        besti = g;
      }
    }
    s_best[t] = best;
    s_besti[t] = besti;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (t < o && cand_before(s_best[t + o], s_besti[t + o], s_best[t], s_besti[t])) {
        s_best[t] = s_best[t + o];
        s_besti[t] = s_besti[t + o];
      }
      __syncthreads();
    }
    if (t == 0) {
      const size_t o = static_cast<size_t>(prow) * k + r;
      const long long gi = s_besti[0];
      out_idx[o] = gi >= 0 ? gi + first_global_id : -1;
      out_score[o] = gi >= 0 ? s_best[0] : -INFINITY;
      out_score_f32[o] = gi >= 0 ? static_cast<float>(s_best[0]) : -INFINITY;
      if (r == 0) out_accept[prow] = (gi >= 0 && static_cast<float>(s_best[0]) >= thr) ? 1 : 0;
      s_prev = s_best[0];
      s_previ = gi;
    }
    __syncthreads();
    if (s_previ < 0) {  // gallery exhausted: fill the rest
      if (t == 0)
        for (int r2 = r + 1; r2 < k; ++r2) {
          const size_t o = static_cast<size_t>(prow) * k + r2;
          out_idx[o] = -1; out_score[o] = -INFINITY; out_score_f32[o] = -INFINITY;
        }
      break;
    }
  }
}

// ---- exact fix-up of the rows whose filter proof failed, without a host round trip: both kernels are launched
// unconditionally after match_finalize_kernel and read the number of listed rows from the device counter it filled
// (grid-stride over the list, immediate exit when it is empty).  No dense [rows][N] score buffer: every block keeps the
// best k of its share of the gallery, the second kernel merges the kExactBlocks partial lists of a row.
constexpr int kExactBlocks = 74;   // gallery partitions per listed row (grid.x of the partial kernel)
constexpr int kExactRowsY = 16;    // listed rows in flight (grid.y): they share each gallery row through L2
constexpr int kExactMaxK = 32;

struct ExactFixParams {
  const float* gallery; long long N;
  const float* probes;          // [P][512] normalised probes
  const int* rows;              // listed probe rows
  const int* count;             // number of listed rows (device)
  int k; float thr; long long first_global_id;
  TopkRec* part;                // [P][kExactBlocks][k] scratch, indexed by list position (idx = local row, -1 = none)
  double* out_score; long long* out_idx; float* out_score_f32; unsigned char* out_accept;
  PeerPush push;
};

// Few listed rows (count <= gridDim.y, the usual case: zero to a handful) and k <= kExactFewK: ONE pass over the gallery
// serves all of them - the whole grid (1184 blocks) partitions the gallery, a warp loads a gallery row into registers
// once and scores it against every listed probe (all of them sit in shared memory).  A flagged row used to be scanned
// by 74 blocks on its own: one flagged row added ~0.3 ms to a sharded match of 4096 probes, four rows 0.24 ms even with
// the whole grid per row (profiles/r02_summary.md).  Larger k / more rows: one row at a time per block, rows strided
// over blockIdx.y, 74 partitions each.
constexpr int kExactFewK = 8;
constexpr int kExactFewRows = 16;  // rows the one-pass path takes (= kExactRowsY)
__device__ __forceinline__ int exact_parts(int count, int blocks_x, int rows_y, int k) {
  return (count > 0 && count <= kExactFewRows && count <= rows_y && k <= kExactFewK) ? blocks_x * rows_y : blocks_x;
}

__global__ void __launch_bounds__(256)
match_exact_part_kernel(const ExactFixParams p) {
  __shared__ float s_probe[kExactFewRows][512];                     // few-rows path: every listed probe; else row [0]
  __shared__ double w_sc[8][kExactFewRows * kExactFewK];            // per warp: few-rows path [row][k]; else [kExactMaxK] (fits: 8*8 >= 32)
  __shared__ int w_ix[8][kExactFewRows * kExactFewK];
  static_assert(kExactFewRows * kExactFewK >= kExactMaxK, "list storage");
  pdl_launch_dependents();
  pdl_wait();
  const int count = *p.count;
  if (count == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, k = p.k;
  const int rows_y = static_cast<int>(gridDim.y), blocks_x = static_cast<int>(gridDim.x);
  const int parts = exact_parts(count, blocks_x, rows_y, k);
  if (parts != blocks_x) {
    // ---- few rows: one gallery pass for all of them
    const int part = blockIdx.y * blocks_x + blockIdx.x;
    for (int i = threadIdx.x; i < count * 512; i += 256) s_probe[i >> 9][i & 511] = p.probes[static_cast<size_t>(p.rows[i >> 9]) * 512 + (i & 511)];
    for (int i = threadIdx.x; i < 8 * kExactFewRows * kExactFewK; i += 256) {
      (&w_sc[0][0])[i] = -INFINITY;
      (&w_ix[0][0])[i] = -1;
    }
    __syncthreads();
    // (ncu: 149 us for 4 flagged rows x 125 k gallery rows, bound by instruction issue - 197 instructions per (gallery
    // row, probe), half of them float -> double conversions on the XU pipe.  Variants measured and dropped,
    // tools/bench_match.py with FRB_PROBES=random: requesting the next gallery row before scoring the current one, a grid
    // resident in one wave, probes kept as f64 in shared memory: 163-172 us; an fp32 screen in front of the f64 dot:
    // 201 us - with 9472 warp-private lists each seeing ~13 rows the lists never get selective)
    for (long long g = static_cast<long long>(part) * 8 + warp; g < p.N; g += static_cast<long long>(parts) * 8) {
      float gr[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) gr[j] = p.gallery[g * 512 + lane + 32 * j];
      for (int f = 0; f < count; ++f) {
        double* ls = &w_sc[warp][f * kExactFewK];
        int* li = &w_ix[warp][f * kExactFewK];
        // exact score: same summation order as warp_dot512_f64 (the same bits on every path)
        double sdot = 0.0;
#pragma unroll
        for (int j = 0; j < 16; ++j) sdot = fma(static_cast<double>(gr[j]), static_cast<double>(s_probe[f][lane + 32 * j]), sdot);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
        if (lane == 0 && cand_before(sdot, g, ls[k - 1], li[k - 1])) {
          int j = k - 1;
          while (j > 0 && cand_before(sdot, g, ls[j - 1], li[j - 1])) {
            ls[j] = ls[j - 1];
            li[j] = li[j - 1];
            --j;
          }
          ls[j] = sdot;
          li[j] = static_cast<int>(g);
        }
      }
    }
    __syncthreads();
    for (int f = threadIdx.x; f < count; f += 256) {   // thread f: 8-way merge of row f's sorted warp lists
      int head[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      TopkRec* dst = p.part + (static_cast<size_t>(f) * parts + part) * k;
      for (int r = 0; r < k; ++r) {
        int bw = -1;
        for (int w = 0; w < 8; ++w) {
          if (head[w] >= k) continue;
          const int ci = w_ix[w][f * kExactFewK + head[w]];
          if (ci < 0) continue;
          if (bw < 0 || cand_before(w_sc[w][f * kExactFewK + head[w]], ci, w_sc[bw][f * kExactFewK + head[bw]], w_ix[bw][f * kExactFewK + head[bw]])) bw = w;
        }
        TopkRec rec;
        rec.score = bw >= 0 ? w_sc[bw][f * kExactFewK + head[bw]] : -INFINITY;
        rec.idx = bw >= 0 ? w_ix[bw][f * kExactFewK + head[bw]] : -1;
        if (bw >= 0) ++head[bw];
        dst[r] = rec;
      }
    }
    return;
  }
  // ---- many rows (or k > kExactFewK): rows strided over blockIdx.y, 74 partitions each
  const int part = blockIdx.x;
  for (int f = blockIdx.y; f < count; f += rows_y) {
    const int prow = p.rows[f];
    __syncthreads();   // the previous row's merge has read the lists
    float* probe_f = &s_probe[0][0];
    for (int i = threadIdx.x; i < 512; i += 256) probe_f[i] = p.probes[static_cast<size_t>(prow) * 512 + i];
    for (int i = threadIdx.x; i < 8 * kExactFewRows * kExactFewK; i += 256) {
      (&w_sc[0][0])[i] = -INFINITY;
      (&w_ix[0][0])[i] = -1;
    }
    __syncthreads();
    double pd[16];
    probe_lane_f64(probe_f, lane, pd);
    for (long long g = static_cast<long long>(part) * 8 + warp; g < p.N; g += static_cast<long long>(parts) * 8) {
      const double s = warp_dot512_f64_pre(p.gallery + g * 512, pd, lane);
      if (lane == 0 && cand_before(s, g, w_sc[warp][k - 1], w_ix[warp][k - 1])) {
        int j = k - 1;   // sorted insertion (canonical order); rare once the list has settled
        while (j > 0 && cand_before(s, g, w_sc[warp][j - 1], w_ix[warp][j - 1])) {
          w_sc[warp][j] = w_sc[warp][j - 1];
          w_ix[warp][j] = w_ix[warp][j - 1];
          --j;
        }
        w_sc[warp][j] = s;
        w_ix[warp][j] = static_cast<int>(g);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {   // 8-way merge of the sorted warp lists -> this block's best k
      int head[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      TopkRec* dst = p.part + (static_cast<size_t>(f) * parts + part) * k;
      for (int r = 0; r < k; ++r) {
        int bw = -1;
        for (int w = 0; w < 8; ++w)
          if (head[w] < k && w_ix[w][head[w]] >= 0 &&
              (bw < 0 || cand_before(w_sc[w][head[w]], w_ix[w][head[w]], w_sc[bw][head[bw]], w_ix[bw][head[bw]])))
            bw = w;
        TopkRec rec;
        rec.score = bw >= 0 ? w_sc[bw][head[bw]] : -INFINITY;
        rec.idx = bw >= 0 ? w_ix[bw][head[bw]] : -1;
        if (bw >= 0) ++head[bw];
        dst[r] = rec;
      }
    }
  }
}

// merge of a listed row's kExactBlocks partial lists: k rounds of block arg-max in canonical order (as
// match_exact_topk_kernel), outputs for the probe row, and - sharded - the row goes to the peers
// dynamic shared memory: the row's partial lists (f64 score + i32 local gallery row), copied once with coalesced loads -
// the k selection rounds then run out of shared memory (reading up to 1184 x k records from L2 in every round, a few
// dependent loads per thread at a time, made the merge of a single flagged row cost ~100 us)
constexpr int kExactFixMaxCand = (kExactBlocks * kExactRowsY * kExactFewK > kExactBlocks * kExactMaxK) ? kExactBlocks * kExactRowsY * kExactFewK   // parts (74 x 16) x k
                                                                                                     : kExactBlocks * kExactMaxK;
constexpr int kExactFixSmemBytes = kExactFixMaxCand * 12;

__global__ void __launch_bounds__(128)
match_exact_fix_kernel(const ExactFixParams p, int blocks_x, int rows_y) {
  extern __shared__ __align__(16) uint8_t fx_raw[];
  double* c_sc = reinterpret_cast<double*>(fx_raw);
  int* c_ix = reinterpret_cast<int*>(c_sc + kExactFixMaxCand);
  __shared__ double s_best[128];
  __shared__ long long s_besti[128];
  __shared__ double s_out[kExactMaxK];
  __shared__ long long s_outi[kExactMaxK];
  pdl_launch_dependents();
  pdl_wait();
  const int count = *p.count;
  const int t = threadIdx.x, k = p.k;
  const int parts = exact_parts(count, blocks_x, rows_y, k);   // as match_exact_part_kernel laid the lists out
  for (int f = blockIdx.x; f < count; f += gridDim.x) {
    const int prow = p.rows[f];
    const TopkRec* cand = p.part + static_cast<size_t>(f) * parts * k;
    const int C = parts * k;
    __syncthreads();
    for (int c0 = t; c0 < C; c0 += 128 * 8) {   // eight independent 16-byte loads in flight per thread
      TopkRec rec[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (c0 + u * 128 < C) rec[u] = cand[c0 + u * 128];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (c0 + u * 128 < C) {
          c_sc[c0 + u * 128] = rec[u].score;
          c_ix[c0 + u * 128] = static_cast<int>(rec[u].idx);
        }
    }
    __syncthreads();
    double pv = INFINITY;
    long long pi = -1;
    for (int r = 0; r < k; ++r) {
      double best = -INFINITY;
      long long besti = -1;
      for (int c = t; c < C; c += 128) {
        const int ci = c_ix[c];
        if (ci < 0) continue;
        const double cs = c_sc[c];
        const bool after = (pi < 0) ? true : ((cs < pv) || (cs == pv && ci > pi));
        if (after && cand_before(cs, ci, best, besti)) {
          best = cs;
          besti = ci;
        }
      }
      s_best[t] = best;
      s_besti[t] = besti;
      __syncthreads();
      for (int o = 64; o > 0; o >>= 1) {
        if (t < o && cand_before(s_best[t + o], s_besti[t + o], s_best[t], s_besti[t])) {
          s_best[t] = s_best[t + o];
          s_besti[t] = s_besti[t + o];
        }
        __syncthreads();
      }
      pv = s_best[0];
      pi = s_besti[0];
      if (t == 0) {
        s_out[r] = pi >= 0 ? pv : -INFINITY;
        s_outi[r] = pi >= 0 ? pi + p.first_global_id : -1;
      }
      __syncthreads();
      if (pi < 0) {   // gallery exhausted (block-uniform)
        if (t == 0)
          for (int r2 = r + 1; r2 < k; ++r2) {
            s_out[r2] = -INFINITY;
            s_outi[r2] = -1;
          }
        break;
      }
    }
    __syncthreads();
    if (t < k) {
      const size_t o = static_cast<size_t>(prow) * k + t;
      p.out_idx[o] = s_outi[t];
      p.out_score[o] = s_out[t];
      p.out_score_f32[o] = static_cast<float>(s_out[t]);
    }
    if (t == 0) p.out_accept[prow] = (s_outi[0] >= 0 && static_cast<float>(s_out[0]) >= p.thr) ? 1 : 0;
    if (p.push.world > 0) peer_push_row(p.push, prow, k, t, t < k ? s_out[t] : 0.0, t < k ? s_outi[t] : -1);
    __syncthreads();
  }
}

// constant rows for an empty gallery (idx -1, score -inf, accept 0)
__global__ void match_fill_empty_kernel(int P, int k, double* __restrict__ out_score, long long* __restrict__ out_idx,
                                        float* __restrict__ out_score_f32, unsigned char* __restrict__ out_accept) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P * k) {
    out_idx[i] = -1;
    if (out_score) out_score[i] = -INFINITY;
    out_score_f32[i] = -INFINITY;
  }
  if (i < P) out_accept[i] = 0;
}

// ------------------------------------------------------------------ per-identity matching over gallery SAMPLES
// SURVEY §8f row 1 (reference evaluate_models_v2.ipynb cells 3-5: compute_all_similarities, aggregate_max / mean /
// topk, identify_probe): the gallery holds every enrolled sample; identity i owns rows [seg[i], seg[i+1]) and its score
// for a probe is max / mean / mean-of-top-k of the probe's similarities to those rows; identities are ranked by
// (score desc, identity index asc) - the order Python's stable sort gives the notebook.
enum IdAgg : int { ID_AGG_MAX = 0, ID_AGG_MEAN = 1, ID_AGG_TOPK = 2 };

__device__ __forceinline__ double identity_aggregate(double* sc, int n, int agg, int agg_k) {
  if (n <= 0) return -1.0;                       // aggregate_* return -1 for an identity without samples
  if (agg == ID_AGG_MEAN) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += sc[i];
    return s / n;
  }
  if (agg == ID_AGG_TOPK) {
    const int kk = agg_k < n ? agg_k : n;       // np.mean(sorted(sims, reverse=True)[:min(k, len)])
    double s = 0.0;
    for (int r = 0; r < kk; ++r) {               // selection of the kk largest, in place
      int b = r;
      for (int i = r + 1; i < n; ++i)
        if (sc[i] > sc[b]) b = i;
      const double t = sc[r]; sc[r] = sc[b]; sc[b] = t;
      s += sc[r];
    }
    return s / kk;
  }
  double m = sc[0];
  for (int i = 1; i < n; ++i) m = sc[i] > m ? sc[i] : m;
  return m;
}

// dense exact sample scores [F][T] (f64) -> identity scores [F][S] (f64 and/or f32); one thread per (probe, identity)
__global__ void identity_reduce_kernel(const double* __restrict__ sample_scores, long long T, const long long* __restrict__ seg,
                                       long long S, int agg, int agg_k, double* __restrict__ out64, float* __restrict__ out32) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int f = blockIdx.y;
  if (i >= S) return;
  const long long r0 = seg[i];
  const int n = static_cast<int>(seg[i + 1] - r0);
  double buf[kAggMaxRows];
  for (int j = 0; j < n; ++j) buf[j] = sample_scores[static_cast<size_t>(f) * T + r0 + j];
  const double v = identity_aggregate(buf, n, agg, agg_k);
  if (out64) out64[static_cast<size_t>(f) * S + i] = v;
  if (out32) out32[static_cast<size_t>(f) * S + i] = static_cast<float>(v);
}

struct IdentityCandParams {
  const float* probes;          // [P][512] normalised probes
  const float* gallery;         // [T][512] fp32 samples
  const long long* seg;         // [S+1]
  const int* sample_identity;   // [T]
  const long long* top_idx;     // [P][KS] exact top-KS sample rows (canonical order), -1 = none
  const double* top_sc;         // [P][KS]
  int KS;
  long long T, S;
  int agg, agg_k, k;
  float thr;
  double* out_score; long long* out_idx; float* out_score_f32; unsigned char* out_accept;
  int* flagged;                 // 1 = the candidate set could not be proven sufficient -> exact scan
};

// One block (128 threads) per probe.  Candidates = the distinct identities among the probe's exact top-KS samples; each
// candidate's aggregate is computed exactly (f64) from ALL its samples.  Any identity outside the candidate set has
// every sample score <= the KS-th sample score, hence aggregate <= that bound (mean <= max): the top-k is proven when
// the k-th candidate aggregate beats the bound strictly.
__global__ void __launch_bounds__(128)
identity_candidates_kernel(const IdentityCandParams p) {
  __shared__ float s_probe[512];
  __shared__ int s_cand[64];
  __shared__ double s_agg[64];
  __shared__ double s_sc[kAggMaxRows];
  __shared__ int s_ncand;
  const int f = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  for (int i = t; i < 512; i += 128) s_probe[i] = p.probes[static_cast<size_t>(f) * 512 + i];
  if (t == 0) {
    int nc = 0;
    for (int j = 0; j < p.KS; ++j) {
      const long long r = p.top_idx[static_cast<size_t>(f) * p.KS + j];
      if (r < 0) break;
      const int id = p.sample_identity[r];
      bool seen = false;
      for (int c = 0; c < nc; ++c) seen |= (s_cand[c] == id);
      if (!seen) s_cand[nc++] = id;
    }
    s_ncand = nc;
  }
  __syncthreads();
  const int nc = s_ncand;
  for (int c = 0; c < nc; ++c) {
    const long long r0 = p.seg[s_cand[c]];
    const int n = static_cast<int>(p.seg[s_cand[c] + 1] - r0);
    for (int j = warp; j < n; j += 4) {
      const double d = warp_dot512_f64(p.gallery + (r0 + j) * 512, s_probe, lane);
      if (lane == 0) s_sc[j] = d;
    }
    __syncthreads();
    if (t == 0) s_agg[c] = identity_aggregate(s_sc, n, p.agg, p.agg_k);
    __syncthreads();
  }
  if (t == 0) {
    const long long last = p.top_idx[static_cast<size_t>(f) * p.KS + p.KS - 1];
    const bool saw_everything = last < 0;           // fewer than KS samples exist: all identities with samples are candidates
    const double bound = saw_everything ? -INFINITY : p.top_sc[static_cast<size_t>(f) * p.KS + p.KS - 1];
    double ps = INFINITY;
    long long pi = -1;
    bool proven = true;
    for (int r = 0; r < p.k; ++r) {
      double best = -INFINITY;
      long long besti = -1;
      for (int c = 0; c < nc; ++c) {
        const double v = s_agg[c];
        const long long ix = s_cand[c];
        const bool after = (pi < 0) ? true : ((v < ps) || (v == ps && ix > pi));
        if (after && cand_before(v, ix, best, besti)) {
          best = v;
          besti = ix;
        }
      }
      const size_t o = static_cast<size_t>(f) * p.k + r;
      p.out_idx[o] = besti;
      p.out_score[o] = besti >= 0 ? best : -INFINITY;
      p.out_score_f32[o] = besti >= 0 ? static_cast<float>(best) : -INFINITY;
      if (besti < 0) proven = false;                        // fewer candidates than k: let the exact scan decide
      else if (!(best > bound)) proven = proven && saw_everything;
      if (besti >= 0) {
        ps = best;
        pi = besti;
      }
    }
    p.out_accept[f] = (p.out_idx[static_cast<size_t>(f) * p.k] >= 0 &&
                       p.out_score_f32[static_cast<size_t>(f) * p.k] >= p.thr) ? 1 : 0;
    p.flagged[f] = proven ? 0 : 1;
  }
}

// ------------------------------------------------------------------ track-level consensus
// FaceMatcher._aggregate_matches + _get_best_candidate (face_matcher.py:321-385) for T tracks at once.  Track t owns
// frames [seg[t], seg[t+1]); a frame is its top-1 (gallery row, f32 score) from frb_match, row < 0 = the frame had no
// match and is skipped (face_matcher.py:180-181).  Votes are counted the way collections.Counter.most_common orders
// them (count descending, first occurrence first); means follow numpy's pairwise float64 summation bit for bit.
struct frb_track_result_dev {      // == frb_track_result (include/frb200.h)
  long long winner;                // consensus identity (gallery row) or -1
  double confidence;               // mean of the winner's quality-frame scores
  double consensus_strength;       // winner votes / quality frames
  int num_quality_frames;          // winner's quality frames
  int total_frames_evaluated;      // frames that had a match
  long long candidate;             // _get_best_candidate identity (-1 when the track has no matched frame)
  double candidate_confidence;
  int candidate_num_quality_frames;
  int recognized;                  // 1 = consensus reached and confidence >= threshold
};

// numpy pairwise_sum of float64 (numpy/_core/src/umath/loops_utils.h.src): < 8 sequential, <= 128 eight
// interleaved accumulators, above that recursive halves (multiples of 8)
__device__ inline double np_pairwise_sum(const double* a, int n) {
  if (n < 8) {
    double res = 0.;
    for (int i = 0; i < n; ++i) res += a[i];
    return res;
  }
  if (n <= 128) {
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

constexpr int kTrackMaxFrames = 1024;

// one block per track; the vote itself is a short sequential scan (tracks hold tens of frames), done by thread 0
// on shared-memory copies the whole block loads
__global__ void __launch_bounds__(128)
track_consensus_kernel(const long long* __restrict__ top_idx, const float* __restrict__ top_score, int stride,
                       const long long* __restrict__ seg, double min_quality, int min_frames, double thr,
                       frb_track_result_dev* __restrict__ out) {
  __shared__ long long s_id[kTrackMaxFrames];
  __shared__ double s_sc[kTrackMaxFrames];
  __shared__ double s_tmp[kTrackMaxFrames];
  const int t = blockIdx.x;
  const long long f0 = seg[t];
  int F = static_cast<int>(seg[t + 1] - f0);
  if (F > kTrackMaxFrames) F = kTrackMaxFrames;   // the entry point rejects longer tracks; never overrun shared memory
  for (int i = threadIdx.x; i < F; i += blockDim.x) {
    s_id[i] = top_idx[(f0 + i) * stride];
    s_sc[i] = static_cast<double>(top_score[(f0 + i) * stride]);
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  frb_track_result_dev r;
  r.winner = -1; r.confidence = 0.; r.consensus_strength = 0.; r.num_quality_frames = 0; r.total_frames_evaluated = 0;
  r.candidate = -1; r.candidate_confidence = 0.; r.candidate_num_quality_frames = 0; r.recognized = 0;
  int total = 0, voters = 0;
  for (int i = 0; i < F; ++i) {
    if (s_id[i] < 0) continue;
    ++total;
    if (s_sc[i] >= min_quality) ++voters;
  }
  r.total_frames_evaluated = total;
  // Counter.most_common over a pool (quality frames, or every matched frame): best and runner-up by
  // (count desc, first occurrence asc)
  auto rank2 = [&](bool quality_only, long long* id1, int* c1, int* c2) {
    *id1 = -1; *c1 = 0; *c2 = 0;
    long long id2 = -1;
    for (int i = 0; i < F; ++i) {
      const long long id = s_id[i];
      if (id < 0 || (quality_only && !(s_sc[i] >= min_quality))) continue;
      bool first = true;
      for (int j = 0; j < i && first; ++j)
        if (s_id[j] == id && !(quality_only && !(s_sc[j] >= min_quality))) first = false;
      if (!first) continue;
      int c = 0;
      for (int j = i; j < F; ++j)
        if (s_id[j] == id && !(quality_only && !(s_sc[j] >= min_quality))) ++c;
      if (c > *c1) { id2 = *id1; *c2 = *c1; *id1 = id; *c1 = c; }
      else if (c > *c2) { id2 = id; *c2 = c; }
    }
    (void)id2;
  };
  auto mean_of = [&](long long id, bool quality_only, int* n_out) {
    int n = 0;
    for (int i = 0; i < F; ++i)
      if (s_id[i] == id && !(quality_only && !(s_sc[i] >= min_quality))) s_tmp[n++] = s_sc[i];
    *n_out = n;
    return np_pairwise_sum(s_tmp, n) / static_cast<double>(n);
  };
  if (total > 0) {
    const bool pool_quality = voters > 0;            // _get_best_candidate: quality frames, else all frames
    long long cid; int cc1, cc2;
    rank2(pool_quality, &cid, &cc1, &cc2);
    r.candidate = cid;
    r.candidate_confidence = mean_of(cid, pool_quality, &r.candidate_num_quality_frames);
  }
  if (voters >= min_frames) {
    long long wid; int c1, c2;
    rank2(true, &wid, &c1, &c2);
    const double share = static_cast<double>(c1) / static_cast<double>(voters);
    bool agreed = share > 0.5;
    if (!agreed && c2 > 0) agreed = share > 0.4 && c1 >= 2 * c2;
    if (agreed) {
      int n;
      const double m = mean_of(wid, true, &n);
      if (!(m < thr)) {
        r.winner = wid; r.confidence = m; r.consensus_strength = share; r.num_quality_frames = n; r.recognized = 1;
      }
    }
  }
  out[t] = r;
}

// ------------------------------------------------------------------ server best-frame selection
// LiveRecognitionTracker.get_best_frame + the should_recognize gate (face_recognition_server.py:39-85) for T tracks:
// quality = det * min(blur / 100, 1) in f64 as Python evaluates it; Python's max() keeps the FIRST maximal frame;
// ready = best frame's det_score > min_det.  One warp per track.
__global__ void __launch_bounds__(128)
best_frames_kernel(const double* __restrict__ det, const double* __restrict__ blur, const long long* __restrict__ seg,
                   int T, double min_det, long long* __restrict__ out_idx, double* __restrict__ out_quality,
                   unsigned char* __restrict__ out_ready) {
  const int t = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (t >= T) return;
  const long long f0 = seg[t], f1 = seg[t + 1];
  double best = 0.0;
  long long besti = -1;
  for (long long f = f0 + lane; f < f1; f += 32) {
    const double q = __dmul_rn(det[f], fmin(__ddiv_rn(blur[f], 100.0), 1.0));
    if (besti < 0 || q > best) {   // strictly greater: the earliest frame of this lane's stride wins a tie
      best = q;
      besti = f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (oi >= 0 && (besti < 0 || ob > best || (ob == best && oi < besti))) {
      best = ob;
      besti = oi;
    }
  }
  if (lane == 0) {
    out_idx[t] = besti >= 0 ? besti - f0 : -1;
    if (out_quality) out_quality[t] = besti >= 0 ? best : 0.0;
    out_ready[t] = (besti >= 0 && det[besti] > min_det) ? 1 : 0;
  }
}

// ------------------------------------------------------------------ identity-sharded gallery: exchange + merge
// (new work, no reference counterpart: the reference is single-device)
struct XchgWait {
  int world;                       // 0 = nothing to wait for
  unsigned epoch;
  const unsigned* flag;            // [world] flag words (kFlagStride apart) in the LOCAL exchange buffer, raised by the peers
  unsigned long long timeout_ns;
  int* status;                     // pinned host word: set to rank+1 of the first missing peer before trapping
};

// lanes 0..world-1 of the calling warp poll one peer flag each (acquire at system scope)
__device__ __forceinline__ void xchg_wait_flags(const XchgWait& w, int lane) {
  if (lane < w.world) {
    const unsigned long long t0 = global_timer_ns();
    while (static_cast<int>(ld_acquire_sys(w.flag + lane * kFlagStride) - w.epoch) < 0) {
      __nanosleep(200);
      if (global_timer_ns() - t0 > w.timeout_ns) {   // a peer never arrived: fail loudly instead of hanging the GPU
        if (w.status) *reinterpret_cast<volatile int*>(w.status) = lane + 1;
        __threadfence_system();
        __trap();
      }
    }
  }
  __syncwarp();
}

// one warp: returns when every peer's probes of this epoch have landed in the local buffer; the kernels launched
// after it (filter: TMA loads) then read them
__global__ void xchg_wait_kernel(const XchgWait w) { xchg_wait_flags(w, threadIdx.x); }

// q = q / (||q|| + 1e-8) as probe_prepare_kernel, written to row (row0 + b) of EVERY rank's probe buffers
struct ProbePush {
  int world;
  int row0;                        // first global probe row of this rank
  int rows;                        // local rows (== gridDim.x)
  unsigned epoch;
  int* done_rows;
  float* f32[kMaxPeers];           // [P][512] inside peer g's exchange buffer
  __nv_bfloat16* bf16[kMaxPeers];
  unsigned* flag[kMaxPeers];       // this rank's probe flag inside peer g's buffer
};

__global__ void __launch_bounds__(128)
probe_push_kernel(const float* __restrict__ in, int normalize, const ProbePush pp) {
  const int b = blockIdx.x, t = threadIdx.x;
  __shared__ float red[4];
  // same element-to-thread mapping and summation order as probe_prepare_kernel: a probe normalises to the same bits
  // whether it is matched against the whole gallery or a shard
  float x[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) x[j] = in[static_cast<size_t>(b) * 512 + t + 128 * j];
  if (normalize) {
    float ss = x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((t & 31) == 0) red[t >> 5] = ss;
    __syncthreads();
    const float n2 = sqrtf(red[0] + red[1] + red[2] + red[3]) + 1e-8f;
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = x[j] / n2;
  }
  for (int g = 0; g < pp.world; ++g) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const size_t o = static_cast<size_t>(pp.row0 + b) * 512 + t + 128 * j;
      pp.f32[g][o] = x[j];
      pp.bf16[g][o] = __float2bfloat16_rn(x[j]);
    }
  }
  __threadfence_system();
  __syncthreads();
  if (t == 0) {
    __threadfence();
    if (atomicAdd(pp.done_rows, 1) + 1 == pp.rows) {
      __threadfence_system();
      for (int g = 0; g < pp.world; ++g) st_release_sys(pp.flag[g], pp.epoch);
    }
  }
}
// a rank without probes of its own still has to raise its flags
__global__ void probe_push_empty_kernel(const ProbePush pp) {
  if (threadIdx.x == 0)
    for (int g = 0; g < pp.world; ++g) st_release_sys(pp.flag[g], pp.epoch);
}

// local [P][k] results (dense exact path, empty shard) -> the peers
__global__ void __launch_bounds__(128)
xchg_push_rows_kernel(const double* __restrict__ score, const long long* __restrict__ idx, int P, int k, const PeerPush pp) {
  const int row0 = blockIdx.x * 32;
  const int nrows = min(32, P - row0);
  for (int i = threadIdx.x; i < nrows * k; i += blockDim.x) {
    TopkRec r;
    r.score = score[static_cast<size_t>(row0) * k + i];
    r.idx = idx[static_cast<size_t>(row0) * k + i];
    for (int g = 0; g < pp.world; ++g) pp.slot[g][static_cast<size_t>(row0) * k + i] = r;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) peer_rows_done(pp, nrows);
}

// merge G per-rank top-k lists into the global top-k in canonical order; one thread per probe.  in: [G][P][k]
// records.  With wait.world > 0 the lists are the slots of the local exchange buffer and the first warp of every
// block first waits for the peers' result flags.
__global__ void __launch_bounds__(128)
topk_merge_kernel(const TopkRec* __restrict__ in, int G, int P, int k, float thr, double* __restrict__ out_score,
                  long long* __restrict__ out_idx, float* __restrict__ out_score_f32,
                  unsigned char* __restrict__ out_accept, const XchgWait wait) {
  if (wait.world > 0) {
    if (threadIdx.x < 32) xchg_wait_flags(wait, threadIdx.x);
    __syncthreads();
  }
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= P) return;
  double ps = INFINITY;
  long long pi = -1;
  for (int r = 0; r < k; ++r) {
    double best = -INFINITY;
    long long besti = -1;
    for (int g = 0; g < G; ++g)
      for (int j = 0; j < k; ++j) {
        const TopkRec rec = in[(static_cast<size_t>(g) * P + row) * k + j];
        if (rec.idx < 0) continue;
        const bool after = (pi < 0) ? true : ((rec.score < ps) || (rec.score == ps && rec.idx > pi));
        if (after && cand_before(rec.score, rec.idx, best, besti)) {
          best = rec.score;
          besti = rec.idx;
        }
      }
    const size_t o = static_cast<size_t>(row) * k + r;
    out_idx[o] = besti;
    if (out_score) out_score[o] = besti >= 0 ? best : -INFINITY;
    out_score_f32[o] = besti >= 0 ? static_cast<float>(best) : -INFINITY;
    if (r == 0) out_accept[row] = (besti >= 0 && static_cast<float>(best) >= thr) ? 1 : 0;
    if (besti < 0) {
      for (int r2 = r + 1; r2 < k; ++r2) {
        const size_t o2 = static_cast<size_t>(row) * k + r2;
        out_idx[o2] = -1;
        if (out_score) out_score[o2] = -INFINITY;
        out_score_f32[o2] = -INFINITY;
      }
      break;
    }
    ps = best;
    pi = besti;
  }
}

// separate score / id arrays (frb_topk_merge's layout) -> records
__global__ void topk_pack_kernel(const double* __restrict__ score, const long long* __restrict__ idx, size_t n,
                                 TopkRec* __restrict__ out) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < n) {
    TopkRec r;
    r.score = score[i];
    r.idx = idx[i];
    out[i] = r;
  }
}

}  // namespace frb
