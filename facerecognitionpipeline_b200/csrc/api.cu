// api.cu — C ABI (include/frb200.h) over the sm_100a kernels: context, TMA descriptor
// construction, the backbone layer program, gallery residency and the match pipeline.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/frb200.h"
#include "gemm_sm100.cuh"
#include "gemm2_sm100.cuh"
#include "gemm2_multi_sm100.cuh"
#include "conv_slab_sm100.cuh"
#include "conv_slab_multi_sm100.cuh"
#include "match_sm100.cuh"
#include "stem_sm100.cuh"
#include "simple_kernels.cuh"

using namespace frb;

static_assert(sizeof(frb_warp_job) == sizeof(WarpJob), "warp job ABI mismatch");

namespace {

struct Plan {  // per-batch-size launch plan of the backbone
  int B = 0;
  bool dataflow = false;
  const void* d_in = nullptr;  // the descriptors bake in the batch size and the input pointer
  std::vector<CUtensorMap> tmA, tmA2, tmB;
  std::vector<CUtensorMap> tmR;    // slab layers with sp.res_tma: the shortcut as a TMA tile source
  std::vector<GemmParams> gp;
  std::vector<int> block_n, grid;
  size_t tail_flags = 0;           // split-K flags of all layers (zeroed at the start of every embed)
  // persistent runs (gemm2_multi_sm100_kernel): run_len[i] > 1 = layers i .. i+run_len[i]-1 go out as ONE launch
  std::vector<int> run_len, run_grid, run_index;
  int num_runs = 0;
  size_t run_layers = 0;           // layers in all runs = grid-barrier counters (zeroed at the start of every embed)
  std::vector<size_t> run_off;     // offset of the run's first Gemm2Layer in ctx->d_runs
  std::vector<int> srun_len;       // same for runs of identical slab layers (conv_slab_multi_sm100_kernel)
  std::vector<size_t> srun_off;    // offset of the run's first SlabLayer in ctx->d_sruns
  std::vector<int> use_slab;       // 1 = conv_slab_sm100_kernel (3x3 stride-1, W in {28,56,112})
  std::vector<SlabParams> sp;
  std::vector<int> slab_smem;
  std::vector<int> wait_target;    // dataflow: progress units every image must have before this layer may read it
};

// The FRONT of the network (stem + the 112x112 / 56x56 layers, Cout = 64) run in SUB-BATCHES: for `sub` images at a time
// the front's layers are launched back to back, so the 1.6 MB-per-image activations a layer writes are still in the
// 126 MB L2 when the next layer reads them (at 256 images per pass every one of these layers streams 200-800 MB through
// HBM and runs at about half its speed, profiles/r01e).  Only the front's last output (56x56x64) is kept for the
// whole batch.  Same kernels, same per-tile instruction streams: no bit changes.
struct FrontPlan {
  int B = 0, sub = 0, n_sub = 0, nf = 0;   // batch, images per sub-batch, sub-batches, layers [0, nf)
  std::vector<int> c0;                      // first image of every sub-batch (the last one is right-aligned)
  std::vector<int> use_slab, block_n, grid, slab_smem;   // per layer
  std::vector<std::vector<CUtensorMap>> tmA, tmA2, tmB;  // [sub-batch][layer]
  std::vector<std::vector<GemmParams>> gp;
  std::vector<std::vector<SlabParams>> sp;
};

}  // namespace

struct frb_ctx {
  int device = 0;
  int num_sms = 148;
  std::string err;
  std::mutex mu;
  long long launches = 0;
  PFN_cuTensorMapEncodeTiled_v12000 encode_tiled = nullptr;
  PFN_cuTensorMapEncodeIm2col_v12000 encode_im2col = nullptr;
  int driver_version = 0;
  int conv_mode = 2;  // 2 = CTA-pair kernel, 1 = 1-CTA kernel with weight multicast
  int tail_split = 0;  // FRB_TAIL_SPLIT=1: split-K for the last partial round of the pair conv kernel (see TileItem)
  float* d_tail_partial = nullptr; size_t tail_partial_cap = 0;
  int* d_tail_flags = nullptr; size_t tail_flags_cap = 0;
  int slab_multi = 0;  // FRB_SLAB_MULTI=1: runs of identical slab layers as ONE persistent launch as well.  Bit-identical,
                       // but measured 0.35 ms slower per batch-256 embed: the resident weights (147 KB per CTA) must
                       // be swapped at every layer boundary with the tensor pipe idle, while separate launches fetch
                       // the next layer's weights before their dependency wait, under the previous layer's tail.
  int multi_coop = 1;  // FRB_MULTI_COOP=0: no cooperative-launch attribute on the persistent run kernel
  int conv_multi = 2;  // persistent multi-layer runs of the pair conv kernel: 2 = per-image dataflow between the run's
                       // layers (default), 1 = grid barrier between them, 0 = every conv layer is a launch of its own
  Gemm2Layer* d_runs = nullptr; size_t runs_cap = 0;
  SlabLayer* d_sruns = nullptr; size_t sruns_cap = 0;
  int* d_run_bar = nullptr; size_t run_bar_cap = 0;
  int conv_quad = 0;  // FRB_QUAD: gemm2_sm100_kernel<., 4> for Cout >= 256 (1) / >= 128 (2) layers
  int quad_clusters = 0;
  int use_slab = 1;   // activation-slab kernel for eligible 3x3 stride-1 layers (FRB_SLAB=0 disables)
  int slab_stage_out = 1;   // Cout = 64 slab layers: output staged in shared memory + one TMA store per tile (FRB_SLAB_STAGE=0: direct stores)
  int slab_res_tma = 1;     // shortcut tile of the staged Cout = 64 layers by TMA into the staging buffer (FRB_SLAB_RES_TMA=0: per-thread loads)
  int embed_chunk = 256;  // faces per pass of the layer program (FRB_EMBED_CHUNK; 0 = whole batch in one pass)
  int use_pdl = 1;    // programmatic dependent launch between backbone kernels (FRB_PDL=0 disables)
  int match_pair = 1;   // CTA-pair match filter for P > 128 (FRB_MATCH_PAIR=0 disables)
  int use_dataflow = 0;  // per-image progress counters instead of whole-grid dependencies (FRB_DATAFLOW=0 disables; needs PDL)
  int* d_progress = nullptr;
  int progress_cap = 0;

  // constants
  unsigned short* d_lut = nullptr;  // 256 bf16
  unsigned short* d_wtab = nullptr;  // 32*32*4 uint16 (values 0..32768)

  // backbone
  std::vector<frb_layer_desc> layers;
  uint8_t* d_blob = nullptr;
  size_t blob_bytes = 0;
  int n_bufs = 0;
  std::vector<size_t> buf_elems_per_face;  // bf16 elements per face per buffer
  std::vector<__nv_bfloat16*> d_bufs;
  int bufs_capacity_B = 0;
  float* d_fc_partial = nullptr;
  size_t fc_partial_elems = 0;
  float* d_emb2 = nullptr;  // flip fusion scratch [2B][512]
  size_t emb2_elems = 0;
  Plan plan;
  FrontPlan front;
  int front_sub = 0;    // images per front sub-batch (FRB_FRONT_SUB=32 ...; 0 = the front runs with the whole pass like the rest).
                        // MEASURED slower on (embed 5.70-5.86 ms at 32, 5.53 at 64, 5.96 at 16 vs 5.37-5.44 ms off): the
                        // per-launch fixed costs of 8 x 7 small launches outweigh the L2 residency they buy.
  std::vector<int> prof_marks;   // frb_embed_profile: the layer every recorded event precedes
  double flops_per_face = 0.0;
  bool profiling = false;               // frb_embed_profile: CUDA events around every layer launch
  std::vector<cudaEvent_t> prof_events;

  // gallery
  float* d_gal = nullptr;
  __nv_bfloat16* d_gal_bf16 = nullptr;
  long long gal_N = 0, gal_first = 0, gal_cap = 0;
  float* d_gal_maxnorm = nullptr;
  CUtensorMap tmG;
  CUtensorMap tmG2;  // same gallery, 128-row box (CTA-pair match kernel)
  // sample gallery (per-identity matching): identity i owns gallery rows [seg[i], seg[i+1])
  long long* d_seg = nullptr; int* d_sample_identity = nullptr; long long gal_S = 0; size_t seg_cap = 0, sid_cap = 0;
  long long* d_id_top_idx = nullptr; double* d_id_top_sc = nullptr; unsigned char* d_id_acc = nullptr; float* d_id_sc32 = nullptr;
  size_t id_top_cap = 0, id_top_sc_cap = 0, id_acc_cap = 0, id_sc32_cap = 0;
  double* d_id_scores = nullptr; size_t id_scores_elems = 0;

  // match workspace
  float* d_probe_f32 = nullptr;
  __nv_bfloat16* d_probe_bf16 = nullptr;
  float* d_cand_score = nullptr;
  int* d_cand_idx = nullptr;
  int* d_flagged = nullptr;
  int* d_flag_rows = nullptr;
  unsigned* d_row_floor = nullptr;   // [match_cap_P] shared admission floors of the filter (zeroed per match)
  double* d_exact = nullptr;
  size_t exact_elems = 0;
  int match_cap_P = 0, match_cap_slices = 0;
  std::vector<int> h_flagged;
  int last_flagged = 0;
  // match without a host round trip: device counters (0 = flagged rows, 1 = rows pushed to the peers, 2 = probe rows
  // pushed), the flagged count mirrored into pinned host memory behind an event, the fix-up's partial lists
  int* d_match_ctr = nullptr;
  int* h_flag_count = nullptr;            // pinned
  cudaEvent_t flag_event = nullptr;
  bool flag_event_pending = false;
  TopkRec* d_exact_part = nullptr; size_t exact_part_cap = 0;
  TopkRec* d_merge_rec = nullptr; size_t merge_rec_cap = 0;
  cudaEvent_t match_prof_ev[4] = {nullptr, nullptr, nullptr, nullptr};   // frb_match_profile: prepare | filter | finalize + fix-up
  bool match_profiling = false;
  // generations: bumped by every upload, so a wrapper can tell whether what it uploaded is still resident
  long long gallery_gen = 0, backbone_gen = 0;
  // the workspaces above are shared by every call on this ctx: a call on another stream than the previous one first
  // waits for the previous call's work
  cudaEvent_t ws_event = nullptr;
  cudaStream_t ws_stream = nullptr;
  bool ws_valid = false;
  // identity-sharded exchange over peer memory (frb_xchg_*)
  struct Xchg {
    int world = 0, rank = 0, max_probes = 0, max_k = 0;
    uint8_t* local = nullptr;               // this rank's exchange buffer (cudaMalloc, exported through cudaIpc)
    uint8_t* peer[kMaxPeers] = {};          // every rank's buffer as mapped here (peer[rank] == local)
    bool peer_ipc[kMaxPeers] = {};          // mapped with cudaIpcOpenMemHandle (to be closed)
    size_t bytes = 0, off_f32 = 0, off_bf16 = 0, off_slots = 0, off_pflag = 0, off_rflag = 0;
    unsigned epoch = 0;
    int* h_status = nullptr;                // pinned: set by a wait that timed out
    unsigned long long timeout_ns = 20ull * 1000 * 1000 * 1000;
    bool connected = false;
  } xchg;
  double* d_scores64_tmp = nullptr;
  size_t scores64_tmp_elems = 0;

  // host-API staging
  uint8_t* d_stage_u8 = nullptr; size_t stage_u8_bytes = 0;
  // frb_prefetch_host: crops of LATER calls travel to these buffers on copy_stream while the current call computes
  struct PrefetchSlot {
    uint8_t* d_buf = nullptr; size_t cap = 0;
    const uint8_t* h_ptr = nullptr; size_t bytes = 0;   // h_ptr != nullptr: an unconsumed prefetch
    cudaEvent_t done = nullptr;
    unsigned long long seq = 0;
  } prefetch[2];
  unsigned long long prefetch_seq = 0;
  cudaStream_t copy_stream = nullptr;
  __nv_bfloat16* d_stage_in = nullptr; size_t stage_in_elems = 0;
  float* d_stage_emb = nullptr; float* d_stage_norm = nullptr; size_t stage_emb_rows = 0;
  float* d_stage_sc = nullptr; long long* d_stage_idx = nullptr; unsigned char* d_stage_acc = nullptr;
  size_t stage_match_rows = 0; int stage_match_k = 0;
  WarpJob* d_jobs = nullptr; int jobs_cap = 0;
  std::vector<int> h_boxes;      // per-face source boxes of the last frb_warp_normalize call
  int slab_min_w = 0;            // FRB_SLAB_MINW=56: 28-pixel layers use the im2col pair kernel instead of the slab kernel
  int match_prefetch = 0;        // FRB_MATCH_PREFETCH=n: L2-prefetch gallery tiles n ahead of the TMA ring (pair kernel)
  int warp_band = 0;             // FRB_WARP_BAND=1: band-staged kernel (every band's source spans copied to shared memory with
                                 // 16-byte loads first; bit-identical).  MEASURED 3x SLOWER than the direct gather (2.28 vs 0.75 ms for
                                 // 8192 faces, profiles/r02_summary.md): six block-wide phases per 2016 pixels at 4 blocks per SM.
  int warp_staged = 0;           // FRB_WARP_STAGED=1: stage each face's source box in shared memory first (bit-identical;
                                 // measured SLOWER, 1.14 vs 0.75 ms for 8192 faces: one 200 KB block per SM serialises
                                 // load and gather).  Default: the global-memory gather.
  cudaStream_t own_stream = nullptr;
  std::set<const void*> smem_attr_done;   // kernels whose dynamic shared-memory limit was raised on this device
};

namespace {

int fail(frb_ctx* c, const char* fmt, ...);
int stage_match(frb_ctx* ctx, int P, int k);
int ws_begin(frb_ctx* ctx, cudaStream_t st);
int ws_end(frb_ctx* ctx, cudaStream_t st);
bool stream_capturing(cudaStream_t st);
// dynamic shared memory opt-in: a function attribute is per device, so it is remembered per context, not per process
int set_smem_attr(frb_ctx* ctx, const void* kernel, int bytes);

int fail(frb_ctx* c, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return 1;
}

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(ctx, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

int set_smem_attr(frb_ctx* ctx, const void* kernel, int bytes) {
  if (ctx->smem_attr_done.count(kernel)) return 0;
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  ctx->smem_attr_done.insert(kernel);
  return 0;
}

template <typename T>
int ensure(frb_ctx* ctx, T** p, size_t* cap, size_t need) {
  if (*cap >= need && *p) return 0;
  if (*p) CK(cudaFree(*p));
  *p = nullptr;
  *cap = 0;
  CK(cudaMalloc(reinterpret_cast<void**>(p), need * sizeof(T)));
  *cap = need;
  return 0;
}

// CUTLASS applies the same fix-up for drivers <= 13.1 on tensors smaller than 128 KiB.
void small_tensor_fixup(frb_ctx* ctx, CUtensorMap* m, size_t bytes) {
  if (ctx->driver_version <= 13010 && bytes < 131072)
    reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
}

// 2D row-major bf16 tensor [rows][cols]; box = 64 cols x box_rows, 128B swizzle.
int make_tmap_2d(frb_ctx* ctx, CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = ctx->encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
                                 es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(ctx, "cuTensorMapEncodeTiled failed (%d) cols=%llu rows=%llu box_rows=%u", (int)r,
                (unsigned long long)cols, (unsigned long long)rows, box_rows);
  small_tensor_fixup(ctx, m, cols * rows * 2);
  return 0;
}

// NHWC bf16 activation [N][H][W][C] in im2col mode: 128 pixels x 64 channels per load.
int make_tmap_im2col(frb_ctx* ctx, CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int ksize, int stride,
                     int pad) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  // fprop corners (CUTLASS conv/collective/detail.hpp): lower = -pad, upper = pad - (ksize-1)*dilation
  int lower[2] = {-pad, -pad};
  int upper[2] = {pad - (ksize - 1), pad - (ksize - 1)};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = ctx->encode_im2col(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, lower,
                                  upper, 64, 128, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(ctx, "cuTensorMapEncodeIm2col failed (%d) N=%d H=%d W=%d C=%d k=%d s=%d p=%d", (int)r, N, H, W, C,
                ksize, stride, pad);
  small_tensor_fixup(ctx, m, (size_t)N * H * W * C * 2);
  return 0;
}

// cluster dimension + programmatic dependent launch (see ptx.cuh: pdl_wait)
void fill_launch_attrs(cudaLaunchAttribute* attr, int cluster) {
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
}

// launch with programmatic stream serialization (the kernel calls pdl_wait() before it touches its predecessor's output)
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, int cluster,
                       Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <int BN, int MODE, int CL>
int launch_gemm_t(frb_ctx* ctx, const CUtensorMap& a, const CUtensorMap& a2, const CUtensorMap& b, const GemmParams& gp,
                  int grid, cudaStream_t st) {
  auto kern = gemm_sm100_kernel<BN, MODE, CL>;
  if (set_smem_attr(ctx, reinterpret_cast<const void*>(kern), GemmSmem<BN>::kTotal)) return 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = GemmSmem<BN>::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  fill_launch_attrs(attr, CL);
  cfg.attrs = attr;
  cfg.numAttrs = ctx->use_pdl ? 2 : 1;
  CK(cudaLaunchKernelEx(&cfg, kern, a, a2, b, gp));
  ctx->launches++;
  return 0;
}

// cluster = CTAs sharing (multicasting) one weight tile; grid must be a multiple of it
int launch_gemm(frb_ctx* ctx, int block_n, int mode, int cluster, const CUtensorMap& a, const CUtensorMap& a2,
                const CUtensorMap& b, const GemmParams& gp, int grid, cudaStream_t st) {
  if (mode == A_IM2COL && cluster == 2) {
    if (block_n == 64) return launch_gemm_t<64, A_IM2COL, 2>(ctx, a, a2, b, gp, grid, st);
    if (block_n == 128) return launch_gemm_t<128, A_IM2COL, 2>(ctx, a, a2, b, gp, grid, st);
    if (block_n == 256) return launch_gemm_t<256, A_IM2COL, 2>(ctx, a, a2, b, gp, grid, st);
  } else if (mode == A_IM2COL && cluster == 1) {
    if (block_n == 64) return launch_gemm_t<64, A_IM2COL, 1>(ctx, a, a2, b, gp, grid, st);
    if (block_n == 128) return launch_gemm_t<128, A_IM2COL, 1>(ctx, a, a2, b, gp, grid, st);
    if (block_n == 256) return launch_gemm_t<256, A_IM2COL, 1>(ctx, a, a2, b, gp, grid, st);
  } else if (mode == A_TILED && cluster == 1) {
    if (block_n == 64) return launch_gemm_t<64, A_TILED, 1>(ctx, a, a2, b, gp, grid, st);
    if (block_n == 128) return launch_gemm_t<128, A_TILED, 1>(ctx, a, a2, b, gp, grid, st);
    if (block_n == 256) return launch_gemm_t<256, A_TILED, 1>(ctx, a, a2, b, gp, grid, st);
  }
  return fail(ctx, "unsupported gemm config block_n=%d mode=%d cluster=%d", block_n, mode, cluster);
}

constexpr int kConvCluster = 2;

template <int BN, int CL>
int launch_gemm2_t(frb_ctx* ctx, const CUtensorMap& a, const CUtensorMap& a2, const CUtensorMap& b, const GemmParams& gp,
                   int grid, cudaStream_t st) {
  auto kern = gemm2_sm100_kernel<BN, CL>;
  if (set_smem_attr(ctx, reinterpret_cast<const void*>(kern), Gemm2Smem<BN>::kTotal)) return 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemm2Threads);
  cfg.dynamicSmemBytes = Gemm2Smem<BN>::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  fill_launch_attrs(attr, CL);
  cfg.attrs = attr;
  cfg.numAttrs = ctx->use_pdl ? 2 : 1;
  CK(cudaLaunchKernelEx(&cfg, kern, a, a2, b, gp));
  ctx->launches++;
  return 0;
}

// How many 4-CTA clusters of the quad kernel can be resident at once (the GPCs of a B200 hold 16-20 SMs, so 4-CTA
// clusters strand a few SMs: 33 clusters = 132 of 148).  The persistent tile loop needs exactly one wave.
int quad_max_clusters(frb_ctx* ctx) {
  if (ctx->quad_clusters > 0) return ctx->quad_clusters;
  auto kern = gemm2_sm100_kernel<256, 4>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Smem<256>::kTotal);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctx->num_sms / 4 * 4);
  cfg.blockDim = dim3(kGemm2Threads);
  cfg.dynamicSmemBytes = Gemm2Smem<256>::kTotal;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = ctx->num_sms / 4 * 7 / 8;
  }
  ctx->quad_clusters = n;
  return n;
}

// Convolution launch: CTA-pair kernel (cta_group::2) by default; FRB_CONV_MODE=1 selects the
// 1-CTA kernel with 2-way weight multicast (kept for A/B profiling and as the FC/GEMM core).
int launch_conv(frb_ctx* ctx, int block_n, const CUtensorMap& a, const CUtensorMap& a2, const CUtensorMap& b,
                const GemmParams& gp, int grid, cudaStream_t st) {
  if (ctx->conv_mode == 1) return launch_gemm(ctx, block_n, A_IM2COL, kConvCluster, a, a2, b, gp, grid, st);
  if (gp.quad) {
    if (block_n == 128) return launch_gemm2_t<128, 4>(ctx, a, a2, b, gp, grid, st);
    if (block_n == 256) return launch_gemm2_t<256, 4>(ctx, a, a2, b, gp, grid, st);
    return fail(ctx, "unsupported quad conv block_n=%d", block_n);
  }
  if (block_n == 64) return launch_gemm2_t<64, 2>(ctx, a, a2, b, gp, grid, st);
  if (block_n == 128) return launch_gemm2_t<128, 2>(ctx, a, a2, b, gp, grid, st);
  if (block_n == 256) return launch_gemm2_t<256, 2>(ctx, a, a2, b, gp, grid, st);
  return fail(ctx, "unsupported conv block_n=%d", block_n);
}

// One persistent launch for a run of consecutive CTA-pair conv layers (gemm2_multi_sm100.cuh).
template <int BN>
int launch_gemm2_multi_t(frb_ctx* ctx, const Gemm2Layer* d_layers, int n, int* d_bar, int grid, cudaStream_t st) {
  auto kern = gemm2_multi_sm100_kernel<BN>;
  if (set_smem_attr(ctx, reinterpret_cast<const void*>(kern), Gemm2Smem<BN>::kTotal)) return 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemm2Threads);
  cfg.dynamicSmemBytes = Gemm2Smem<BN>::kTotal;
  cfg.stream = st;
  // The layers of a run wait for each other inside the kernel, so every CTA must be resident at the same time.
  // grid <= number of SMs with one CTA per SM gives that on an otherwise idle GPU; the cooperative attribute makes
  // the driver guarantee it (gang scheduling) even when other work shares the device.
  cudaLaunchAttribute base[2], attr[3];
  fill_launch_attrs(base, 2);   // [0] cluster dimension, [1] programmatic stream serialization
  auto launch = [&](bool pdl) {
    int na = 0;
    attr[na++] = base[0];
    if (pdl) attr[na++] = base[1];
    if (ctx->multi_coop) {
      attr[na].id = cudaLaunchAttributeCooperative;
      attr[na].val.cooperative = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, d_layers, n, d_bar);
  };
  // The cooperative launch is not combined with programmatic stream serialization (FRB_MULTI_COOP_PDL=1 does, for
  // experiments): how griddepcontrol.wait interacts with gang scheduling is not documented, and one plain stream
  // dependency per embed costs nothing measurable.
  const bool pdl = ctx->use_pdl && (!ctx->multi_coop || getenv("FRB_MULTI_COOP_PDL") != nullptr);
  cudaError_t e = launch(pdl);
  if (e != cudaSuccess && ctx->multi_coop) {
    // the driver refused the cooperative attribute (e.g. a partitioned or shared device where the grid cannot be
    // gang-scheduled): say so once and go on without the guarantee - correct whenever the GPU is not shared
    fprintf(stderr, "frb: cooperative launch of the persistent conv run refused (%s); launching it without the attribute\n",
            cudaGetErrorString(e));
    cudaGetLastError();
    ctx->multi_coop = 0;
    e = launch(ctx->use_pdl != 0);
  }
  CK(e);
  ctx->launches++;
  return 0;
}

int launch_gemm2_multi(frb_ctx* ctx, int block_n, const Gemm2Layer* d_layers, int n, int* d_bar, int grid, cudaStream_t st) {
  if (block_n == 64) return launch_gemm2_multi_t<64>(ctx, d_layers, n, d_bar, grid, st);
  if (block_n == 128) return launch_gemm2_multi_t<128>(ctx, d_layers, n, d_bar, grid, st);
  if (block_n == 256) return launch_gemm2_multi_t<256>(ctx, d_layers, n, d_bar, grid, st);
  return fail(ctx, "unsupported multi-layer conv block_n=%d", block_n);
}

// One persistent launch for a run of identical slab conv layers (conv_slab_multi_sm100.cuh).
template <int BN, int CH>
int launch_slab_multi_t(frb_ctx* ctx, const SlabLayer* d_layers, int n, int smem_bytes, int grid, cudaStream_t st) {
  auto kern = conv_slab_multi_sm100_kernel<BN, CH>;
  if (set_smem_attr(ctx, reinterpret_cast<const void*>(kern), 227 * 1024)) return 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemm2Threads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute base[2], attr[3];
  fill_launch_attrs(base, 2);
  auto launch = [&](bool pdl) {
    int na = 0;
    attr[na++] = base[0];
    if (pdl) attr[na++] = base[1];
    if (ctx->multi_coop) {
      attr[na].id = cudaLaunchAttributeCooperative;
      attr[na].val.cooperative = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, d_layers, n);
  };
  cudaError_t e = launch(ctx->use_pdl && !ctx->multi_coop);
  if (e != cudaSuccess && ctx->multi_coop) {
    fprintf(stderr, "frb: cooperative launch of the persistent slab run refused (%s); launching it without the attribute\n",
            cudaGetErrorString(e));
    cudaGetLastError();
    ctx->multi_coop = 0;
    e = launch(ctx->use_pdl != 0);
  }
  CK(e);
  ctx->launches++;
  return 0;
}

int launch_slab_multi(frb_ctx* ctx, int chunks, int N, const SlabLayer* d_layers, int n, int smem_bytes, int grid, cudaStream_t st) {
  if (chunks == 1 && N == 64) return launch_slab_multi_t<64, 1>(ctx, d_layers, n, smem_bytes, grid, st);
  if (chunks == 1 && N == 128) return launch_slab_multi_t<128, 1>(ctx, d_layers, n, smem_bytes, grid, st);
  if (chunks == 2 && N == 128) return launch_slab_multi_t<128, 2>(ctx, d_layers, n, smem_bytes, grid, st);
  return fail(ctx, "slab run: unsupported Cin chunks %d / Cout %d", chunks, N);
}

// NHWC bf16 activation [N][H][W][C] as a tiled 4-D tensor; box = 64 channels x box_w x box_h x 1 image.
int make_tmap_4d_tiled(frb_ctx* ctx, CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int box_w, int box_h) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = ctx->encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, "cuTensorMapEncodeTiled(4d) failed (%d)", (int)r);
  small_tensor_fixup(ctx, m, (size_t)N * H * W * C * 2);
  return 0;
}

int slab_rows(int w) { return w == 112 ? 1 : (w == 56 ? 2 : (w == 28 ? 4 : 0)); }

bool slab_eligible(const frb_ctx* ctx, const frb_layer_desc& L, bool has_sc) {
  if (ctx->conv_mode != 2 || !ctx->use_slab) return false;
  if (L.ksize != 3 || L.stride != 1 || L.pad != 1 || has_sc) return false;
  if (L.win < ctx->slab_min_w) return false;   // experiments: narrower layers go to the im2col pair kernel (and its persistent run)
  if (L.hin != L.win || slab_rows(L.win) == 0 || L.hin % slab_rows(L.win)) return false;
  // Cout = 256 (weights cannot stay resident) measured slower than the im2col pair kernel: not eligible
  if (!((L.cin == 64 && (L.cout == 64 || L.cout == 128)) || (L.cin == 128 && (L.cout == 128 || (L.cout == 256 && getenv("FRB_SLAB_N256")))))) return false;
  if (L.res_buf >= 0 && (L.res_stride != 1 || L.res_h != L.hin || L.res_w != L.win)) return false;
  return true;
}

int setup_slab(frb_ctx* ctx, const frb_layer_desc& L, int B, const void* d_in, const void* d_res, const void* d_w,
               const float* d_bias, const float* d_prelu, void* d_out, CUtensorMap* tmX, CUtensorMap* tmB,
               SlabParams* sp, int* smem_bytes, int* grid, CUtensorMap* tmOut = nullptr, CUtensorMap* tmRes = nullptr) {
  const int W = L.win, R = slab_rows(W);
  const int chunks = L.cin / 64, num_kb = 9 * chunks;
  memset(sp, 0, sizeof(*sp));
  sp->B = B; sp->H = L.hin; sp->W = W; sp->R = R;
  sp->N = L.cout;
  sp->box_bytes = 128 * (W + 2) * (R + 2);
  // the last tap's descriptor reads 128 rows starting 2 padded rows + 2 pixels into the buffer
  const int need = (128 + 2 * (W + 2) + 2) * 128;
  sp->slab_bytes = std::max((need + 1023) / 1024 * 1024, (sp->box_bytes + 1023) / 1024 * 1024);
  const int b_bytes = (L.cout / 2) * 128;
  // Cout = 64 layers (112 / 56 pixels): output staged in shared memory + one TMA store per tile (conv_slab_sm100.cuh)
  const bool stage_out = tmOut != nullptr && ctx->slab_stage_out && !ctx->slab_multi && L.cout == 64 && L.cin == 64 && !ctx->use_dataflow && R * W == 112;
  // ... and their shortcut tile comes in by TMA into the same staging buffers (three instead of two: fetched two tiles ahead)
  const bool res_tma = stage_out && tmRes != nullptr && d_res != nullptr && ctx->slab_res_tma;
  const int misc = 1024 + 10 * L.cout * 4 + 9 * kBiasPad * 4 + 1024 + slab_stage_bufs(stage_out, res_tma) * kSlabStageBuf;
  const int total = 227 * 1024;
  // A unit (one 64-channel chunk of a tile's slab) is only 36 MMAs (0.6-2.4 us): several must be in flight
  // to hide the ~2 us load latency.  Weights stay resident when three units still fit beside them.
  int nbuf, st;
  if (num_kb * b_bytes + 3 * sp->slab_bytes + misc <= total) {
    st = num_kb;  // resident
    nbuf = std::min(kSlabMaxBuf, (total - misc - st * b_bytes) / sp->slab_bytes);
  } else {
    nbuf = 3;
    st = std::min(kSlabMaxBStages, (total - misc - nbuf * sp->slab_bytes) / b_bytes);
    if (st >= num_kb) st = num_kb - 1;  // == num_kb means "resident" to the kernel
  }
  if (st < 3 || nbuf < 2) return fail(ctx, "slab conv: not enough shared memory (stages %d, buffers %d)", st, nbuf);
  sp->b_stages = st;
  sp->nbuf = nbuf;
  sp->stage_out = stage_out ? 1 : 0;
  sp->res_tma = res_tma ? 1 : 0;
  *smem_bytes = nbuf * sp->slab_bytes + st * b_bytes + misc;
  sp->bias = d_bias; sp->bias_cases = L.bias_cases;
  sp->prelu = L.has_prelu ? d_prelu : nullptr;
  sp->residual = reinterpret_cast<const __nv_bfloat16*>(d_res);
  sp->out = reinterpret_cast<__nv_bfloat16*>(d_out);
  if (const char* e = getenv("FRB_SLAB_DEBUG")) sp->debug = atoi(e);
  if (const char* e = getenv("FRB_SLAB_TRACE")) sp->trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  if (make_tmap_4d_tiled(ctx, tmX, d_in, B, L.hin, W, L.cin, (sp->debug & 32) ? W : W + 2, (sp->debug & 16) ? 1 : R + 2)) return 1;
  if (make_tmap_2d(ctx, tmB, d_w, num_kb * 64, (sp->debug & 64) ? 74 * L.cout : L.cout, L.cout / 2)) return 1;
  if (tmOut) {
    if (stage_out) {
      if (make_tmap_2d(ctx, tmOut, d_out, L.cout, static_cast<uint64_t>(B) * L.hin * W, R * W)) return 1;
    } else {
      *tmOut = *tmB;   // unused by the kernel
    }
  }
  if (tmRes) {
    if (res_tma) {
      if (make_tmap_2d(ctx, tmRes, d_res, L.cout, static_cast<uint64_t>(B) * L.hin * W, R * W)) return 1;
    } else {
      *tmRes = *tmB;   // unused by the kernel
    }
  }
  const int pairs = (B * (L.hin / R) + 1) / 2;
  *grid = std::min(pairs, ctx->num_sms / 2) * 2;
  return 0;
}

template <int BN, int CH, bool RES = false>
int launch_slab_t(frb_ctx* ctx, const CUtensorMap& x, const CUtensorMap& b, const CUtensorMap& o, const CUtensorMap& r,
                  const SlabParams& sp, int smem_bytes, int grid, cudaStream_t st) {
  auto kern = conv_slab_sm100_kernel<BN, CH, RES>;
  if (set_smem_attr(ctx, reinterpret_cast<const void*>(kern), 227 * 1024)) return 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemm2Threads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  fill_launch_attrs(attr, 2);
  cfg.attrs = attr;
  cfg.numAttrs = ctx->use_pdl ? 2 : 1;
  CK(cudaLaunchKernelEx(&cfg, kern, x, b, o, r, sp));
  ctx->launches++;
  return 0;
}

// r: tensor map of the shortcut for sp.res_tma layers (setup_slab's tmRes), any valid map otherwise
int launch_slab(frb_ctx* ctx, int chunks, const CUtensorMap& x, const CUtensorMap& b, const CUtensorMap& o, const CUtensorMap& r,
                const SlabParams& sp, int smem_bytes, int grid, cudaStream_t st) {
  if (chunks == 1) {
    if (sp.N == 64 && sp.res_tma) return launch_slab_t<64, 1, true>(ctx, x, b, o, r, sp, smem_bytes, grid, st);
    if (sp.N == 64) return launch_slab_t<64, 1>(ctx, x, b, o, r, sp, smem_bytes, grid, st);
    if (sp.N == 128) return launch_slab_t<128, 1>(ctx, x, b, o, r, sp, smem_bytes, grid, st);
  } else if (chunks == 2) {
    if (sp.N == 128) return launch_slab_t<128, 2>(ctx, x, b, o, r, sp, smem_bytes, grid, st);
    if (sp.N == 256) return launch_slab_t<256, 2>(ctx, x, b, o, r, sp, smem_bytes, grid, st);
  }
  return fail(ctx, "slab conv: unsupported Cin chunks %d / Cout %d", chunks, sp.N);
}

int pick_block_n(int cout) { return cout >= 256 ? 256 : (cout >= 128 ? 128 : 64); }

int out_dim(int in, int ksize, int stride, int pad) { return (in + 2 * pad - ksize) / stride + 1; }

// Fill GemmParams + tensor maps for one CONV layer.
int setup_conv(frb_ctx* ctx, const frb_layer_desc& L, int B, const void* d_in, const void* d_sc, const void* d_res,
               const void* d_w, const float* d_bias, const float* d_prelu, void* d_out, CUtensorMap* tmA,
               CUtensorMap* tmA2, CUtensorMap* tmB, GemmParams* gp, int* block_n, int* grid) {
  if (L.cin % 64 || L.cout % 64) return fail(ctx, "conv channels must be multiples of 64 (cin=%d cout=%d)", L.cin, L.cout);
  if (!((L.ksize == 3 && L.pad == 1) || (L.ksize == 1 && L.pad == 0))) return fail(ctx, "unsupported conv geometry");
  const int P = out_dim(L.hin, L.ksize, L.stride, L.pad), Q = out_dim(L.win, L.ksize, L.stride, L.pad);
  memset(gp, 0, sizeof(*gp));
  gp->M = B * P * Q;
  gp->N = L.cout;
  gp->cin_chunks = L.cin / 64;
  gp->num_kb_main = L.ksize * L.ksize * gp->cin_chunks;
  gp->num_kb_sc = (L.sc_buf >= 0 || d_sc) && L.sc_cin > 0 ? L.sc_cin / 64 : 0;
  gp->num_splits = 1;
  gp->P = P;
  gp->Q = Q;
  gp->stride = L.stride;
  gp->pad = L.pad;
  gp->sc_stride = L.sc_stride;
  gp->sc_chunks = gp->num_kb_sc;
  gp->bias = d_bias;
  gp->bias_cases = L.bias_cases;
  gp->prelu = L.has_prelu ? d_prelu : nullptr;
  gp->residual = reinterpret_cast<const __nv_bfloat16*>(d_res);
  gp->res_stride = L.res_stride;
  gp->RH = L.res_h;
  gp->RW = L.res_w;
  gp->out = reinterpret_cast<__nv_bfloat16*>(d_out);
  gp->out_f32 = nullptr;
  if (make_tmap_im2col(ctx, tmA, d_in, B, L.hin, L.win, L.cin, L.ksize, L.stride, L.pad)) return 1;
  if (gp->num_kb_sc > 0) {
    if (out_dim(L.sc_hin, 1, L.sc_stride, 0) != P || out_dim(L.sc_win, 1, L.sc_stride, 0) != Q)
      return fail(ctx, "shortcut conv output size mismatch");
    if (make_tmap_im2col(ctx, tmA2, d_sc, B, L.sc_hin, L.sc_win, L.sc_cin, 1, L.sc_stride, 0)) return 1;
  } else {
    *tmA2 = *tmA;
  }
  const int ktot = (gp->num_kb_main + gp->num_kb_sc) * 64;
  *block_n = pick_block_n(L.cout);
  const int m_super = ((gp->M + 127) / 128 + kConvCluster - 1) / kConvCluster;
  // quad clusters (weight multicast across two CTA pairs) where the layer is bound by L2->SM ingest: Cout >= 256
  // (FRB_QUAD: 0 = never, 1 = block_n 256, 2 = block_n >= 128); not with the per-image dataflow schedule
  gp->quad = (ctx->conv_mode == 2 && !ctx->use_dataflow && ctx->conv_quad > 0 &&
              (*block_n == 256 || (*block_n == 128 && ctx->conv_quad >= 2)) && m_super >= 2) ? 1 : 0;
  if (gp->quad) {
    if (make_tmap_2d(ctx, tmB, d_w, ktot, L.cout, *block_n / 4)) return 1;
    const int units = (m_super + 1) / 2 * (L.cout / *block_n);
    *grid = std::min(units, quad_max_clusters(ctx)) * 4;
    return 0;
  }
  if (make_tmap_2d(ctx, tmB, d_w, ktot, L.cout, *block_n / kConvCluster)) return 1;
  const int units = m_super * (L.cout / *block_n);
  *grid = std::min(units, ctx->num_sms / kConvCluster) * kConvCluster;
  // split-K for the last, partially filled round (14x14 layers at batch 256: 196 tiles on 74 pairs = 2 rounds + 48
  // tiles; cut in 3 K ranges those 48 tiles are 144 items = 2 short sub-rounds, 2.67 rounds instead of 3)
  gp->tail_split = 0;
  const int pairs = *grid / 2;
  if (ctx->tail_split && ctx->conv_mode == 2 && !ctx->use_dataflow && gp->num_kb_sc == 0 && units >= pairs && units % pairs) {
    const int full = units / pairs, rem = units % pairs;
    int best = 1;
    double best_cost = 1.0;
    for (int sp = 2; sp <= 4; ++sp) {
      if (gp->num_kb_main % sp) continue;
      const double cost = static_cast<double>((rem * sp + pairs - 1) / pairs) / sp + 0.05;  // + the partial round trip
      if (cost < best_cost - 1e-9) { best_cost = cost; best = sp; }
    }
    if (best > 1 && (1.0 - best_cost) / (full + 1) >= 0.02) gp->tail_split = best;
    if (const char* e = getenv("FRB_TS_DEBUG")) gp->tail_debug = atoi(e);
    if (const char* e = getenv("FRB_TS_FORCE")) gp->tail_split = atoi(e);
  }
  return 0;
}

// split-K scratch of one conv launch: rem tiles x (split-1) fp32 partial tiles + rem flags
void tail_split_needs(const GemmParams& gp, int block_n, int grid, size_t* partial_floats, size_t* flags) {
  *partial_floats = 0;
  *flags = 0;
  if (gp.tail_split <= 1) return;
  const int pairs = grid / 2;
  const int units = (((gp.M + 127) / 128 + 1) / 2) * (gp.N / block_n);
  const int rem = units % pairs;
  *flags = static_cast<size_t>(rem);
  *partial_floats = static_cast<size_t>(rem) * (gp.tail_split - 1) * 256 * block_n;
}

int choose_splits(int num_kb, int want) {
  want = std::max(1, std::min(want, num_kb));
  const int per = (num_kb + want - 1) / want;
  return (num_kb + per - 1) / per;
}

}  // namespace

// ====================================================================== ctx
extern "C" int frb_ctx_create(int device, frb_ctx** out) {
  if (!out) return 1;
  *out = nullptr;
  frb_ctx* ctx = new frb_ctx();
  ctx->device = device;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    fprintf(stderr, "frb_ctx_create: cudaSetDevice(%d): %s\n", device, cudaGetErrorString(e));
    delete ctx;
    return 1;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess || prop.major != 10) {
    fprintf(stderr, "frb_ctx_create: device %d is not sm_100 (cc %d.%d); this library has no fallback path\n", device,
            prop.major, prop.minor);
    delete ctx;
    return 2;
  }
  ctx->num_sms = prop.multiProcessorCount;
  cudaDriverGetVersion(&ctx->driver_version);
  if (const char* e = getenv("FRB_CONV_MODE")) ctx->conv_mode = atoi(e);
  if (const char* e = getenv("FRB_SLAB")) ctx->use_slab = atoi(e);
  if (const char* e = getenv("FRB_SLAB_STAGE")) ctx->slab_stage_out = atoi(e);
  if (const char* e = getenv("FRB_SLAB_RES_TMA")) ctx->slab_res_tma = atoi(e);
  if (const char* e = getenv("FRB_QUAD")) ctx->conv_quad = atoi(e);
  if (const char* e = getenv("FRB_MULTI")) ctx->conv_multi = atoi(e);
  // Nsight Compute / compute-sanitizer cannot run the cooperative cluster launch of the persistent runs (the launch
  // fails and takes the CUDA context with it): when the process was started under such a tool (they announce
  // themselves through these variables) the runs go out without the cooperative attribute.  Tools serialise kernels,
  // so every CTA of the run is resident anyway - the guarantee the attribute exists for.
  for (const char* v : {"NV_COMPUTE_PROFILER_PERFWORKS_DIR", "NV_NSIGHT_INJECTION_TRANSPORT_TYPE", "NV_TPS_LAUNCH_TOKEN",
                        "CUDA_INJECTION64_PATH", "NV_SANITIZER_INJECTION_PORT_BASE"})
    if (getenv(v)) {
      ctx->multi_coop = 0;
      break;
    }
  if (const char* e = getenv("FRB_MULTI_COOP")) ctx->multi_coop = atoi(e);
  if (const char* e = getenv("FRB_SLAB_MULTI")) ctx->slab_multi = atoi(e);
  if (const char* e = getenv("FRB_WARP_STAGED")) ctx->warp_staged = atoi(e);
  if (const char* e = getenv("FRB_WARP_BAND")) ctx->warp_band = atoi(e);
  if (const char* e = getenv("FRB_MATCH_PREFETCH")) ctx->match_prefetch = atoi(e);
  if (const char* e = getenv("FRB_SLAB_MINW")) ctx->slab_min_w = atoi(e);
  if (const char* e = getenv("FRB_TAIL_SPLIT")) ctx->tail_split = atoi(e);
  if (const char* e = getenv("FRB_PDL")) ctx->use_pdl = atoi(e);
  if (const char* e = getenv("FRB_EMBED_CHUNK")) ctx->embed_chunk = atoi(e);
  if (const char* e = getenv("FRB_FRONT_SUB")) ctx->front_sub = atoi(e);
  if (const char* e = getenv("FRB_DATAFLOW")) ctx->use_dataflow = atoi(e);
  if (const char* e = getenv("FRB_MATCH_PAIR")) ctx->match_pair = atoi(e);
  if (!ctx->use_pdl) ctx->use_dataflow = 0;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || !fn) { fprintf(stderr, "frb: no cuTensorMapEncodeTiled\n"); delete ctx; return 3; }
  ctx->encode_tiled = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  fn = nullptr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || !fn) { fprintf(stderr, "frb: no cuTensorMapEncodeIm2col\n"); delete ctx; return 3; }
  ctx->encode_im2col = reinterpret_cast<PFN_cuTensorMapEncodeIm2col_v12000>(fn);

  // normalisation LUT: float32((x/255 - 0.5)/0.5) in f64 as numpy does (face_embedder.py:100), then bf16 RN
  unsigned short lut[256];
  for (int i = 0; i < 256; ++i) {
    const float f = static_cast<float>((static_cast<double>(i) / 255.0 - 0.5) / 0.5);
    const __nv_bfloat16 b = __float2bfloat16_rn(f);
    memcpy(&lut[i], &b, 2);
  }
  // bilinear weight table as cv::initInterTab2D(INTER_LINEAR, fixpt=true) builds it
  std::vector<unsigned short> wtab(32 * 32 * 4);
  for (int ay = 0; ay < 32; ++ay)
    for (int ax = 0; ax < 32; ++ax) {
      const float fx = ax * (1.f / 32), fy = ay * (1.f / 32);
      const float tx[2] = {1.f - fx, fx}, ty[2] = {1.f - fy, fy};
      int iw[4], sum = 0;
      for (int k1 = 0; k1 < 2; ++k1)
        for (int k2 = 0; k2 < 2; ++k2) {
          const float v = ty[k1] * tx[k2];
          int q = static_cast<int>(lrintf(v * 32768.f));
          q = std::max(-32768, std::min(32767, q));
          iw[k1 * 2 + k2] = q;
          sum += q;
        }
      if (sum != 32768) {
        const int diff = sum - 32768;
        int kmax = 0, kmin = 0;  // first max / first min, scanning k1 then k2 within the central 2x2
        for (int k = 1; k < 4; ++k) {
          if (iw[k] > iw[kmax]) kmax = k;
          if (iw[k] < iw[kmin]) kmin = k;
        }
        if (diff < 0) iw[kmax] -= diff; else iw[kmin] -= diff;
      }
      for (int k = 0; k < 4; ++k) wtab[(ay * 32 + ax) * 4 + k] = static_cast<unsigned short>(iw[k]);
    }
  if (cudaMalloc(&ctx->d_lut, sizeof(lut)) != cudaSuccess || cudaMalloc(&ctx->d_wtab, wtab.size() * 2) != cudaSuccess ||
      cudaMemcpy(ctx->d_lut, lut, sizeof(lut), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(ctx->d_wtab, wtab.data(), wtab.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMalloc(&ctx->d_gal_maxnorm, 8) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    fprintf(stderr, "frb_ctx_create: constant upload failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    delete ctx;
    return 4;
  }
  *out = ctx;
  return 0;
}

extern "C" void frb_ctx_destroy(frb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  void* ptrs[] = {ctx->d_lut, ctx->d_wtab, ctx->d_blob, ctx->d_fc_partial, ctx->d_emb2, ctx->d_gal, ctx->d_gal_bf16,
                  ctx->d_gal_maxnorm, ctx->d_probe_f32, ctx->d_probe_bf16, ctx->d_cand_score, ctx->d_cand_idx,
                  ctx->d_flagged, ctx->d_flag_rows, ctx->d_row_floor, ctx->d_exact, ctx->d_scores64_tmp, ctx->d_stage_u8, ctx->prefetch[0].d_buf, ctx->prefetch[1].d_buf,
                  ctx->d_stage_in, ctx->d_stage_emb, ctx->d_stage_norm, ctx->d_stage_sc, ctx->d_stage_idx,
                  ctx->d_stage_acc, ctx->d_jobs, ctx->d_progress, ctx->d_seg, ctx->d_sample_identity, ctx->d_id_top_idx,
                  ctx->d_id_top_sc, ctx->d_id_acc, ctx->d_id_sc32, ctx->d_id_scores, ctx->d_tail_partial, ctx->d_tail_flags, ctx->d_runs, ctx->d_run_bar, ctx->d_sruns};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (void* p : {static_cast<void*>(ctx->d_match_ctr), static_cast<void*>(ctx->d_exact_part), static_cast<void*>(ctx->d_merge_rec),
                  static_cast<void*>(ctx->xchg.local)})
    if (p) cudaFree(p);
  for (int g = 0; g < kMaxPeers; ++g)
    if (ctx->xchg.peer_ipc[g]) cudaIpcCloseMemHandle(ctx->xchg.peer[g]);
  if (ctx->h_flag_count) cudaFreeHost(ctx->h_flag_count);
  if (ctx->xchg.h_status) cudaFreeHost(ctx->xchg.h_status);
  if (ctx->flag_event) cudaEventDestroy(ctx->flag_event);
  if (ctx->ws_event) cudaEventDestroy(ctx->ws_event);
  for (auto e : ctx->prof_events) cudaEventDestroy(e);
  for (auto e : ctx->match_prof_ev) if (e) cudaEventDestroy(e);
  for (auto* b : ctx->d_bufs)
    if (b) cudaFree(b);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (auto& sl : ctx->prefetch) if (sl.done) cudaEventDestroy(sl.done);
  delete ctx;
}

extern "C" const char* frb_last_error(frb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
extern "C" long long frb_launch_count(frb_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ====================================================================== preprocessing
extern "C" int frb_preprocess_u8(frb_ctx* ctx, const void* d_in, int B, int S, void* d_out, int flip, void* stream) {
  if (!ctx) return 1;
  if (S != 112 && S != 224) return fail(ctx, "frb_preprocess_u8: S must be 112 or 224 (got %d)", S);
  if (B <= 0) return 0;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  const size_t total = static_cast<size_t>(B) * 112 * 112;
  preprocess_u8_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint8_t*>(d_in), B, S, ctx->d_lut, reinterpret_cast<__nv_bfloat16*>(d_out), flip);
  CK(cudaGetLastError());
  ctx->launches++;
  return 0;
}

extern "C" int frb_warp_normalize(frb_ctx* ctx, const void* d_src_base, const frb_warp_job* h_jobs, int B, int S,
                                  void* d_out_u8, void* d_out_bf16, void* stream) {
  if (!ctx) return 1;
  if (B <= 0) return 0;
  if (d_out_bf16 && S != 112) return fail(ctx, "frb_warp_normalize: bf16 output requires S == 112");
  if (!d_out_u8 && !d_out_bf16) return fail(ctx, "frb_warp_normalize: no output requested");
  if (S < 64 || S > kWarpMaxS) return fail(ctx, "frb_warp_normalize: S must be in [64, %d] (got %d)", kWarpMaxS, S);
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ws_begin(ctx, st)) return 1;
  if (ctx->jobs_cap < B) {
    if (ctx->d_jobs) CK(cudaFree(ctx->d_jobs));
    ctx->d_jobs = nullptr;
    const int cap = (B + 1) & ~1;   // even: the boxes behind the 72-byte jobs stay 16-byte aligned
    CK(cudaMalloc(&ctx->d_jobs, (sizeof(WarpJob) + sizeof(int4)) * cap));   // jobs, then one source box per face
    ctx->jobs_cap = cap;
  }
  CK(cudaMemcpyAsync(ctx->d_jobs, h_jobs, sizeof(WarpJob) * B, cudaMemcpyHostToDevice, st));
  const uint8_t* src = reinterpret_cast<const uint8_t*>(d_src_base);
  uint8_t* o8 = reinterpret_cast<uint8_t*>(d_out_u8);
  __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(d_out_bf16);
  // Source box of every face (the output square's corners through the inverse map, two pixels of slack, clamped to the
  // image).  When all of them fit in shared memory the staged kernel reads each source byte once with coalesced loads.
  if (ctx->warp_staged) {
    ctx->h_boxes.resize(static_cast<size_t>(B) * 4);
    size_t max_bytes = 0;
    const WarpJob* hj = reinterpret_cast<const WarpJob*>(h_jobs);
    for (int i = 0; i < B; ++i) {
      const double* M = hj[i].M;
      double D = M[0] * M[4] - M[1] * M[3];
      D = D != 0.0 ? 1.0 / D : 0.0;
      const double m00 = M[4] * D, m01 = -M[1] * D, m10 = -M[3] * D, m11 = M[0] * D;
      const double b1 = -m00 * M[2] - m01 * M[5], b2 = -m10 * M[2] - m11 * M[5];
      double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
      for (int c = 0; c < 4; ++c) {
        const double x = (c & 1) ? S - 1 : 0, y = (c & 2) ? S - 1 : 0;
        const double sx = m00 * x + m01 * y + b1, sy = m10 * x + m11 * y + b2;
        xmin = std::min(xmin, sx); xmax = std::max(xmax, sx); ymin = std::min(ymin, sy); ymax = std::max(ymax, sy);
      }
      int x0 = static_cast<int>(std::floor(xmin)) - 2, x1 = static_cast<int>(std::ceil(xmax)) + 3;
      int y0 = static_cast<int>(std::floor(ymin)) - 2, y1 = static_cast<int>(std::ceil(ymax)) + 3;
      x0 = std::max(x0, 0); y0 = std::max(y0, 0); x1 = std::min(x1, hj[i].W - 1); y1 = std::min(y1, hj[i].H - 1);
      if (!(std::isfinite(xmin) && std::isfinite(xmax) && std::isfinite(ymin) && std::isfinite(ymax))) { x0 = y0 = 0; x1 = y1 = -1; }
      int* bx = &ctx->h_boxes[static_cast<size_t>(i) * 4];
      bx[0] = x0; bx[1] = y0; bx[2] = x1; bx[3] = y1;
      if (x1 >= x0 && y1 >= y0)
        max_bytes = std::max(max_bytes, static_cast<size_t>(y1 - y0 + 1) * warp_staged_pitch(x0, x1));
    }
    if (max_bytes <= static_cast<size_t>(kWarpStagedMaxBytes)) {
      int4* d_boxes = reinterpret_cast<int4*>(reinterpret_cast<uint8_t*>(ctx->d_jobs) + sizeof(WarpJob) * static_cast<size_t>(ctx->jobs_cap));
      CK(cudaMemcpyAsync(d_boxes, ctx->h_boxes.data(), sizeof(int4) * B, cudaMemcpyHostToDevice, st));
      if (set_smem_attr(ctx, reinterpret_cast<const void*>(warp_normalize_staged_kernel<true, true>), kWarpStagedMaxBytes)) return 1;
      if (set_smem_attr(ctx, reinterpret_cast<const void*>(warp_normalize_staged_kernel<true, false>), kWarpStagedMaxBytes)) return 1;
      if (set_smem_attr(ctx, reinterpret_cast<const void*>(warp_normalize_staged_kernel<false, true>), kWarpStagedMaxBytes)) return 1;
      const size_t smem = std::max<size_t>(max_bytes, 16);
      if (d_out_u8 && d_out_bf16)
        warp_normalize_staged_kernel<true, true><<<B, kWarpStagedThreads, smem, st>>>(src, ctx->d_jobs, d_boxes, S, ctx->d_wtab, ctx->d_lut, o8, ob);
      else if (d_out_u8)
        warp_normalize_staged_kernel<true, false><<<B, kWarpStagedThreads, smem, st>>>(src, ctx->d_jobs, d_boxes, S, ctx->d_wtab, ctx->d_lut, o8, ob);
      else
        warp_normalize_staged_kernel<false, true><<<B, kWarpStagedThreads, smem, st>>>(src, ctx->d_jobs, d_boxes, S, ctx->d_wtab, ctx->d_lut, o8, ob);
      CK(cudaGetLastError());
      ctx->launches++;
      return ws_end(ctx, st);
    }
  }
  if (ctx->warp_band) {   // experiment: every band's source spans staged in shared memory (bit-identical, coalesced, slower)
    const int rpb = warp_band_rows(S);
    dim3 bgrid((S + rpb - 1) / rpb, B);
    if (set_smem_attr(ctx, reinterpret_cast<const void*>(warp_normalize_band_kernel<true, true>), kWarpBandStageBytes)) return 1;
    if (set_smem_attr(ctx, reinterpret_cast<const void*>(warp_normalize_band_kernel<true, false>), kWarpBandStageBytes)) return 1;
    if (set_smem_attr(ctx, reinterpret_cast<const void*>(warp_normalize_band_kernel<false, true>), kWarpBandStageBytes)) return 1;
    if (d_out_u8 && d_out_bf16)
      warp_normalize_band_kernel<true, true><<<bgrid, kWarpBandThreads, kWarpBandStageBytes, st>>>(src, ctx->d_jobs, S, ctx->d_wtab, ctx->d_lut, o8, ob);
    else if (d_out_u8)
      warp_normalize_band_kernel<true, false><<<bgrid, kWarpBandThreads, kWarpBandStageBytes, st>>>(src, ctx->d_jobs, S, ctx->d_wtab, ctx->d_lut, o8, ob);
    else
      warp_normalize_band_kernel<false, true><<<bgrid, kWarpBandThreads, kWarpBandStageBytes, st>>>(src, ctx->d_jobs, S, ctx->d_wtab, ctx->d_lut, o8, ob);
    CK(cudaGetLastError());
    ctx->launches++;
    return ws_end(ctx, st);
  }
  dim3 grid((S * S + kWarpPixPerBlock - 1) / kWarpPixPerBlock, B);
  if (d_out_u8 && d_out_bf16)
    warp_normalize_kernel<true, true><<<grid, kWarpThreads, 0, st>>>(src, ctx->d_jobs, S, ctx->d_wtab, ctx->d_lut, o8, ob);
  else if (d_out_u8)
    warp_normalize_kernel<true, false><<<grid, kWarpThreads, 0, st>>>(src, ctx->d_jobs, S, ctx->d_wtab, ctx->d_lut, o8, ob);
  else
    warp_normalize_kernel<false, true><<<grid, kWarpThreads, 0, st>>>(src, ctx->d_jobs, S, ctx->d_wtab, ctx->d_lut, o8, ob);
  CK(cudaGetLastError());
  ctx->launches++;
  return ws_end(ctx, st);
}

// ====================================================================== backbone
extern "C" int frb_backbone_load(frb_ctx* ctx, const frb_layer_desc* layers, int n_layers, const void* h_blob,
                                 size_t blob_bytes, int n_bufs) {
  if (!ctx) return 1;
  if (!layers || n_layers <= 0 || !h_blob || n_bufs <= 0) return fail(ctx, "frb_backbone_load: bad arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  // validate everything before touching the resident state: a rejected program leaves the ctx as it was
  std::vector<size_t> elems(n_bufs, 0);
  double flops = 0.0;
  auto in_blob = [&](long long off, long long bytes) {
    return off >= 0 && bytes >= 0 && static_cast<unsigned long long>(off) + static_cast<unsigned long long>(bytes) <= blob_bytes;
  };
  for (int i = 0; i < n_layers; ++i) {
    const frb_layer_desc& L = layers[i];
    if (L.op != FRB_OP_STEM && L.op != FRB_OP_CONV && L.op != FRB_OP_FC) return fail(ctx, "layer %d: unknown op %d", i, L.op);
    if (L.cin <= 0 || L.cout <= 0 || L.hin <= 0 || L.win <= 0 || L.ksize <= 0 || L.stride <= 0 || L.pad < 0)
      return fail(ctx, "layer %d: bad geometry", i);
    if (!in_blob(L.w_off, L.w_bytes)) return fail(ctx, "layer %d: weights outside blob", i);
    const int cases = L.bias_cases == 9 ? 9 : 1;
    if (L.bias_cases != 1 && L.bias_cases != 9) return fail(ctx, "layer %d: bias_cases must be 1 or 9", i);
    if (!in_blob(L.bias_off, static_cast<long long>(cases) * L.cout * 4)) return fail(ctx, "layer %d: bias outside blob", i);
    if (L.has_prelu && !in_blob(L.prelu_off, static_cast<long long>(L.cout) * 4)) return fail(ctx, "layer %d: PReLU slopes outside blob", i);
    if (L.out_buf < 0 || L.out_buf >= n_bufs) return fail(ctx, "layer %d: bad out_buf", i);
    if (L.in_buf >= n_bufs) return fail(ctx, "layer %d: bad in_buf", i);
    if (L.sc_buf >= n_bufs) return fail(ctx, "layer %d: bad sc_buf", i);
    if (L.res_buf >= n_bufs) return fail(ctx, "layer %d: bad res_buf", i);
    if (L.op != FRB_OP_FC && (L.in_buf == L.out_buf || (L.sc_buf >= 0 && L.sc_buf == L.out_buf)))
      return fail(ctx, "layer %d: a layer cannot write the buffer it reads", i);
    const int P = out_dim(L.hin, L.ksize, L.stride, L.pad), Q = out_dim(L.win, L.ksize, L.stride, L.pad);
    if (L.op != FRB_OP_FC && (P <= 0 || Q <= 0)) return fail(ctx, "layer %d: empty output", i);
    size_t out_elems;
    if (L.op == FRB_OP_FC) {
      out_elems = 0;  // FC writes fp32 partials, not an activation buffer
      flops += 2.0 * L.cin * L.cout;
    } else {
      out_elems = static_cast<size_t>(P) * Q * L.cout;
      flops += 2.0 * P * Q * L.cout * (static_cast<double>(L.ksize) * L.ksize * L.cin + (L.sc_buf >= 0 ? L.sc_cin : 0));
    }
    elems[L.out_buf] = std::max(elems[L.out_buf], out_elems);
  }
  uint8_t* d_new = nullptr;
  CK(cudaMalloc(&d_new, blob_bytes));
  if (cudaMemcpy(d_new, h_blob, blob_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(d_new);
    return fail(ctx, "frb_backbone_load: weight upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  // commit
  CK(cudaDeviceSynchronize());
  if (ctx->d_blob) CK(cudaFree(ctx->d_blob));
  ctx->d_blob = d_new;
  ctx->blob_bytes = blob_bytes;
  ctx->layers.assign(layers, layers + n_layers);
  for (auto* b : ctx->d_bufs)
    if (b) CK(cudaFree(b));
  ctx->d_bufs.assign(n_bufs, nullptr);
  ctx->bufs_capacity_B = 0;
  ctx->n_bufs = n_bufs;
  ctx->buf_elems_per_face = elems;
  ctx->plan = Plan();
  ctx->flops_per_face = flops;
  ctx->backbone_gen++;
  return 0;
}

extern "C" double frb_backbone_flops_per_face(frb_ctx* ctx) { return ctx ? ctx->flops_per_face : 0.0; }
// schedule of the last embed on this ctx: out4 = {faces per pass (chunk), front layers, images per front sub-batch, sub-batches}
extern "C" int frb_embed_schedule(frb_ctx* ctx, int* out4) {
  if (!ctx || !out4) return 1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  out4[0] = ctx->plan.B;
  out4[1] = ctx->front.nf;
  out4[2] = ctx->front.nf > 0 ? ctx->front.sub : 0;
  out4[3] = ctx->front.nf > 0 ? ctx->front.n_sub : 0;
  return 0;
}

namespace {

int build_plan(frb_ctx* ctx, int B, const void* d_in) {
  // (re)allocate activation buffers
  if (ctx->bufs_capacity_B < B) {
    for (int i = 0; i < ctx->n_bufs; ++i) {
      if (ctx->d_bufs[i]) CK(cudaFree(ctx->d_bufs[i]));
      ctx->d_bufs[i] = nullptr;
      const size_t elems = std::max<size_t>(ctx->buf_elems_per_face[i], 64) * B;
      CK(cudaMalloc(reinterpret_cast<void**>(&ctx->d_bufs[i]), elems * 2));
    }
    ctx->bufs_capacity_B = B;
  }
  const size_t nl = ctx->layers.size();
  Plan& pl = ctx->plan;
  if (ctx->progress_cap < B) {
    if (ctx->d_progress) CK(cudaFree(ctx->d_progress));
    ctx->d_progress = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&ctx->d_progress), sizeof(int) * B));
    ctx->progress_cap = B;
  }
  pl.wait_target.assign(nl, 0);
  {
    long long units = 0;  // (output row x 32-channel chunk) units per image written by the layers so far
    for (size_t i = 0; i < nl; ++i) {
      const frb_layer_desc& L = ctx->layers[i];
      pl.wait_target[i] = static_cast<int>(units);
      if (L.op != FRB_OP_FC)
        units += static_cast<long long>(out_dim(L.hin, L.ksize, L.stride, L.pad)) * out_dim(L.win, L.ksize, L.stride, L.pad) * (L.cout / 32);
    }
    if (units > 0x7fffffffLL) return fail(ctx, "dataflow progress counter would overflow");
  }
  pl.B = B;
  pl.tmA.resize(nl); pl.tmA2.resize(nl); pl.tmB.resize(nl); pl.tmR.resize(nl); pl.gp.resize(nl); pl.block_n.assign(nl, 0); pl.grid.assign(nl, 0);
  pl.use_slab.assign(nl, 0); pl.sp.resize(nl); pl.slab_smem.assign(nl, 0);
  for (size_t i = 0; i < nl; ++i) {
    const frb_layer_desc& L = ctx->layers[i];
    const void* in = L.in_buf < 0 ? d_in : ctx->d_bufs[L.in_buf];
    const uint8_t* blob = ctx->d_blob;
    if (L.op == FRB_OP_STEM) {
      if (L.win != 112 || L.hin % kStemRows || L.cin != 3 || L.cout != 64 || L.w_bytes != 64 * 32 * 2)
        return fail(ctx, "stem layer must be Conv3x3(3->64) on 112-wide rows with [64][32] bf16 weights");
      if (make_tmap_2d(ctx, &pl.tmA[i], ctx->d_bufs[L.out_buf], 64, static_cast<uint64_t>(B) * L.hin * L.win, L.win)) return 1;
    } else if (L.op == FRB_OP_CONV) {
      const void* sc = L.sc_buf >= 0 ? ctx->d_bufs[L.sc_buf] : nullptr;
      const void* res = L.res_buf >= 0 ? ctx->d_bufs[L.res_buf] : nullptr;
      if (slab_eligible(ctx, L, sc != nullptr)) {
        pl.use_slab[i] = 1;
        if (setup_slab(ctx, L, B, in, res, blob + L.w_off, reinterpret_cast<const float*>(blob + L.bias_off),
                       reinterpret_cast<const float*>(blob + L.prelu_off), ctx->d_bufs[L.out_buf], &pl.tmA[i], &pl.tmB[i],
                       &pl.sp[i], &pl.slab_smem[i], &pl.grid[i], &pl.tmA2[i], &pl.tmR[i]))
          return 1;
        memset(&pl.gp[i], 0, sizeof(GemmParams));
        pl.gp[i].M = 1;  // plan marker: non-empty
        if (ctx->use_dataflow && !ctx->profiling) {
          pl.sp[i].progress = ctx->d_progress;
          pl.sp[i].wait_target = ctx->use_dataflow == 2 ? -1 : pl.wait_target[i];
          pl.sp[i].sig_fence = getenv("FRB_DF_NOFENCE") ? 0 : 1;
        }
        continue;
      }
      if (setup_conv(ctx, L, B, in, sc, res, blob + L.w_off, reinterpret_cast<const float*>(blob + L.bias_off),
                     reinterpret_cast<const float*>(blob + L.prelu_off), ctx->d_bufs[L.out_buf], &pl.tmA[i],
                     &pl.tmA2[i], &pl.tmB[i], &pl.gp[i], &pl.block_n[i], &pl.grid[i]))
        return 1;
      if (ctx->use_dataflow && !ctx->profiling && ctx->conv_mode == 2) {
        pl.gp[i].progress = ctx->d_progress;
        pl.gp[i].wait_target = ctx->use_dataflow == 2 ? -1 : pl.wait_target[i];
        pl.gp[i].sig_fence = getenv("FRB_DF_NOFENCE") ? 0 : 1;
      }
    } else if (L.op == FRB_OP_FC) {
      if (L.cin % 64 || L.cout % 256) return fail(ctx, "FC dims unsupported (%d -> %d)", L.cin, L.cout);
      GemmParams& gp = pl.gp[i];
      memset(&gp, 0, sizeof(gp));
      gp.M = B;
      gp.N = L.cout;
      gp.num_kb_main = L.cin / 64;
      gp.num_kb_sc = 0;
      const int mn_tiles = ((B + 127) / 128) * (L.cout / 256);
      // fixed split count (independent of the batch) so that a face's embedding is bit-identical
      // whatever batch it is embedded in: the fp32 summation order over K never changes
      gp.num_splits = choose_splits(gp.num_kb_main, 37);
      const size_t need = static_cast<size_t>(gp.num_splits) * B * L.cout;
      if (ctx->fc_partial_elems < need) {
        if (ctx->d_fc_partial) CK(cudaFree(ctx->d_fc_partial));
        ctx->d_fc_partial = nullptr;
        CK(cudaMalloc(&ctx->d_fc_partial, need * 4));
        ctx->fc_partial_elems = need;
      }
      gp.out_f32 = ctx->d_fc_partial;
      if (make_tmap_2d(ctx, &pl.tmA[i], in, L.cin, B, 128)) return 1;
      pl.tmA2[i] = pl.tmA[i];
      if (make_tmap_2d(ctx, &pl.tmB[i], blob + L.w_off, L.cin, L.cout, 256)) return 1;
      pl.block_n[i] = 256;
      pl.grid[i] = std::min(mn_tiles * gp.num_splits, ctx->num_sms);
    }
  }
  // split-K scratch: one partial area shared by all layers (they run one after the other), one flag region per layer
  size_t max_partial = 0, total_flags = 0;
  std::vector<size_t> flag_off(nl, 0);
  for (size_t i = 0; i < nl; ++i) {
    if (ctx->layers[i].op != FRB_OP_CONV || pl.use_slab[i]) continue;
    size_t pf, fl;
    tail_split_needs(pl.gp[i], pl.block_n[i], pl.grid[i], &pf, &fl);
    flag_off[i] = total_flags;
    total_flags += fl;
    max_partial = std::max(max_partial, pf);
  }
  // persistent runs: maximal sequences of consecutive pair-kernel conv layers with the same Cout tile
  pl.run_len.assign(nl, 0); pl.run_grid.assign(nl, 0); pl.run_index.assign(nl, -1); pl.run_off.assign(nl, 0);
  pl.num_runs = 0;
  if (ctx->conv_multi && ctx->conv_mode == 2 && !ctx->use_dataflow && ctx->num_sms >= 2) {
    auto eligible = [&](size_t i) {
      return ctx->layers[i].op == FRB_OP_CONV && !pl.use_slab[i] && !pl.gp[i].quad && pl.gp[i].tail_split <= 1 &&
             pl.grid[i] <= ctx->num_sms;
    };
    std::vector<Gemm2Layer> host;
    size_t i = 0;
    while (i < nl) {
      if (!eligible(i)) { ++i; continue; }
      size_t j = i + 1;
      const size_t max_run = getenv("FRB_MULTI_MAXRUN") ? static_cast<size_t>(atoi(getenv("FRB_MULTI_MAXRUN"))) : nl;
      while (j < nl && j - i < max_run && eligible(j) && pl.block_n[j] == pl.block_n[i]) ++j;
      if (j - i >= 2) {
        pl.run_len[i] = static_cast<int>(j - i);
        pl.run_index[i] = pl.num_runs++;
        pl.run_off[i] = host.size();
        int g = 0;
        long long units = 0;   // run-local progress units per image written by the run's layers so far
        std::vector<std::pair<const void*, long long>> out_layout;   // output buffer -> elements per image at its last write
        for (size_t q = i; q < j; ++q) {
          g = std::max(g, pl.grid[q]);
          Gemm2Layer gl;
          gl.tmA = pl.tmA[q]; gl.tmA2 = pl.tmA2[q]; gl.tmB = pl.tmB[q]; gl.p = pl.gp[q];
          if (ctx->conv_multi >= 2) {   // dataflow inside the run (FRB_MULTI=2): per-image progress instead of grid barriers
            gl.p.progress = ctx->d_progress;
            gl.p.wait_target = static_cast<int>(units);
            gl.p.sig_fence = 1;
            units += static_cast<long long>(gl.p.P) * gl.p.Q * (gl.p.N / 32);
            // stage transitions: the output buffer is re-used with another per-image size -> full wait (see kernel)
            const long long per_image = static_cast<long long>(gl.p.P) * gl.p.Q * gl.p.N;
            bool seen = false;
            for (auto& ol : out_layout)
              if (ol.first == gl.p.out) {
                seen = true;
                if (ol.second != per_image) gl.p.full_wait = 1;
                ol.second = per_image;
              }
            // first write of the run into this buffer: its old content has whatever layout the layers before the run
            // (or an earlier layer's INPUT view) used - treat as a change unless it is the run's first layer
            if (!seen) {
              out_layout.push_back({gl.p.out, per_image});
              if (q > i) gl.p.full_wait = 1;
            }
          }
          if (const char* e = getenv("FRB_MULTI_DEBUG")) gl.p.tail_debug = atoi(e);
          host.push_back(gl);
        }
        if (const char* e = getenv("FRB_MULTI_GRID")) g = std::min(g, std::max(2, atoi(e) & ~1));  // experiments: fewer SMs
        pl.run_grid[i] = g;
      }
      i = j;
    }
    if (pl.num_runs > 0) {
      if (ensure(ctx, &ctx->d_runs, &ctx->runs_cap, host.size())) return 1;
      if (ensure(ctx, &ctx->d_run_bar, &ctx->run_bar_cap, host.size())) return 1;   // one barrier counter per layer
      pl.run_layers = host.size();
      CK(cudaMemcpy(ctx->d_runs, host.data(), host.size() * sizeof(Gemm2Layer), cudaMemcpyHostToDevice));
    }
  }
  // persistent runs of identical slab layers with resident weights (stage 2 of IR-101: 24 layers; stage 1: 4)
  pl.srun_len.assign(nl, 0); pl.srun_off.assign(nl, 0);
  if (ctx->conv_multi >= 2 && ctx->conv_mode == 2 && !ctx->use_dataflow && ctx->slab_multi) {
    auto same = [&](size_t a, size_t b) {
      const frb_layer_desc &A = ctx->layers[a], &Bq = ctx->layers[b];
      return A.cin == Bq.cin && A.cout == Bq.cout && A.hin == Bq.hin && A.win == Bq.win && pl.sp[a].nbuf == pl.sp[b].nbuf &&
             pl.sp[a].slab_bytes == pl.sp[b].slab_bytes && pl.slab_smem[a] == pl.slab_smem[b] && pl.grid[a] == pl.grid[b];
    };
    auto eligible = [&](size_t i) {
      return ctx->layers[i].op == FRB_OP_CONV && pl.use_slab[i] && pl.sp[i].b_stages == 9 * (ctx->layers[i].cin / 64) &&
             pl.sp[i].debug == 0 && pl.sp[i].trace == nullptr && pl.grid[i] <= ctx->num_sms;
    };
    std::vector<SlabLayer> host;
    size_t i = 0;
    while (i < nl) {
      if (!eligible(i)) { ++i; continue; }
      size_t j = i + 1;
      while (j < nl && eligible(j) && same(i, j)) ++j;
      if (j - i >= 2) {
        pl.srun_len[i] = static_cast<int>(j - i);
        pl.srun_off[i] = host.size();
        long long units = 0;
        for (size_t q = i; q < j; ++q) {
          SlabLayer sl;
          sl.tmX = pl.tmA[q]; sl.tmB = pl.tmB[q]; sl.p = pl.sp[q];
          sl.p.progress = ctx->d_progress;
          sl.p.wait_target = static_cast<int>(units);
          sl.p.sig_fence = 1;
          units += static_cast<long long>(sl.p.H) * sl.p.W * (sl.p.N / 32);
          host.push_back(sl);
        }
      }
      i = j;
    }
    if (!host.empty()) {
      if (ensure(ctx, &ctx->d_sruns, &ctx->sruns_cap, host.size())) return 1;
      CK(cudaMemcpy(ctx->d_sruns, host.data(), host.size() * sizeof(SlabLayer), cudaMemcpyHostToDevice));
    }
  }
  pl.tail_flags = total_flags;
  if (total_flags > 0) {
    if (ensure(ctx, &ctx->d_tail_partial, &ctx->tail_partial_cap, max_partial)) return 1;
    if (ensure(ctx, &ctx->d_tail_flags, &ctx->tail_flags_cap, total_flags)) return 1;
    for (size_t i = 0; i < nl; ++i)
      if (ctx->layers[i].op == FRB_OP_CONV && !pl.use_slab[i] && pl.gp[i].tail_split > 1) {
        pl.gp[i].tail_partial = ctx->d_tail_partial;
        pl.gp[i].tail_flag = ctx->d_tail_flags + flag_off[i];
      }
  }
  return 0;
}

int launch_stem(frb_ctx* ctx, const frb_layer_desc& L, const CUtensorMap& tm_out, const void* in, int B, int* progress, cudaStream_t st) {
  const uint8_t* blob = ctx->d_blob;
  if (set_smem_attr(ctx, reinterpret_cast<const void*>(stem_tc_kernel), kStemSmemBytes)) return 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(B * (L.hin / kStemRows));
  cfg.blockDim = dim3(kStemThreads);
  cfg.dynamicSmemBytes = kStemSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  fill_launch_attrs(attr, 1);
  cfg.attrs = attr + 1;
  cfg.numAttrs = ctx->use_pdl ? 1 : 0;
  CK(cudaLaunchKernelEx(&cfg, stem_tc_kernel, tm_out, reinterpret_cast<const __nv_bfloat16*>(in),
                        reinterpret_cast<const __nv_bfloat16*>(blob + L.w_off), reinterpret_cast<const float*>(blob + L.bias_off),
                        reinterpret_cast<const float*>(blob + L.prelu_off), static_cast<int>(L.hin), static_cast<int>(L.win), progress));
  ctx->launches++;
  return 0;
}

// How many leading layers form the front: stem + conv layers on >= 56-pixel inputs with Cout <= 64, provided the buffer
// the front ends in is used with ONE per-image size inside the front (it is addressed at the sub-batch's image offset).
int front_layers(const frb_ctx* ctx) {
  const auto& Ls = ctx->layers;
  size_t nf = 0;
  while (nf < Ls.size() && (Ls[nf].op == FRB_OP_STEM || (Ls[nf].op == FRB_OP_CONV && Ls[nf].hin >= 56 && Ls[nf].cout <= 64))) ++nf;
  if (nf < 2 || nf >= Ls.size() || Ls[0].op != FRB_OP_STEM) return 0;
  const int ext = Ls[nf - 1].out_buf;
  const auto& F = Ls[nf - 1];
  const long long want = static_cast<long long>(out_dim(F.hin, F.ksize, F.stride, F.pad)) * out_dim(F.win, F.ksize, F.stride, F.pad) * F.cout;
  for (size_t i = 0; i < nf; ++i) {
    const auto& L = Ls[i];
    const long long out_e = static_cast<long long>(out_dim(L.hin, L.ksize, L.stride, L.pad)) * out_dim(L.win, L.ksize, L.stride, L.pad) * L.cout;
    if (L.out_buf == ext && out_e != want) return 0;
    if (L.in_buf == ext && static_cast<long long>(L.hin) * L.win * L.cin != want) return 0;
    if (L.sc_buf == ext && static_cast<long long>(L.sc_hin) * L.sc_win * L.sc_cin != want) return 0;
    if (L.res_buf == ext && static_cast<long long>(L.res_h) * L.res_w * L.cout != want) return 0;
  }
  return static_cast<int>(nf);
}

int build_front(frb_ctx* ctx, int B) {
  FrontPlan& fp = ctx->front;
  fp = FrontPlan();
  const int nf = front_layers(ctx);
  if (nf == 0 || ctx->front_sub <= 0 || B <= ctx->front_sub || ctx->conv_mode != 2) return 0;   // fp.nf == 0: no front
  fp.B = B; fp.nf = nf;
  fp.n_sub = (B + ctx->front_sub - 1) / ctx->front_sub;
  fp.sub = (B + fp.n_sub - 1) / fp.n_sub;
  const int ext = ctx->layers[nf - 1].out_buf;
  fp.use_slab.assign(nf, 0); fp.block_n.assign(nf, 0); fp.grid.assign(nf, 0); fp.slab_smem.assign(nf, 0);
  fp.tmA.assign(fp.n_sub, std::vector<CUtensorMap>(nf)); fp.tmA2 = fp.tmA; fp.tmB = fp.tmA;
  fp.gp.assign(fp.n_sub, std::vector<GemmParams>(nf)); fp.sp.assign(fp.n_sub, std::vector<SlabParams>(nf));
  const uint8_t* blob = ctx->d_blob;
  for (int c = 0; c < fp.n_sub; ++c) {
    const int c0 = std::min(c * fp.sub, B - fp.sub);
    fp.c0.push_back(c0);
    // buffer `id` as this sub-batch sees it: the front's final buffer at the sub-batch's images, everything else from 0
    auto buf = [&](int id, long long per_image) -> __nv_bfloat16* {
      if (id < 0) return nullptr;
      return ctx->d_bufs[id] + (id == ext ? static_cast<size_t>(c0) * per_image : 0);
    };
    for (int i = 0; i < nf; ++i) {
      const frb_layer_desc& L = ctx->layers[i];
      const long long out_e = static_cast<long long>(out_dim(L.hin, L.ksize, L.stride, L.pad)) * out_dim(L.win, L.ksize, L.stride, L.pad) * L.cout;
      __nv_bfloat16* out = buf(L.out_buf, out_e);
      if (L.op == FRB_OP_STEM) {
        if (make_tmap_2d(ctx, &fp.tmA[c][i], out, 64, static_cast<uint64_t>(fp.sub) * L.hin * L.win, L.win)) return 1;
        continue;
      }
      const void* in = buf(L.in_buf, static_cast<long long>(L.hin) * L.win * L.cin);
      const void* sc = buf(L.sc_buf, static_cast<long long>(L.sc_hin) * L.sc_win * L.sc_cin);
      const void* res = buf(L.res_buf, static_cast<long long>(L.res_h) * L.res_w * L.cout);
      const float* bias = reinterpret_cast<const float*>(blob + L.bias_off);
      const float* prelu = reinterpret_cast<const float*>(blob + L.prelu_off);
      if (slab_eligible(ctx, L, sc != nullptr)) {
        fp.use_slab[i] = 1;
        if (setup_slab(ctx, L, fp.sub, in, res, blob + L.w_off, bias, prelu, out, &fp.tmA[c][i], &fp.tmB[c][i], &fp.sp[c][i],
                       &fp.slab_smem[i], &fp.grid[i], &fp.tmA2[c][i]))
          return 1;
      } else {
        if (setup_conv(ctx, L, fp.sub, in, sc, res, blob + L.w_off, bias, prelu, out, &fp.tmA[c][i], &fp.tmA2[c][i], &fp.tmB[c][i],
                       &fp.gp[c][i], &fp.block_n[i], &fp.grid[i]))
          return 1;
        if (fp.gp[c][i].tail_split > 1 || fp.gp[c][i].quad) { fp = FrontPlan(); return 0; }   // experiments only: no front then
      }
    }
  }
  return 0;
}

// Bn faces through the network in ONE pass of the layer program (no flip pairing here: l2 / renorm as given).
int embed_chunk_locked(frb_ctx* ctx, const void* d_in, int Bn, int l2, int renorm, float* d_emb, float* d_norm,
                       void* d_emb_bf16, cudaStream_t st) {
  const bool want_df = ctx->use_dataflow && !ctx->profiling;
  // the plan bakes in the batch size only: the network input is read by the stem kernel through a plain pointer
  if (ctx->plan.B != Bn || ctx->plan.gp.empty() || ctx->plan.dataflow != want_df) {
    ctx->plan.gp.clear();
    if (build_plan(ctx, Bn, d_in)) return 1;
    ctx->plan.d_in = d_in;
    ctx->plan.dataflow = want_df;
    ctx->front = FrontPlan();
    if (!want_df && build_front(ctx, Bn)) return 1;
  }
  Plan& pl = ctx->plan;
  const FrontPlan& fp = ctx->front;
  const size_t face_bytes = static_cast<size_t>(112) * 112 * 3 * 2;
  const bool dataflow = pl.dataflow && ctx->conv_mode == 2;
  if (dataflow) CK(cudaMemsetAsync(ctx->d_progress, 0, sizeof(int) * Bn, st));
  if (pl.tail_flags > 0) CK(cudaMemsetAsync(ctx->d_tail_flags, 0, sizeof(int) * pl.tail_flags, st));
  if (pl.num_runs > 0) CK(cudaMemsetAsync(ctx->d_run_bar, 0, sizeof(int) * pl.run_layers, st));
  size_t run_end = 0;  // layers below this index were covered by a persistent run launch
  // frb_embed_profile: an event before every launch, tagged with its layer (a front layer is launched once per sub-batch)
  ctx->prof_marks.clear();
  auto mark = [&](int layer) -> int {
    if (!ctx->profiling) return 0;
    const size_t k = ctx->prof_marks.size();
    while (ctx->prof_events.size() <= k) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      ctx->prof_events.push_back(e);
    }
    CK(cudaEventRecord(ctx->prof_events[k], st));
    ctx->prof_marks.push_back(layer);
    return 0;
  };
  // ---- the front, sub-batch by sub-batch
  for (int c = 0; c < (fp.nf > 0 ? fp.n_sub : 0); ++c)
    for (int i = 0; i < fp.nf; ++i) {
      const frb_layer_desc& L = ctx->layers[i];
      if (mark(i)) return 1;
      if (L.op == FRB_OP_STEM) {
        if (launch_stem(ctx, L, fp.tmA[c][i], static_cast<const uint8_t*>(d_in) + fp.c0[c] * face_bytes, fp.sub, nullptr, st)) return 1;
      } else if (fp.use_slab[i]) {
        if (launch_slab(ctx, L.cin / 64, fp.tmA[c][i], fp.tmB[c][i], fp.tmA2[c][i], fp.tmB[c][i], fp.sp[c][i], fp.slab_smem[i], fp.grid[i], st)) return 1;
      } else if (launch_conv(ctx, fp.block_n[i], fp.tmA[c][i], fp.tmA2[c][i], fp.tmB[c][i], fp.gp[c][i], fp.grid[i], st)) return 1;
    }
  for (size_t i = static_cast<size_t>(fp.nf); i < ctx->layers.size(); ++i) {
    const frb_layer_desc& L = ctx->layers[i];
    const uint8_t* blob = ctx->d_blob;
    if (i < run_end) continue;
    if (mark(static_cast<int>(i))) return 1;
    if (pl.srun_len[i] > 1) {
      CK(cudaMemsetAsync(ctx->d_progress, 0, sizeof(int) * Bn, st));  // run-local progress
      if (launch_slab_multi(ctx, L.cin / 64, L.cout, ctx->d_sruns + pl.srun_off[i], pl.srun_len[i], pl.slab_smem[i], pl.grid[i], st))
        return 1;
      run_end = i + pl.srun_len[i];
      continue;
    }
    if (pl.run_len[i] > 1) {
      if (ctx->conv_multi >= 2) CK(cudaMemsetAsync(ctx->d_progress, 0, sizeof(int) * Bn, st));  // run-local progress
      if (launch_gemm2_multi(ctx, pl.block_n[i], ctx->d_runs + pl.run_off[i], pl.run_len[i], ctx->d_run_bar + pl.run_off[i],
                             pl.run_grid[i], st))
        return 1;
      run_end = i + pl.run_len[i];
      continue;
    }
    if (L.op == FRB_OP_STEM) {
      const void* in = L.in_buf < 0 ? d_in : ctx->d_bufs[L.in_buf];
      if (launch_stem(ctx, L, pl.tmA[i], in, Bn, dataflow ? ctx->d_progress : static_cast<int*>(nullptr), st)) return 1;
    } else if (L.op == FRB_OP_CONV) {
      if (pl.use_slab[i]) {
        if (launch_slab(ctx, L.cin / 64, pl.tmA[i], pl.tmB[i], pl.tmA2[i], pl.tmR[i], pl.sp[i], pl.slab_smem[i], pl.grid[i], st)) return 1;
      } else if (launch_conv(ctx, pl.block_n[i], pl.tmA[i], pl.tmA2[i], pl.tmB[i], pl.gp[i], pl.grid[i], st)) return 1;
    } else if (L.op == FRB_OP_FC) {
      if (launch_gemm(ctx, 256, A_TILED, 1, pl.tmA[i], pl.tmA2[i], pl.tmB[i], pl.gp[i], pl.grid[i], st)) return 1;
      fc_finalize_kernel<<<Bn, 128, 0, st>>>(ctx->d_fc_partial, pl.gp[i].num_splits, Bn,
                                             reinterpret_cast<const float*>(blob + L.bias_off), l2, renorm, d_emb, d_norm,
                                             reinterpret_cast<__nv_bfloat16*>(d_emb_bf16));
      CK(cudaGetLastError());
      ctx->launches++;
    }
  }
  if (mark(-1)) return 1;   // closing event
  return 0;
}

// B faces (2B crops with FRB_EMBED_FLIP).  Large batches go through the network in chunks of at most
// ctx->embed_chunk faces (default 256; FRB_EMBED_CHUNK=0 disables): per face the network is FASTER at 256 than at
// 1024 (21.3 vs 24.4 us measured) because a layer's output stays in the 126 MB L2 for the next layer only while
// the activations of a chunk fit there.  A face's embedding does not depend on the batch it travels in
// (tests/test_gpu_embed.py), so chunking changes no bit; chunks are equal-sized (one plan), the last one is
// right-aligned and recomputes a few faces of its predecessor when the batch is not a multiple.
int embed_locked(frb_ctx* ctx, const void* d_in, int B, int flags, float* d_emb, float* d_norm, void* d_emb_bf16,
                 cudaStream_t st) {
  if (ctx->layers.empty()) return fail(ctx, "frb_embed: no backbone loaded");
  const bool flipf = (flags & FRB_EMBED_FLIP) != 0;
  const int Bn = flipf ? 2 * B : B;  // faces through the network
  float* emb_dst = d_emb;
  float* norm_dst = d_norm;
  void* bf_dst = d_emb_bf16;
  if (flipf) {
    if (ensure(ctx, &ctx->d_emb2, &ctx->emb2_elems, static_cast<size_t>(Bn) * 512)) return 1;
    emb_dst = ctx->d_emb2;
    norm_dst = nullptr;
    bf_dst = nullptr;
  }
  const int l2 = (flags & FRB_EMBED_L2) ? 1 : 0;
  // with flip fusion each half is normalised the way extract_embeddings_batch(normalize=True) does
  const int renorm = ((flags & FRB_EMBED_RENORM) || flipf) ? 1 : 0;
  const int limit = (ctx->profiling || ctx->embed_chunk <= 0) ? Bn : ctx->embed_chunk;
  const int n_chunks = (Bn + limit - 1) / limit;
  const int Bc = (Bn + n_chunks - 1) / n_chunks;
  const size_t face_bytes = static_cast<size_t>(112) * 112 * 3 * 2;
  for (int c = 0; c < n_chunks; ++c) {
    const int c0 = std::min(c * Bc, Bn - Bc);   // the last chunk is right-aligned
    if (embed_chunk_locked(ctx, static_cast<const uint8_t*>(d_in) + c0 * face_bytes, Bc, l2, renorm,
                           emb_dst ? emb_dst + static_cast<size_t>(c0) * 512 : nullptr, norm_dst ? norm_dst + c0 : nullptr,
                           bf_dst ? static_cast<uint8_t*>(bf_dst) + static_cast<size_t>(c0) * 512 * 2 : nullptr, st))
      return 1;
  }
  if (flipf) {
    flip_fuse_kernel<<<B, 128, 0, st>>>(ctx->d_emb2, B, d_emb, reinterpret_cast<__nv_bfloat16*>(d_emb_bf16));
    CK(cudaGetLastError());
    ctx->launches++;
  }
  return 0;
}

}  // namespace

// Profiling variant of frb_embed: the same launches with a CUDA event between consecutive layers (which also
// serialises them: no programmatic overlap), so per-layer durations can be read back.  kernel_id: 0 stem,
// 1/2/3 = CTA-pair im2col conv with Cout tile 64/128/256, 4/5/6 = slab conv <64,1>/<128,1>/<128,2>, 7 = FC + finalize,
// 8 = layer of a persistent multi-layer run (gemm2_multi_sm100_kernel): the whole run's time is reported on its first
// layer, the following layers of the run are reported with id -8 and ~0 ms.
extern "C" int frb_embed_profile(frb_ctx* ctx, const void* d_in, int B, int flags, float* d_emb, void* stream,
                                 float* h_layer_ms, int* h_kernel_id, double* h_layer_flops, int max_layers,
                                 int* n_layers) {
  if (!ctx) return 1;
  if (B <= 0) return 0;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ws_begin(ctx, st)) return 1;
  ctx->profiling = true;
  const int rc = embed_locked(ctx, d_in, B, flags, d_emb, nullptr, nullptr, st);
  ctx->profiling = false;
  if (rc) return rc;
  CK(cudaStreamSynchronize(st));
  const int nl = static_cast<int>(ctx->layers.size());
  const int Bn = (flags & FRB_EMBED_FLIP) ? 2 * B : B;
  if (n_layers) *n_layers = nl;
  std::vector<float> layer_ms(nl, 0.f);
  for (size_t k = 0; k + 1 < ctx->prof_marks.size(); ++k) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->prof_events[k], ctx->prof_events[k + 1]));
    if (ctx->prof_marks[k] >= 0 && ctx->prof_marks[k] < nl) layer_ms[ctx->prof_marks[k]] += ms;
  }
  for (int i = 0; i < nl && i < max_layers; ++i) {
    const frb_layer_desc& L = ctx->layers[i];
    if (h_layer_ms) h_layer_ms[i] = layer_ms[i];
    int id = 7;
    if (L.op == FRB_OP_STEM) id = 0;
    else if (L.op == FRB_OP_CONV) {
      if (ctx->plan.use_slab[i]) id = L.cout == 64 ? 4 : (L.cin == 64 ? 5 : 6);
      else id = ctx->plan.block_n[i] == 64 ? 1 : (ctx->plan.block_n[i] == 128 ? 2 : 3);
    }
    // layers inside a persistent run: the run's time is on its first layer (the others follow it without a gap)
    for (int q = i; q >= 0; --q)
      if (ctx->plan.run_len[q] > 1) {
        if (q + ctx->plan.run_len[q] > i) id = (q == i) ? 8 : -8;   // -8: continuation, not a launch of its own
        break;
      }
    for (int q = i; q >= 0; --q)
      if (ctx->plan.srun_len[q] > 1) {
        if (q + ctx->plan.srun_len[q] > i) id = (q == i) ? 9 : -9;   // persistent slab run
        break;
      }
    if (h_kernel_id) h_kernel_id[i] = id;
    const int P = out_dim(L.hin, L.ksize, L.stride, L.pad), Q = out_dim(L.win, L.ksize, L.stride, L.pad);
    double fl = L.op == FRB_OP_FC ? 2.0 * L.cin * L.cout
                                  : 2.0 * P * Q * L.cout * (static_cast<double>(L.ksize) * L.ksize * L.cin + (L.sc_buf >= 0 ? L.sc_cin : 0));
    if (h_layer_flops) h_layer_flops[i] = fl * Bn;
  }
  return 0;
}

extern "C" int frb_embed(frb_ctx* ctx, const void* d_in, int B, int flags, float* d_emb, float* d_norm,
                         void* d_emb_bf16, void* stream) {
  if (!ctx) return 1;
  if (B <= 0) return 0;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ws_begin(ctx, st)) return 1;
  if (embed_locked(ctx, d_in, B, flags, d_emb, d_norm, d_emb_bf16, st)) return 1;
  return ws_end(ctx, st);
}

// ====================================================================== gallery + match
extern "C" int frb_gallery_upload(frb_ctx* ctx, const float* g, long long N, long long first_global_id, int is_device) {
  if (!ctx) return 1;
  if (N < 0) return fail(ctx, "frb_gallery_upload: negative N");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  CK(cudaDeviceSynchronize());
  if (N > ctx->gal_cap || (ctx->gal_cap > (1ll << 20) && N < ctx->gal_cap / 8)) {   // grow, or give a much larger old gallery back
    if (ctx->d_gal) CK(cudaFree(ctx->d_gal));
    if (ctx->d_gal_bf16) CK(cudaFree(ctx->d_gal_bf16));
    ctx->d_gal = nullptr;
    ctx->d_gal_bf16 = nullptr;
    const long long cap = (std::max<long long>(N, 256) + 255) / 256 * 256;  // whole 128-row groups (K-blocked bf16 copy)
    CK(cudaMalloc(&ctx->d_gal, static_cast<size_t>(cap) * 512 * 4));
    CK(cudaMalloc(&ctx->d_gal_bf16, static_cast<size_t>(cap) * 512 * 2));
    ctx->gal_cap = cap;
  }
  ctx->gallery_gen++;
  ctx->gal_N = N;
  ctx->gal_S = 0;  // a plain template gallery: per-identity queries need frb_gallery_upload_samples
  ctx->gal_first = first_global_id;
  CK(cudaMemset(ctx->d_gal_maxnorm, 0, 8));
  if (N == 0) return 0;
  CK(cudaMemcpy(ctx->d_gal, g, static_cast<size_t>(N) * 512 * 4, is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
  const long long groups = (N + kGalGroup - 1) / kGalGroup;
  gallery_prepare_kernel<<<static_cast<unsigned>(groups * kGalGroup / 8), 256>>>(ctx->d_gal, N, ctx->d_gal_bf16, ctx->d_gal_maxnorm);
  CK(cudaGetLastError());
  ctx->launches++;
  CK(cudaDeviceSynchronize());
  // K-blocked copy as a [groups * 8 * 128][64] tensor: one box = one K block of one 128-row group = 16 KB contiguous
  if (make_tmap_2d(ctx, &ctx->tmG2, ctx->d_gal_bf16, 64, static_cast<uint64_t>(groups) * kMatchKB * kGalGroup, kGalGroup)) return 1;
  ctx->tmG = ctx->tmG2;
  return 0;
}

// Batched enrollment aggregation: templates of S identities from their embeddings, on the device.
extern "C" int frb_aggregate_templates(frb_ctx* ctx, const float* d_emb, const long long* d_seg, int S, int max_rows,
                                       int method, float min_similarity, float* d_templates, int* d_kept, void* stream) {
  if (!ctx) return 1;
  if (S <= 0) return 0;
  if (method < 0 || method > 2) return fail(ctx, "frb_aggregate_templates: method must be 0 (mean), 1 (median) or 2 (weighted_mean)");
  if (max_rows > kAggMaxRows) return fail(ctx, "frb_aggregate_templates: at most %d embeddings per identity (got %d)", kAggMaxRows, max_rows);
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  aggregate_templates_kernel<<<S, 128, 0, static_cast<cudaStream_t>(stream)>>>(d_emb, d_seg, method, min_similarity, d_templates, d_kept);
  CK(cudaGetLastError());
  ctx->launches++;
  return 0;
}

// Track-level consensus: FaceMatcher._aggregate_matches / _get_best_candidate for T tracks at once.
extern "C" int frb_track_consensus(frb_ctx* ctx, const long long* d_top_idx, const float* d_top_score, int k_stride,
                                   const long long* d_seg, int T, int max_frames, double min_quality, int min_frames,
                                   double threshold, frb_track_result* d_out, void* stream) {
  if (!ctx) return 1;
  if (T <= 0) return 0;
  static_assert(sizeof(frb_track_result) == sizeof(frb_track_result_dev), "frb_track_result layout");
  if (k_stride < 1) return fail(ctx, "frb_track_consensus: k_stride must be >= 1");
  if (max_frames > kTrackMaxFrames) return fail(ctx, "frb_track_consensus: at most %d frames per track (got %d)", kTrackMaxFrames, max_frames);
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  track_consensus_kernel<<<T, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      d_top_idx, d_top_score, k_stride, d_seg, min_quality, min_frames, threshold, reinterpret_cast<frb_track_result_dev*>(d_out));
  CK(cudaGetLastError());
  ctx->launches++;
  return 0;
}

// Server best-frame selection (second half): arg-max of det * min(blur / 100, 1) per track.
extern "C" int frb_best_frames(frb_ctx* ctx, const double* d_det, const double* d_blur, const long long* d_seg, int T,
                               double min_det, long long* d_best_idx, double* d_best_quality, unsigned char* d_ready,
                               void* stream) {
  if (!ctx) return 1;
  if (T <= 0) return 0;
  if (!d_det || !d_blur || !d_seg || !d_best_idx || !d_ready) return fail(ctx, "frb_best_frames: null argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  best_frames_kernel<<<(T + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(d_det, d_blur, d_seg, T, min_det, d_best_idx,
                                                                             d_best_quality, d_ready);
  CK(cudaGetLastError());
  ctx->launches++;
  return 0;
}

extern "C" long long frb_gallery_size(frb_ctx* ctx) { return ctx ? ctx->gal_N : 0; }
// waits for the last match on this ctx (only for the 4-byte count behind its event)
extern "C" int frb_match_last_flagged(frb_ctx* ctx) {
  if (!ctx) return 0;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (ctx->flag_event_pending) {
    cudaSetDevice(ctx->device);
    if (cudaEventSynchronize(ctx->flag_event) != cudaSuccess) return -1;
    ctx->flag_event_pending = false;
    ctx->last_flagged = ctx->h_flag_count[0];
  }
  return ctx->last_flagged;
}
extern "C" long long frb_gallery_generation(frb_ctx* ctx) { return ctx ? ctx->gallery_gen : 0; }
extern "C" long long frb_backbone_generation(frb_ctx* ctx) { return ctx ? ctx->backbone_gen : 0; }

namespace {

constexpr long long kExactOnlyBelow = 4096;  // tiny galleries: the exact scan IS the match
constexpr int kExactChunk = 16;              // probes per dense exact-scan pass

// Calls on one ctx share its workspaces (probe copies, candidate lists, staging buffers, activation buffers).  The
// mutex serialises the enqueue; this orders the WORK when consecutive calls use different streams (the host entry
// points run on own_stream, frb_match / frb_embed on the caller's): the new stream waits for the previous call's tail.
bool stream_capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return cs != cudaStreamCaptureStatusNone;
}
// (a stream that is being captured into a CUDA graph may not depend on work outside the capture: the caller of a
// captured call orders the graph launch against earlier work on this ctx itself)
int ws_begin(frb_ctx* ctx, cudaStream_t st) {
  if (ctx->ws_valid && ctx->ws_stream != st && !stream_capturing(st)) CK(cudaStreamWaitEvent(st, ctx->ws_event, 0));
  return 0;
}
int ws_end(frb_ctx* ctx, cudaStream_t st) {
  if (stream_capturing(st)) {
    ctx->ws_valid = false;
    return 0;
  }
  if (!ctx->ws_event) CK(cudaEventCreateWithFlags(&ctx->ws_event, cudaEventDisableTiming));
  CK(cudaEventRecord(ctx->ws_event, st));
  ctx->ws_stream = st;
  ctx->ws_valid = true;
  return 0;
}

// dense exact scan (host-known row count): tiny galleries and k > kExactMaxK
int run_exact(frb_ctx* ctx, const float* d_probes_norm, int F, int k, float thr, float* d_scores,
              long long* d_idx, unsigned char* d_accept, double* d_scores64, cudaStream_t st) {
  const long long N = ctx->gal_N;
  for (int f0 = 0; f0 < F; f0 += kExactChunk) {
    const int fc = std::min(kExactChunk, F - f0);
    const size_t need = static_cast<size_t>(fc) * N;
    if (ensure(ctx, &ctx->d_exact, &ctx->exact_elems, std::max<size_t>(need, 1))) return 1;
    const float* probes = d_probes_norm + static_cast<size_t>(f0) * 512;
    dim3 grid(static_cast<unsigned>(std::min<long long>((N + 7) / 8, 148 * 8)), fc);
    match_exact_scores_kernel<<<grid, 256, 0, st>>>(ctx->d_gal, N, probes, nullptr, ctx->d_exact);
    CK(cudaGetLastError());
    ctx->launches++;
    const size_t oo = static_cast<size_t>(f0);
    match_exact_topk_kernel<<<fc, 256, 0, st>>>(ctx->d_exact, N, nullptr, k, thr, ctx->gal_first, d_scores64 + oo * k,
                                                d_idx + oo * k, d_scores + oo * k, d_accept + oo);
    CK(cudaGetLastError());
    ctx->launches++;
  }
  return 0;
}

int match_workspace(frb_ctx* ctx, int P) {
  if (!ctx->d_match_ctr) {
    CK(cudaMalloc(&ctx->d_match_ctr, 64));
    CK(cudaMemset(ctx->d_match_ctr, 0, 64));
    CK(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_flag_count), 64, cudaHostAllocDefault));
    ctx->h_flag_count[0] = 0;
    CK(cudaEventCreateWithFlags(&ctx->flag_event, cudaEventDisableTiming));
  }
  if (ctx->match_cap_P < P) {
    if (ctx->d_probe_f32) CK(cudaFree(ctx->d_probe_f32));
    if (ctx->d_probe_bf16) CK(cudaFree(ctx->d_probe_bf16));
    if (ctx->d_flagged) CK(cudaFree(ctx->d_flagged));
    if (ctx->d_flag_rows) CK(cudaFree(ctx->d_flag_rows));
    if (ctx->d_row_floor) CK(cudaFree(ctx->d_row_floor));
    if (ctx->d_cand_score) CK(cudaFree(ctx->d_cand_score));
    if (ctx->d_cand_idx) CK(cudaFree(ctx->d_cand_idx));
    ctx->d_probe_f32 = nullptr; ctx->d_probe_bf16 = nullptr; ctx->d_flagged = nullptr; ctx->d_flag_rows = nullptr;
    ctx->d_row_floor = nullptr;
    ctx->d_cand_score = nullptr; ctx->d_cand_idx = nullptr; ctx->match_cap_slices = 0;
    const int cap = std::max(P, 128);
    CK(cudaMalloc(&ctx->d_probe_f32, static_cast<size_t>(cap) * 512 * 4));
    CK(cudaMalloc(&ctx->d_probe_bf16, static_cast<size_t>(cap) * 512 * 2));
    CK(cudaMalloc(&ctx->d_flagged, static_cast<size_t>(cap) * 4));
    CK(cudaMalloc(&ctx->d_flag_rows, static_cast<size_t>(cap) * 4));
    CK(cudaMalloc(&ctx->d_row_floor, static_cast<size_t>(cap) * 4));
    ctx->match_cap_P = cap;
  }
  return 0;
}

// Match P prepared probes (fp32 + bf16 copies, normalised as search() does) against the resident gallery.  Everything
// is enqueued on `st`; nothing here waits for the device.  push != nullptr: identity-sharded match, finished rows
// also go to the peers' exchange buffers.
// The caller has zeroed the match counters (ctx->d_match_ctr[0..1]) on `st` BEFORE its probe kernel: nothing but
// kernels may sit between the probe kernel and the filter, or the programmatic launch edge between them is lost.
int match_core(frb_ctx* ctx, const float* d_probe_f32, const __nv_bfloat16* d_probe_bf16, int P, int k, float thr,
               float* d_scores, long long* d_idx, unsigned char* d_accept, double* s64, const PeerPush* push,
               cudaStream_t st) {
  const long long N = ctx->gal_N;
  PeerPush no_push;
  memset(&no_push, 0, sizeof(no_push));
  const PeerPush& pp = push ? *push : no_push;
  if (N == 0 || N < kExactOnlyBelow || k > kExactMaxK) {
    if (N == 0) {
      match_fill_empty_kernel<<<(P * k + 255) / 256, 256, 0, st>>>(P, k, s64, d_idx, d_scores, d_accept);
      CK(cudaGetLastError());
      ctx->launches++;
    } else if (run_exact(ctx, d_probe_f32, P, k, thr, d_scores, d_idx, d_accept, s64, st)) {
      return 1;
    }
    if (pp.world > 0) {
      xchg_push_rows_kernel<<<(P + 31) / 32, 128, 0, st>>>(s64, d_idx, P, k, pp);
      CK(cudaGetLastError());
      ctx->launches++;
    }
    ctx->last_flagged = 0;
    ctx->flag_event_pending = false;
    return 0;
  }
  // ---- bf16 tensor-core filter ----
  MatchParams mp;
  mp.P = P;
  mp.N = N;
  mp.prefetch_tiles = ctx->match_prefetch;
  const bool pair_mode = P > 128 && ctx->match_pair;   // two probe tiles per CTA pair (match_filter2_kernel)
  const int units = pair_mode ? ctx->num_sms / 2 : ctx->num_sms;  // concurrently running work items
  mp.p_tiles = pair_mode ? (P + 255) / 256 : (P + 127) / 128;
  mp.g_tiles = static_cast<int>((N + kMatchBN - 1) / kMatchBN);
  // Slices: long enough that the per-row top-k lists settle (few insertions => cheap epilogue), and
  // p_tiles * slices close to a multiple of the number of concurrent work items (whole waves).
  {
    const int smin = std::max(1, (units + mp.p_tiles - 1) / mp.p_tiles);
    const int smax = std::max(1, std::min(std::min(mp.g_tiles, kMaxCandPad / kCand), smin * 8));
    long best_cost = -1;
    int best_s = 1;
    for (int sl = std::min(smin, smax); sl <= smax; ++sl) {
      const int tps = (mp.g_tiles + sl - 1) / sl;
      const int real = (mp.g_tiles + tps - 1) / tps;
      const long waves = (static_cast<long>(mp.p_tiles) * real + units - 1) / units;
      const long cost = waves * (tps + 2);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best_s = real;
      }
    }
    if (const char* e = getenv("FRB_MATCH_SLICES")) best_s = std::max(1, std::min(atoi(e), smax));
    mp.tiles_per_slice = (mp.g_tiles + best_s - 1) / best_s;
    mp.slices = (mp.g_tiles + mp.tiles_per_slice - 1) / mp.tiles_per_slice;
  }
  if (ctx->match_cap_slices < mp.slices || !ctx->d_cand_score) {
    if (ctx->d_cand_score) CK(cudaFree(ctx->d_cand_score));
    if (ctx->d_cand_idx) CK(cudaFree(ctx->d_cand_idx));
    ctx->d_cand_score = nullptr; ctx->d_cand_idx = nullptr;
    const size_t n = static_cast<size_t>(ctx->match_cap_P) * mp.slices * kCand;
    CK(cudaMalloc(&ctx->d_cand_score, n * 4));
    CK(cudaMalloc(&ctx->d_cand_idx, n * 4));
    ctx->match_cap_slices = mp.slices;
  }
  if (ensure(ctx, &ctx->d_exact_part, &ctx->exact_part_cap, static_cast<size_t>(std::max(P, kExactRowsY * kExactRowsY)) * kExactBlocks * k)) return 1;
  mp.cand_score = ctx->d_cand_score;
  mp.cand_idx = ctx->d_cand_idx;
  mp.row_floor = ctx->d_row_floor;
  CUtensorMap tmP;
  if (make_tmap_2d(ctx, &tmP, d_probe_bf16, 512, static_cast<uint64_t>(P), 128)) return 1;
  if (ctx->match_profiling) CK(cudaEventRecord(ctx->match_prof_ev[1], st));
  // The four kernels of the chain are launched with programmatic stream serialization: each sets itself up while its
  // predecessor drains and blocks in griddepcontrol.wait before it reads the predecessor's output.  In the plain
  // match the filter's predecessor is probe_prepare_kernel (which releases its dependents at once); in the sharded
  // match it is xchg_wait_kernel (no early release: the filter waits for its completion).
  const bool pdl = ctx->use_pdl != 0 && !ctx->match_profiling;   // an event between two launches breaks the programmatic edge
  if (pair_mode) {
    if (set_smem_attr(ctx, reinterpret_cast<const void*>(match_filter2_kernel), Match2Smem::kTotal)) return 1;
    CK(launch_pdl(match_filter2_kernel, dim3(std::min(mp.p_tiles * mp.slices, units) * 2), dim3(kMatch2Threads), Match2Smem::kTotal, st,
                  pdl, 2, tmP, ctx->tmG2, mp));
  } else {
    if (set_smem_attr(ctx, reinterpret_cast<const void*>(match_filter_kernel), MatchSmem::kTotal)) return 1;
    CK(launch_pdl(match_filter_kernel, dim3(std::min(mp.p_tiles * mp.slices, ctx->num_sms)), dim3(kMatchThreads), MatchSmem::kTotal, st,
                  pdl, 1, tmP, ctx->tmG, mp));
  }
  ctx->launches++;
  if (ctx->match_profiling) CK(cudaEventRecord(ctx->match_prof_ev[2], st));
  FinalizeParams fp;
  fp.cand_score = ctx->d_cand_score; fp.cand_idx = ctx->d_cand_idx; fp.slices = mp.slices;
  fp.probes = d_probe_f32; fp.gallery = ctx->d_gal; fp.N = N; fp.first_global_id = ctx->gal_first;
  fp.k = k; fp.thr = thr; fp.max_norm = ctx->d_gal_maxnorm;
  fp.rescore = std::min(kRescore, std::max(24, 8 * k));
  fp.out_score = s64; fp.out_idx = d_idx; fp.out_score_f32 = d_scores; fp.out_accept = d_accept;
  fp.flagged = ctx->d_flagged; fp.flag_rows = ctx->d_flag_rows; fp.flag_count = ctx->d_match_ctr;
  fp.push = pp;
  CK(launch_pdl(match_finalize_kernel, dim3(P), dim3(128), 0, st, pdl, 1, fp));
  ctx->launches++;
  // Rows whose proof failed get the exact scan.  Rare (none on the synthetic workloads), so the two kernels are
  // launched unconditionally and take the row list and its length from the device: no D2H, no synchronisation.
  ExactFixParams xp;
  xp.gallery = ctx->d_gal; xp.N = N; xp.probes = d_probe_f32; xp.rows = ctx->d_flag_rows; xp.count = ctx->d_match_ctr;
  xp.k = k; xp.thr = thr; xp.first_global_id = ctx->gal_first; xp.part = ctx->d_exact_part;
  xp.out_score = s64; xp.out_idx = d_idx; xp.out_score_f32 = d_scores; xp.out_accept = d_accept;
  xp.push = pp;
  CK(launch_pdl(match_exact_part_kernel, dim3(kExactBlocks, std::min(P, kExactRowsY)), dim3(256), 0, st, pdl, 1, xp));
  ctx->launches++;
  if (set_smem_attr(ctx, reinterpret_cast<const void*>(match_exact_fix_kernel), kExactFixSmemBytes)) return 1;
  CK(launch_pdl(match_exact_fix_kernel, dim3(std::min(P, 64)), dim3(128), kExactFixSmemBytes, st, pdl, 1, xp,
                static_cast<int>(kExactBlocks), std::min(P, kExactRowsY)));
  ctx->launches++;
  if (ctx->match_profiling) CK(cudaEventRecord(ctx->match_prof_ev[3], st));
  // frb_match_last_flagged: the count travels to pinned memory behind an event, read only when somebody asks
  CK(cudaMemcpyAsync(ctx->h_flag_count, ctx->d_match_ctr, 4, cudaMemcpyDeviceToHost, st));
  if (stream_capturing(st)) {   // inside a graph: the count still lands in pinned memory on every replay
    ctx->flag_event_pending = false;
    ctx->last_flagged = -1;
    return 0;
  }
  CK(cudaEventRecord(ctx->flag_event, st));
  ctx->flag_event_pending = true;
  return 0;
}

int scores64_scratch(frb_ctx* ctx, size_t need, double** out) {
  if (ensure(ctx, &ctx->d_scores64_tmp, &ctx->scores64_tmp_elems, need)) return 1;
  *out = ctx->d_scores64_tmp;
  return 0;
}

int match_locked(frb_ctx* ctx, const float* d_probes, int P, int k, float thr, int normalize, float* d_scores,
                 long long* d_idx, unsigned char* d_accept, double* d_scores64, cudaStream_t st) {
  if (k <= 0) return fail(ctx, "frb_match: k must be positive");
  if (match_workspace(ctx, P)) return 1;
  double* s64 = d_scores64;
  if (!s64 && scores64_scratch(ctx, static_cast<size_t>(P) * k, &s64)) return 1;
  if (ctx->match_profiling) CK(cudaEventRecord(ctx->match_prof_ev[0], st));
  // the kernel also zeroes this match's device state: its rows' admission floors and the flagged / pushed counters
  probe_prepare_kernel<<<P, 128, 0, st>>>(d_probes, normalize, ctx->d_probe_f32, ctx->d_probe_bf16, ctx->d_row_floor,
                                          ctx->d_match_ctr);
  CK(cudaGetLastError());
  ctx->launches++;
  return match_core(ctx, ctx->d_probe_f32, ctx->d_probe_bf16, P, k, thr, d_scores, d_idx, d_accept, s64, nullptr, st);
}

}  // namespace

// Measurement aid (bench.py per-kernel roofline): frb_match with CUDA events between its three parts; h_ms3 = prepare,
// filter kernel, finalize + exact fix-up (ms).  Needs the tensor-core path (N >= 4096, k <= 32); synchronises.
extern "C" int frb_match_profile(frb_ctx* ctx, const float* d_probes, int P, int k, float thr, int normalize,
                                 float* d_scores, long long* d_idx, unsigned char* d_accept, void* stream, float* h_ms3) {
  if (!ctx) return 1;
  if (P <= 0 || !h_ms3) return fail(ctx, "frb_match_profile: bad arguments");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  if (ctx->gal_N < kExactOnlyBelow || k > kExactMaxK) return fail(ctx, "frb_match_profile: needs the tensor-core path");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (auto& e : ctx->match_prof_ev)
    if (!e) CK(cudaEventCreate(&e));
  if (ws_begin(ctx, st)) return 1;
  ctx->match_profiling = true;
  const int rc = match_locked(ctx, d_probes, P, k, thr, normalize, d_scores, d_idx, d_accept, nullptr, st);
  ctx->match_profiling = false;
  if (rc) return rc;
  CK(cudaStreamSynchronize(st));
  ctx->ws_valid = false;
  for (int i = 0; i < 3; ++i) CK(cudaEventElapsedTime(&h_ms3[i], ctx->match_prof_ev[i], ctx->match_prof_ev[i + 1]));
  return 0;
}

// d_probes: [P][512] f32.  Any k >= 1: k <= 32 takes the tensor-core filter + exact re-score + proof (galleries of
// >= 4096 rows), larger k or smaller galleries the dense exact scan (search() accepts any top_k, gallery_manager.py:197).
extern "C" int frb_match(frb_ctx* ctx, const float* d_probes, int P, int k, float thr, int normalize, float* d_scores,
                         long long* d_idx, unsigned char* d_accept, double* d_scores64, void* stream) {
  if (!ctx) return 1;
  if (P <= 0) return 0;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ws_begin(ctx, st)) return 1;
  if (match_locked(ctx, d_probes, P, k, thr, normalize, d_scores, d_idx, d_accept, d_scores64, st)) return 1;
  return ws_end(ctx, st);
}

// ---- per-identity matching over a SAMPLE gallery (evaluate_models_v2.ipynb cells 3-5)
extern "C" int frb_gallery_upload_samples(frb_ctx* ctx, const float* samples, long long T, const long long* h_seg,
                                          long long S, int is_device) {
  if (!ctx) return 1;
  if (T < 0 || S < 0 || !h_seg || h_seg[0] != 0 || h_seg[S] != T) return fail(ctx, "frb_gallery_upload_samples: bad segments");
  std::vector<int> sid(static_cast<size_t>(T));
  for (long long i = 0; i < S; ++i) {
    const long long n = h_seg[i + 1] - h_seg[i];
    if (n < 0 || n > kAggMaxRows) return fail(ctx, "frb_gallery_upload_samples: identity %lld has %lld samples (0..%d supported)", i, n, kAggMaxRows);
    for (long long r = h_seg[i]; r < h_seg[i + 1]; ++r) sid[static_cast<size_t>(r)] = static_cast<int>(i);
  }
  if (int rc = frb_gallery_upload(ctx, samples, T, 0, is_device)) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  if (ensure(ctx, &ctx->d_seg, &ctx->seg_cap, static_cast<size_t>(S) + 1)) return 1;
  if (ensure(ctx, &ctx->d_sample_identity, &ctx->sid_cap, std::max<size_t>(static_cast<size_t>(T), 1))) return 1;
  CK(cudaMemcpy(ctx->d_seg, h_seg, (static_cast<size_t>(S) + 1) * 8, cudaMemcpyHostToDevice));
  if (T > 0) CK(cudaMemcpy(ctx->d_sample_identity, sid.data(), static_cast<size_t>(T) * 4, cudaMemcpyHostToDevice));
  ctx->gal_S = S;
  ctx->gallery_gen++;
  return 0;
}

namespace {
// exact identity scores of probes f0..f0+fc-1 (or the listed rows) into ctx->d_id_scores [fc][S] f64
int identity_exact_chunk(frb_ctx* ctx, const float* probes, const int* rows, int fc, int agg, int agg_k, float* d_out32,
                         cudaStream_t st) {
  const long long T = ctx->gal_N, S = ctx->gal_S;
  if (ensure(ctx, &ctx->d_exact, &ctx->exact_elems, std::max<size_t>(static_cast<size_t>(fc) * T, 1))) return 1;
  if (ensure(ctx, &ctx->d_id_scores, &ctx->id_scores_elems, std::max<size_t>(static_cast<size_t>(fc) * S, 1))) return 1;
  if (T > 0) {
    dim3 grid(static_cast<unsigned>(std::min<long long>((T + 7) / 8, 148 * 8)), fc);
    match_exact_scores_kernel<<<grid, 256, 0, st>>>(ctx->d_gal, T, probes, rows, ctx->d_exact);
    CK(cudaGetLastError());
    ctx->launches++;
  }
  dim3 g2(static_cast<unsigned>((S + 127) / 128), fc);
  identity_reduce_kernel<<<g2, 128, 0, st>>>(ctx->d_exact, T, ctx->d_seg, S, agg, agg_k, ctx->d_id_scores, d_out32);
  CK(cudaGetLastError());
  ctx->launches++;
  return 0;
}
}  // namespace

// Full [P][S] identity score matrix (what identify_probe's identity_scores dict holds), exact f64 arithmetic, f32 out.
extern "C" int frb_identity_scores(frb_ctx* ctx, const float* d_probes, int P, int normalize, int agg, int agg_k,
                                   float* d_out, long long S_expected, void* stream) {
  if (!ctx) return 1;
  if (P <= 0) return 0;
  if (agg < 0 || agg > 2 || agg_k < 1) return fail(ctx, "frb_identity_scores: bad aggregation");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ctx->gal_S <= 0) return fail(ctx, "frb_identity_scores: no sample gallery uploaded");
  if (S_expected >= 0 && S_expected != ctx->gal_S)
    return fail(ctx, "frb_identity_scores: the resident sample gallery has %lld identities, the caller expects %lld (another "
                     "gallery was uploaded to this ctx in between)", ctx->gal_S, S_expected);
  if (ws_begin(ctx, st)) return 1;
  float* d_norm = nullptr;
  CK(cudaMallocAsync(reinterpret_cast<void**>(&d_norm), static_cast<size_t>(P) * 512 * 4, st));
  probe_prepare_kernel<<<P, 128, 0, st>>>(d_probes, normalize, d_norm, nullptr, nullptr, nullptr);
  CK(cudaGetLastError());
  ctx->launches++;
  for (int f0 = 0; f0 < P; f0 += kExactChunk) {
    const int fc = std::min(kExactChunk, P - f0);
    if (identity_exact_chunk(ctx, d_norm + static_cast<size_t>(f0) * 512, nullptr, fc, agg, agg_k,
                             d_out + static_cast<size_t>(f0) * ctx->gal_S, st))
      return 1;
  }
  CK(cudaFreeAsync(d_norm, st));
  return ws_end(ctx, st);
}

// Top-k identities per probe.  Outputs as frb_match, with identity indices instead of gallery rows.
extern "C" int frb_match_identities(frb_ctx* ctx, const float* d_probes, int P, int k, float thr, int normalize, int agg,
                                    int agg_k, float* d_scores, long long* d_idx, unsigned char* d_accept, void* stream) {
  if (!ctx) return 1;
  if (P <= 0) return 0;
  if (agg < 0 || agg > 2 || agg_k < 1) return fail(ctx, "frb_match_identities: bad aggregation");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long T = ctx->gal_N, S = ctx->gal_S;
  if (S <= 0) return fail(ctx, "frb_match_identities: no sample gallery uploaded");
  if (ws_begin(ctx, st)) return 1;
  if (k <= 0 || k > kRescore / 2) return fail(ctx, "frb_match_identities: k must be in [1, %d]", kRescore / 2);
  constexpr int KS = kRescore / 2;  // exact top-32 samples per probe feed the candidate set
  {
    const size_t need = static_cast<size_t>(P) * KS;
    if (ensure(ctx, &ctx->d_id_top_idx, &ctx->id_top_cap, need)) return 1;
    if (ensure(ctx, &ctx->d_id_top_sc, &ctx->id_top_sc_cap, need)) return 1;
    if (ensure(ctx, &ctx->d_id_acc, &ctx->id_acc_cap, need)) return 1;
    if (ensure(ctx, &ctx->d_id_sc32, &ctx->id_sc32_cap, need)) return 1;
  }
  if (ensure(ctx, &ctx->d_scores64_tmp, &ctx->scores64_tmp_elems, static_cast<size_t>(P) * std::max(k, KS))) return 1;
  // 1. exact top-KS SAMPLES (tensor-core filter + exact re-score + proof, or the exact scan for small galleries)
  if (match_locked(ctx, d_probes, P, KS, -INFINITY, normalize, ctx->d_id_sc32, ctx->d_id_top_idx, ctx->d_id_acc, ctx->d_id_top_sc, st))
    return 1;
  // 2. candidates -> exact aggregates -> proof
  IdentityCandParams cp;
  cp.probes = ctx->d_probe_f32; cp.gallery = ctx->d_gal; cp.seg = ctx->d_seg; cp.sample_identity = ctx->d_sample_identity;
  cp.top_idx = ctx->d_id_top_idx; cp.top_sc = ctx->d_id_top_sc; cp.KS = KS; cp.T = T; cp.S = S;
  cp.agg = agg; cp.agg_k = agg_k; cp.k = k; cp.thr = thr;
  cp.out_score = ctx->d_scores64_tmp; cp.out_idx = d_idx; cp.out_score_f32 = d_scores; cp.out_accept = d_accept;
  cp.flagged = ctx->d_flagged;
  identity_candidates_kernel<<<P, 128, 0, st>>>(cp);
  CK(cudaGetLastError());
  ctx->launches++;
  // 3. exact scan over all identities for the rows whose proof failed
  ctx->h_flagged.resize(P);
  CK(cudaMemcpyAsync(ctx->h_flagged.data(), ctx->d_flagged, static_cast<size_t>(P) * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  std::vector<int> rows;
  for (int i = 0; i < P; ++i)
    if (ctx->h_flagged[i]) rows.push_back(i);
  ctx->last_flagged = static_cast<int>(rows.size());
  ctx->flag_event_pending = false;   // the count reported is the identity proof's, not the sample filter's
  if (!rows.empty()) {
    CK(cudaMemcpyAsync(ctx->d_flag_rows, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice, st));
    const int F = static_cast<int>(rows.size());
    for (int f0 = 0; f0 < F; f0 += kExactChunk) {
      const int fc = std::min(kExactChunk, F - f0);
      if (identity_exact_chunk(ctx, ctx->d_probe_f32, ctx->d_flag_rows + f0, fc, agg, agg_k, nullptr, st)) return 1;
      match_exact_topk_kernel<<<fc, 256, 0, st>>>(ctx->d_id_scores, S, ctx->d_flag_rows + f0, k, thr, 0, ctx->d_scores64_tmp,
                                                  d_idx, d_scores, d_accept);
      CK(cudaGetLastError());
      ctx->launches++;
    }
  }
  return ws_end(ctx, st);
}

namespace {
int merge_launch(frb_ctx* ctx, const TopkRec* d_rec, int G, int P, int k, float thr, float* d_scores, long long* d_idx,
                 unsigned char* d_accept, double* d_scores64, const XchgWait& wait, cudaStream_t st) {
  topk_merge_kernel<<<(P + 127) / 128, 128, 0, st>>>(d_rec, G, P, k, thr, d_scores64, d_idx, d_scores, d_accept, wait);
  CK(cudaGetLastError());
  ctx->launches++;
  return 0;
}
}  // namespace

// merge G gathered per-rank top-k lists given as separate score / id arrays [G][P][k]
extern "C" int frb_topk_merge(frb_ctx* ctx, const double* d_in_scores64, const long long* d_in_idx, int G, int P, int k,
                              float thr, float* d_scores, long long* d_idx, unsigned char* d_accept,
                              double* d_scores64, void* stream) {
  if (!ctx) return 1;
  if (P <= 0) return 0;
  if (G <= 0 || k <= 0) return fail(ctx, "frb_topk_merge: bad G / k");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ws_begin(ctx, st)) return 1;
  const size_t n = static_cast<size_t>(G) * P * k;
  if (ensure(ctx, &ctx->d_merge_rec, &ctx->merge_rec_cap, n)) return 1;
  topk_pack_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(d_in_scores64, d_in_idx, n, ctx->d_merge_rec);
  CK(cudaGetLastError());
  ctx->launches++;
  XchgWait nowait;
  memset(&nowait, 0, sizeof(nowait));
  if (merge_launch(ctx, ctx->d_merge_rec, G, P, k, thr, d_scores, d_idx, d_accept, d_scores64, nowait, st)) return 1;
  return ws_end(ctx, st);
}

// same with the lists as 16-byte records (f64 score, i64 id) [G][P][k]: what ONE all-gather of the per-rank results
// delivers (dist.ShardedGallery's NCCL exchange)
extern "C" int frb_topk_merge_packed(frb_ctx* ctx, const void* d_records, int G, int P, int k, float thr, float* d_scores,
                                     long long* d_idx, unsigned char* d_accept, void* stream) {
  if (!ctx) return 1;
  if (P <= 0) return 0;
  if (G <= 0 || k <= 0) return fail(ctx, "frb_topk_merge_packed: bad G / k");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  XchgWait nowait;
  memset(&nowait, 0, sizeof(nowait));
  return merge_launch(ctx, reinterpret_cast<const TopkRec*>(d_records), G, P, k, thr, d_scores, d_idx, d_accept, nullptr,
                      nowait, static_cast<cudaStream_t>(stream));
}

// frb_match with the result also written as records [P][k] (f64 score, i64 global id): the payload of the exchange
extern "C" int frb_match_packed(frb_ctx* ctx, const float* d_probes, int P, int k, float thr, int normalize,
                                void* d_records, void* stream) {
  if (!ctx) return 1;
  if (P <= 0) return 0;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ws_begin(ctx, st)) return 1;
  if (k <= 0) return fail(ctx, "frb_match_packed: k must be positive");
  if (stage_match(ctx, P, k)) return 1;
  double* s64 = nullptr;
  if (scores64_scratch(ctx, static_cast<size_t>(P) * k, &s64)) return 1;
  if (match_locked(ctx, d_probes, P, k, thr, normalize, ctx->d_stage_sc, ctx->d_stage_idx, ctx->d_stage_acc, s64, st)) return 1;
  const size_t n = static_cast<size_t>(P) * k;
  topk_pack_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(s64, ctx->d_stage_idx, n,
                                                                            reinterpret_cast<TopkRec*>(d_records));
  CK(cudaGetLastError());
  ctx->launches++;
  return ws_end(ctx, st);
}

// ====================================================================== identity-sharded match over peer memory
// One exchange buffer per rank (cudaMalloc, exported with cudaIpc): the probes of ALL ranks (fp32 + bf16), two
// parities of result slots [world][max_probes][max_k] records, and the flag words the peers raise.
extern "C" int frb_xchg_create(frb_ctx* ctx, int world, int rank, int max_probes, int max_k, void* h_handle_out64) {
  if (!ctx) return 1;
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return fail(ctx, "frb_xchg_create: world must be 1..%d", kMaxPeers);
  if (max_probes < 1 || max_k < 1 || max_k > kExactMaxK) return fail(ctx, "frb_xchg_create: bad max_probes / max_k");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  auto& x = ctx->xchg;
  if (x.local) return fail(ctx, "frb_xchg_create: exchange already created on this ctx");
  auto up = [](size_t v) { return (v + 1023) / 1024 * 1024; };
  x.world = world; x.rank = rank; x.max_probes = max_probes; x.max_k = max_k;
  x.off_pflag = 0;                       // [world] u32, 128 B apart
  x.off_rflag = up(static_cast<size_t>(world) * 4 * kFlagStride);
  x.off_f32 = x.off_rflag + up(static_cast<size_t>(world) * 4 * kFlagStride);
  x.off_bf16 = x.off_f32 + up(static_cast<size_t>(max_probes) * 512 * 4);
  x.off_slots = x.off_bf16 + up(static_cast<size_t>(max_probes) * 512 * 2);
  x.bytes = x.off_slots + 2 * up(static_cast<size_t>(world) * max_probes * max_k * sizeof(TopkRec));
  CK(cudaMalloc(reinterpret_cast<void**>(&x.local), x.bytes));
  CK(cudaMemset(x.local, 0, x.bytes));
  CK(cudaDeviceSynchronize());
  CK(cudaHostAlloc(reinterpret_cast<void**>(&x.h_status), 64, cudaHostAllocMapped));
  x.h_status[0] = 0;
  if (const char* e = getenv("FRB_XCHG_TIMEOUT_MS")) x.timeout_ns = static_cast<unsigned long long>(atoll(e)) * 1000000ull;
  x.peer[rank] = x.local;
  if (world > 1) {
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, x.local));
    if (h_handle_out64) memcpy(h_handle_out64, &h, 64);
  } else if (h_handle_out64) {
    memset(h_handle_out64, 0, 64);
  }
  x.connected = (world == 1);
  // every workspace a sharded match of up to max_probes x max_k needs, now: frb_match_sharded itself then never
  // allocates or frees (both may synchronise the device - with a peer's wait kernel spinning on it)
  if (match_workspace(ctx, max_probes)) return 1;
  if (stage_match(ctx, max_probes, max_k)) return 1;
  double* s64 = nullptr;
  if (scores64_scratch(ctx, static_cast<size_t>(max_probes) * max_k, &s64)) return 1;
  if (ensure(ctx, &ctx->d_exact_part, &ctx->exact_part_cap, static_cast<size_t>(std::max(max_probes, kExactRowsY * kExactRowsY)) * kExactBlocks * max_k)) return 1;
  if (ensure(ctx, &ctx->d_exact, &ctx->exact_elems, static_cast<size_t>(kExactChunk) * kExactOnlyBelow)) return 1;
  {
    const int slices = kMaxCandPad / kCand;   // the most the filter ever uses
    if (ctx->match_cap_slices < slices || !ctx->d_cand_score) {
      if (ctx->d_cand_score) CK(cudaFree(ctx->d_cand_score));
      if (ctx->d_cand_idx) CK(cudaFree(ctx->d_cand_idx));
      ctx->d_cand_score = nullptr; ctx->d_cand_idx = nullptr;
      const size_t n = static_cast<size_t>(ctx->match_cap_P) * slices * kCand;
      CK(cudaMalloc(&ctx->d_cand_score, n * 4));
      CK(cudaMalloc(&ctx->d_cand_idx, n * 4));
      ctx->match_cap_slices = slices;
    }
  }
  // ... and every kernel of the sharded path is loaded now: with lazy module loading the FIRST launch of a kernel may
  // synchronise the context - behind this rank's own wait kernel, which spins until a peer arrives
  {
    cudaFuncAttributes fa;
    const void* kernels[] = {reinterpret_cast<const void*>(probe_push_kernel), reinterpret_cast<const void*>(probe_push_empty_kernel),
                             reinterpret_cast<const void*>(xchg_wait_kernel), reinterpret_cast<const void*>(match_filter_kernel),
                             reinterpret_cast<const void*>(match_filter2_kernel), reinterpret_cast<const void*>(match_finalize_kernel),
                             reinterpret_cast<const void*>(match_exact_part_kernel), reinterpret_cast<const void*>(match_exact_fix_kernel),
                             reinterpret_cast<const void*>(match_exact_scores_kernel), reinterpret_cast<const void*>(match_exact_topk_kernel),
                             reinterpret_cast<const void*>(match_fill_empty_kernel), reinterpret_cast<const void*>(xchg_push_rows_kernel),
                             reinterpret_cast<const void*>(topk_merge_kernel)};
    for (const void* k : kernels) CK(cudaFuncGetAttributes(&fa, k));
  }
  // opt the kernels into their shared-memory sizes now as well
  if (set_smem_attr(ctx, reinterpret_cast<const void*>(match_filter2_kernel), Match2Smem::kTotal)) return 1;
  if (set_smem_attr(ctx, reinterpret_cast<const void*>(match_filter_kernel), MatchSmem::kTotal)) return 1;
  if (set_smem_attr(ctx, reinterpret_cast<const void*>(match_exact_fix_kernel), kExactFixSmemBytes)) return 1;
  return 0;
}

// h_handles_all: world x 64 bytes, rank r's handle at offset 64 r (all-gathered by the caller)
extern "C" int frb_xchg_connect(frb_ctx* ctx, const void* h_handles_all) {
  if (!ctx) return 1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  auto& x = ctx->xchg;
  if (!x.local) return fail(ctx, "frb_xchg_connect: call frb_xchg_create first");
  for (int g = 0; g < x.world; ++g) {
    if (g == x.rank || x.peer[g]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const uint8_t*>(h_handles_all) + 64 * g, 64);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(ctx, "frb_xchg_connect: cudaIpcOpenMemHandle(rank %d) -> %s (no peer access between these GPUs?)", g,
                  cudaGetErrorString(e));
    }
    x.peer[g] = static_cast<uint8_t*>(ptr);
    x.peer_ipc[g] = true;
  }
  x.connected = true;
  return 0;
}

// test hook: several ranks living in ONE process (each its own ctx, own stream) cannot open each other's cudaIpc
// handles; they exchange the raw device pointers instead
extern "C" void* frb_xchg_local_buffer(frb_ctx* ctx) { return ctx ? ctx->xchg.local : nullptr; }
extern "C" int frb_xchg_connect_local(frb_ctx* ctx, int peer_rank, void* d_peer_buffer) {
  if (!ctx) return 1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  auto& x = ctx->xchg;
  if (!x.local || peer_rank < 0 || peer_rank >= x.world || !d_peer_buffer) return fail(ctx, "frb_xchg_connect_local: bad arguments");
  x.peer[peer_rank] = static_cast<uint8_t*>(d_peer_buffer);
  bool all = true;
  for (int g = 0; g < x.world; ++g) all = all && x.peer[g] != nullptr;
  x.connected = all;
  return 0;
}

extern "C" int frb_xchg_status(frb_ctx* ctx) { return (ctx && ctx->xchg.h_status) ? ctx->xchg.h_status[0] : 0; }

// Identity-sharded match, collective over the ranks of the exchange (every rank calls it with the same P_total / k /
// thr, its own rows [p_lo, p_lo + p_cnt) of the probe set, its shard resident with first_global_id = shard start):
//   1. probe_push_kernel: normalise the local probes and store them (fp32 + bf16) into EVERY rank's probe buffer
//   2. xchg_wait_kernel: until all ranks' probes have landed here
//   3. filter + finalize (+ exact fix-up) of ALL probes against the local shard; every finished row's k records are
//      stored straight into every rank's result slot, the last row raises this rank's result flags
//   4. topk_merge_kernel: waits for all ranks' result flags, canonical merge -> outputs for ALL P_total probes
// No host synchronisation, no NCCL on the data path.
extern "C" int frb_match_sharded(frb_ctx* ctx, const float* d_local_probes, int p_lo, int p_cnt, int P_total, int k,
                                 float thr, int normalize, float* d_scores, long long* d_idx, unsigned char* d_accept,
                                 void* stream) {
  if (!ctx) return 1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  auto& x = ctx->xchg;
  if (!x.local || !x.connected) return fail(ctx, "frb_match_sharded: exchange not created / connected");
  if (P_total <= 0) return 0;
  if (P_total > x.max_probes || k < 1 || k > x.max_k) return fail(ctx, "frb_match_sharded: P_total %d / k %d exceed the exchange (%d / %d)", P_total, k, x.max_probes, x.max_k);
  if (p_lo < 0 || p_cnt < 0 || p_lo + p_cnt > P_total) return fail(ctx, "frb_match_sharded: bad local probe range");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ws_begin(ctx, st)) return 1;
  if (match_workspace(ctx, P_total)) return 1;
  if (stage_match(ctx, P_total, k)) return 1;
  double* s64 = nullptr;
  if (scores64_scratch(ctx, static_cast<size_t>(P_total) * k, &s64)) return 1;
  const unsigned epoch = ++x.epoch;
  const int G = x.world;
  CK(cudaMemsetAsync(ctx->d_match_ctr, 0, 12, st));   // flagged rows, rows pushed, probe rows pushed
  CK(cudaMemsetAsync(ctx->d_row_floor, 0, static_cast<size_t>(P_total) * 4, st));
  // 1. probes -> everyone
  ProbePush pq;
  memset(&pq, 0, sizeof(pq));
  pq.world = G; pq.row0 = p_lo; pq.rows = p_cnt; pq.epoch = epoch; pq.done_rows = ctx->d_match_ctr + 2;
  for (int g = 0; g < G; ++g) {
    pq.f32[g] = reinterpret_cast<float*>(x.peer[g] + x.off_f32);
    pq.bf16[g] = reinterpret_cast<__nv_bfloat16*>(x.peer[g] + x.off_bf16);
    pq.flag[g] = reinterpret_cast<unsigned*>(x.peer[g] + x.off_pflag + 4 * kFlagStride * x.rank);
  }
  if (p_cnt > 0) probe_push_kernel<<<p_cnt, 128, 0, st>>>(d_local_probes, normalize, pq);
  else probe_push_empty_kernel<<<1, 32, 0, st>>>(pq);
  CK(cudaGetLastError());
  ctx->launches++;
  // 2. wait for everyone's probes
  XchgWait wq;
  wq.world = G; wq.epoch = epoch; wq.flag = nullptr; wq.timeout_ns = x.timeout_ns; wq.status = x.h_status;
  wq.flag = reinterpret_cast<const unsigned*>(x.local + x.off_pflag);
  xchg_wait_kernel<<<1, 32, 0, st>>>(wq);
  CK(cudaGetLastError());
  ctx->launches++;
  // 3. local match of all probes, rows pushed to the peers as they finish
  const size_t slot_bytes = (static_cast<size_t>(G) * x.max_probes * x.max_k * sizeof(TopkRec) + 1023) / 1024 * 1024;
  const size_t parity_off = x.off_slots + (epoch & 1u) * slot_bytes;
  PeerPush pp;
  memset(&pp, 0, sizeof(pp));
  pp.world = G; pp.rank = x.rank; pp.epoch = epoch; pp.P = P_total; pp.done_rows = ctx->d_match_ctr + 1;
  for (int g = 0; g < G; ++g) {
    pp.slot[g] = reinterpret_cast<TopkRec*>(x.peer[g] + parity_off) + static_cast<size_t>(x.rank) * P_total * k;
    pp.flag[g] = reinterpret_cast<unsigned*>(x.peer[g] + x.off_rflag + 4 * kFlagStride * x.rank);
  }
  if (match_core(ctx, reinterpret_cast<const float*>(x.local + x.off_f32),
                 reinterpret_cast<const __nv_bfloat16*>(x.local + x.off_bf16), P_total, k, thr, ctx->d_stage_sc,
                 ctx->d_stage_idx, ctx->d_stage_acc, s64, &pp, st))
    return 1;
  // 4. wait for everyone's rows, merge
  XchgWait wr = wq;
  wr.flag = reinterpret_cast<const unsigned*>(x.local + x.off_rflag);
  if (merge_launch(ctx, reinterpret_cast<const TopkRec*>(x.local + parity_off), G, P_total, k, thr, d_scores, d_idx,
                   d_accept, nullptr, wr, st))
    return 1;
  return ws_end(ctx, st);
}

// ====================================================================== host-buffer entry points
namespace {

int stage_embed(frb_ctx* ctx, int B, int flags) {
  const int Bn = (flags & FRB_EMBED_FLIP) ? 2 * B : B;
  if (ensure(ctx, &ctx->d_stage_in, &ctx->stage_in_elems, static_cast<size_t>(Bn) * 112 * 112 * 3)) return 1;
  if (ctx->stage_emb_rows < static_cast<size_t>(B)) {
    if (ctx->d_stage_emb) CK(cudaFree(ctx->d_stage_emb));
    if (ctx->d_stage_norm) CK(cudaFree(ctx->d_stage_norm));
    ctx->d_stage_emb = nullptr; ctx->d_stage_norm = nullptr;
    CK(cudaMalloc(&ctx->d_stage_emb, static_cast<size_t>(B) * 512 * 4));
    CK(cudaMalloc(&ctx->d_stage_norm, static_cast<size_t>(B) * 4));
    ctx->stage_emb_rows = B;
  }
  return 0;
}

int stage_match(frb_ctx* ctx, int P, int k) {
  if (ctx->stage_match_rows < static_cast<size_t>(P) || ctx->stage_match_k < k) {
    if (ctx->d_stage_sc) CK(cudaFree(ctx->d_stage_sc));
    if (ctx->d_stage_idx) CK(cudaFree(ctx->d_stage_idx));
    if (ctx->d_stage_acc) CK(cudaFree(ctx->d_stage_acc));
    ctx->d_stage_sc = nullptr; ctx->d_stage_idx = nullptr; ctx->d_stage_acc = nullptr;
    const size_t rows = std::max<size_t>(P, ctx->stage_match_rows);
    const int kk = std::max(k, ctx->stage_match_k);
    CK(cudaMalloc(&ctx->d_stage_sc, rows * kk * 4));
    CK(cudaMalloc(&ctx->d_stage_idx, rows * kk * 8));
    CK(cudaMalloc(&ctx->d_stage_acc, rows));
    ctx->stage_match_rows = rows;
    ctx->stage_match_k = kk;
  }
  return 0;
}

int embed_host_locked(frb_ctx* ctx, const uint8_t* h_rgb, int B, int S, int flags, cudaStream_t st) {
  if (S != 112 && S != 224) return fail(ctx, "S must be 112 or 224 (got %d)", S);
  const size_t in_bytes = static_cast<size_t>(B) * S * S * 3;
  if (stage_embed(ctx, B, flags)) return 1;
  const uint8_t* d_u8 = nullptr;
  for (auto& sl : ctx->prefetch)
    if (sl.h_ptr != nullptr && sl.h_ptr == h_rgb && sl.bytes == in_bytes) {
      // these crops are already on their way (frb_prefetch_host): read them where they land, ordered after the copy
      CK(cudaStreamWaitEvent(st, sl.done, 0));
      d_u8 = sl.d_buf;
      sl.h_ptr = nullptr;   // consumed (this call synchronises before it returns, so the slot is free afterwards)
      break;
    }
  if (d_u8 == nullptr) {
    if (ensure(ctx, &ctx->d_stage_u8, &ctx->stage_u8_bytes, in_bytes)) return 1;
    CK(cudaMemcpyAsync(ctx->d_stage_u8, h_rgb, in_bytes, cudaMemcpyHostToDevice, st));
    d_u8 = ctx->d_stage_u8;
  }
  const size_t total = static_cast<size_t>(B) * 112 * 112;
  preprocess_u8_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
      d_u8, B, S, ctx->d_lut, ctx->d_stage_in, (flags & FRB_EMBED_FLIP) ? 1 : 0);
  CK(cudaGetLastError());
  ctx->launches++;
  return embed_locked(ctx, ctx->d_stage_in, B, flags, ctx->d_stage_emb, ctx->d_stage_norm, nullptr, st);
}

}  // namespace

// Start the host->device copy of the crops a LATER frb_embed_host / frb_embed_match_host call will be given, on a
// separate stream, so it overlaps the call that runs in between.
extern "C" int frb_prefetch_host(frb_ctx* ctx, const uint8_t* h_rgb, int B, int S) {
  if (!ctx) return 1;
  if (B <= 0 || !h_rgb) return 0;
  if (S != 112 && S != 224) return fail(ctx, "S must be 112 or 224 (got %d)", S);
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  if (!ctx->copy_stream) CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  const size_t in_bytes = static_cast<size_t>(B) * S * S * 3;
  // two slots: one may hold the batch the NEXT call consumes while this prefetch is for the call after it.
  // A free slot first, else the older unconsumed one is overwritten (its copy is ordered on the copy stream).
  frb_ctx::PrefetchSlot* sl = nullptr;
  for (auto& c : ctx->prefetch)
    if (c.h_ptr == h_rgb && c.bytes == in_bytes) return 0;   // already on its way
  for (auto& c : ctx->prefetch)
    if (c.h_ptr == nullptr) { sl = &c; break; }
  if (!sl) sl = ctx->prefetch[0].seq < ctx->prefetch[1].seq ? &ctx->prefetch[0] : &ctx->prefetch[1];
  if (!sl->done) CK(cudaEventCreateWithFlags(&sl->done, cudaEventDisableTiming));
  if (ensure(ctx, &sl->d_buf, &sl->cap, in_bytes)) return 1;
  CK(cudaMemcpyAsync(sl->d_buf, h_rgb, in_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
  CK(cudaEventRecord(sl->done, ctx->copy_stream));
  sl->h_ptr = h_rgb;
  sl->bytes = in_bytes;
  sl->seq = ++ctx->prefetch_seq;
  return 0;
}

extern "C" int frb_embed_host(frb_ctx* ctx, const uint8_t* h_rgb, int B, int S, int flags, float* h_emb, float* h_norm) {
  if (!ctx) return 1;
  if (B <= 0) return 0;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->own_stream;
  if (ws_begin(ctx, st)) return 1;
  if (embed_host_locked(ctx, h_rgb, B, S, flags, st)) return 1;
  CK(cudaMemcpyAsync(h_emb, ctx->d_stage_emb, static_cast<size_t>(B) * 512 * 4, cudaMemcpyDeviceToHost, st));
  if (h_norm) CK(cudaMemcpyAsync(h_norm, ctx->d_stage_norm, static_cast<size_t>(B) * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  ctx->ws_valid = false;   // nothing of this ctx is in flight any more
  return 0;
}

extern "C" int frb_match_host(frb_ctx* ctx, const float* h_probes, int P, int k, float thr, int normalize,
                              float* h_scores, long long* h_idx, unsigned char* h_accept) {
  if (!ctx) return 1;
  if (P <= 0) return 0;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->own_stream;
  if (ws_begin(ctx, st)) return 1;
  if (stage_match(ctx, P, k)) return 1;
  if (ctx->stage_emb_rows < static_cast<size_t>(P)) {
    if (ctx->d_stage_emb) CK(cudaFree(ctx->d_stage_emb));
    if (ctx->d_stage_norm) CK(cudaFree(ctx->d_stage_norm));
    ctx->d_stage_emb = nullptr; ctx->d_stage_norm = nullptr;
    CK(cudaMalloc(&ctx->d_stage_emb, static_cast<size_t>(P) * 512 * 4));
    CK(cudaMalloc(&ctx->d_stage_norm, static_cast<size_t>(P) * 4));
    ctx->stage_emb_rows = P;
  }
  CK(cudaMemcpyAsync(ctx->d_stage_emb, h_probes, static_cast<size_t>(P) * 512 * 4, cudaMemcpyHostToDevice, st));
  if (match_locked(ctx, ctx->d_stage_emb, P, k, thr, normalize, ctx->d_stage_sc, ctx->d_stage_idx, ctx->d_stage_acc,
                   nullptr, st))
    return 1;
  CK(cudaMemcpyAsync(h_scores, ctx->d_stage_sc, static_cast<size_t>(P) * k * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_idx, ctx->d_stage_idx, static_cast<size_t>(P) * k * 8, cudaMemcpyDeviceToHost, st));
  if (h_accept) CK(cudaMemcpyAsync(h_accept, ctx->d_stage_acc, P, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  ctx->ws_valid = false;   // nothing of this ctx is in flight any more
  return 0;
}

extern "C" int frb_embed_match_host(frb_ctx* ctx, const uint8_t* h_rgb, int B, int S, int flags, int k, float thr,
                                    float* h_emb, float* h_scores, long long* h_idx, unsigned char* h_accept) {
  if (!ctx) return 1;
  if (B <= 0) return 0;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->own_stream;
  if (ws_begin(ctx, st)) return 1;
  if (embed_host_locked(ctx, h_rgb, B, S, flags, st)) return 1;
  if (stage_match(ctx, B, k)) return 1;
  // search() re-normalises the query (gallery_manager.py:195)
  if (match_locked(ctx, ctx->d_stage_emb, B, k, thr, 1, ctx->d_stage_sc, ctx->d_stage_idx, ctx->d_stage_acc, nullptr, st))
    return 1;
  if (h_emb) CK(cudaMemcpyAsync(h_emb, ctx->d_stage_emb, static_cast<size_t>(B) * 512 * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_scores, ctx->d_stage_sc, static_cast<size_t>(B) * k * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h_idx, ctx->d_stage_idx, static_cast<size_t>(B) * k * 8, cudaMemcpyDeviceToHost, st));
  if (h_accept) CK(cudaMemcpyAsync(h_accept, ctx->d_stage_acc, B, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  ctx->ws_valid = false;   // nothing of this ctx is in flight any more
  return 0;
}

// ====================================================================== test hooks
extern "C" int frb_debug_gemm(frb_ctx* ctx, const void* d_A, const void* d_B, int M, int N, int K, int splits,
                              float* d_C, void* stream) {
  if (!ctx) return 1;
  if (K % 64 || N % 64) return fail(ctx, "frb_debug_gemm: K and N must be multiples of 64");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  const int bn = (N % 256 == 0) ? 256 : ((N % 128 == 0) ? 128 : 64);
  GemmParams gp;
  memset(&gp, 0, sizeof(gp));
  gp.M = M; gp.N = N; gp.num_kb_main = K / 64; gp.num_kb_sc = 0;
  gp.num_splits = choose_splits(gp.num_kb_main, std::max(1, splits));
  if (gp.num_splits != 1 && gp.num_splits != splits) return fail(ctx, "frb_debug_gemm: splits=%d not realisable (got %d)", splits, gp.num_splits);
  gp.out_f32 = d_C;
  CUtensorMap a, b;
  if (make_tmap_2d(ctx, &a, d_A, K, M, 128)) return 1;
  if (make_tmap_2d(ctx, &b, d_B, K, N, bn)) return 1;
  const int tiles = ((M + 127) / 128) * (N / bn) * gp.num_splits;
  return launch_gemm(ctx, bn, A_TILED, 1, a, a, b, gp, std::min(tiles, ctx->num_sms), static_cast<cudaStream_t>(stream));
}

extern "C" int frb_debug_conv(frb_ctx* ctx, const frb_layer_desc* L, int B, const void* d_in, const void* d_sc,
                              const void* d_res, const void* d_w, const float* d_bias, const float* d_prelu,
                              void* d_out, int use_ref, void* stream) {
  if (!ctx || !L) return 1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int P = out_dim(L->hin, L->ksize, L->stride, L->pad), Q = out_dim(L->win, L->ksize, L->stride, L->pad);
  if (use_ref) {
    ConvRefParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.in = reinterpret_cast<const __nv_bfloat16*>(d_in);
    rp.sc = reinterpret_cast<const __nv_bfloat16*>(d_sc);
    rp.wt = reinterpret_cast<const __nv_bfloat16*>(d_w);
    rp.bias = d_bias; rp.bias_cases = L->bias_cases;
    rp.prelu = L->has_prelu ? d_prelu : nullptr;
    rp.residual = reinterpret_cast<const __nv_bfloat16*>(d_res);
    rp.res_stride = L->res_stride; rp.RH = L->res_h; rp.RW = L->res_w;
    rp.out = reinterpret_cast<__nv_bfloat16*>(d_out);
    rp.B = B; rp.H = L->hin; rp.W = L->win; rp.Cin = L->cin; rp.Cout = L->cout; rp.P = P; rp.Q = Q;
    rp.stride = L->stride; rp.ksize = L->ksize; rp.pad = L->pad;
    rp.SH = L->sc_hin; rp.SW = L->sc_win; rp.Csc = d_sc ? L->sc_cin : 0; rp.sc_stride = L->sc_stride;
    const size_t total = static_cast<size_t>(B) * P * Q * L->cout;
    conv_ref_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(rp);
    CK(cudaGetLastError());
    ctx->launches++;
    return 0;
  }
  CUtensorMap a, a2, b;
  GemmParams gp;
  int bn, grid;
  frb_layer_desc LL = *L;
  if (!d_sc) { LL.sc_buf = -1; LL.sc_cin = 0; }
  if (!d_res) LL.res_buf = -1;
  if (slab_eligible(ctx, LL, d_sc != nullptr)) {
    SlabParams sp;
    int smem_bytes;
    CUtensorMap r;
    if (setup_slab(ctx, LL, B, d_in, d_res, d_w, d_bias, d_prelu, d_out, &a, &b, &sp, &smem_bytes, &grid, &a2, &r)) return 1;
    return launch_slab(ctx, LL.cin / 64, a, b, a2, r, sp, smem_bytes, grid, st);
  }
  if (setup_conv(ctx, LL, B, d_in, d_sc, d_res, d_w, d_bias, d_prelu, d_out, &a, &a2, &b, &gp, &bn, &grid)) return 1;
  if (gp.tail_split > 1) {
    size_t pf, fl;
    tail_split_needs(gp, bn, grid, &pf, &fl);
    if (ensure(ctx, &ctx->d_tail_partial, &ctx->tail_partial_cap, pf)) return 1;
    if (ensure(ctx, &ctx->d_tail_flags, &ctx->tail_flags_cap, fl)) return 1;
    ctx->plan.gp.clear();  // the resident plan's scratch pointers may have moved
    gp.tail_partial = ctx->d_tail_partial;
    gp.tail_flag = ctx->d_tail_flags;
    CK(cudaMemsetAsync(ctx->d_tail_flags, 0, fl * sizeof(int), st));
  }
  return launch_conv(ctx, bn, a, a2, b, gp, grid, st);
}

namespace {
__global__ void im2col_dump_kernel(const __grid_constant__ CUtensorMap tm, int c, int w, int h, int n, int off_w,
                                   int off_h, uint4* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 16384);
    tma_load_im2col_4d(&tm, bar, smem, c, w, h, n, static_cast<uint16_t>(off_w), static_cast<uint16_t>(off_h));
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = reinterpret_cast<uint4*>(smem)[i];
}
}  // namespace

extern "C" int frb_debug_im2col(frb_ctx* ctx, const void* d_in, int B, int H, int W, int C, int ksize, int stride,
                                int pad, int m0, int c0, int tap_r, int tap_s, void* d_out_16k, void* stream) {
  if (!ctx) return 1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  CUtensorMap tm;
  if (make_tmap_im2col(ctx, &tm, d_in, B, H, W, C, ksize, stride, pad)) return 1;
  const int P = out_dim(H, ksize, stride, pad), Q = out_dim(W, ksize, stride, pad);
  const int img = m0 / (P * Q), rem = m0 % (P * Q), pp = rem / Q, qq = rem % Q;
  if (set_smem_attr(ctx, reinterpret_cast<const void*>(im2col_dump_kernel), 16384 + 64 + 1024)) return 1;
  im2col_dump_kernel<<<1, 128, 16384 + 64 + 1024, static_cast<cudaStream_t>(stream)>>>(
      tm, c0, qq * stride - pad, pp * stride - pad, img, tap_s, tap_r, reinterpret_cast<uint4*>(d_out_16k));
  CK(cudaGetLastError());
  ctx->launches++;
  return 0;
}

// ---- probe: read-only / write-only HBM stream (the yardstick for kernels that only read - the match filter's bf16
// gallery - or only write - the stem's activations; MEASURED_PEAKS.json holds the read+write copy figure).
namespace {
__global__ void __launch_bounds__(256) stream_read_kernel(const uint4* __restrict__ p, size_t n16, unsigned* __restrict__ sink) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {   // four independent 16-byte loads in flight per thread
    const uint4 a = __ldcs(p + i), b = __ldcs(p + i + stride), c = __ldcs(p + i + 2 * stride), d = __ldcs(p + i + 3 * stride);
    acc.x ^= a.x ^ b.x ^ c.x ^ d.x; acc.y ^= a.y ^ b.y ^ c.y ^ d.y; acc.z ^= a.z ^ b.z ^ c.z ^ d.z; acc.w ^= a.w ^ b.w ^ c.w ^ d.w;
  }
  for (; i < n16; i += stride) {
    const uint4 a = __ldcs(p + i);
    acc.x ^= a.x; acc.y ^= a.y; acc.z ^= a.z; acc.w ^= a.w;
  }
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x9e3779b9u) *sink = acc.x;   // practically never: keeps the loads alive
}
__global__ void __launch_bounds__(256) stream_write_kernel(uint4* __restrict__ p, size_t n16, unsigned v) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n16; i += stride) __stcs(p + i, make_uint4(v, v, v, v));
}
}  // namespace

// mode 0: read `bytes` once per iteration, mode 1: write them; h_ms = average duration of one pass (CUDA events)
extern "C" int frb_debug_stream_bw(frb_ctx* ctx, void* d_buf, size_t bytes, int mode, int iters, float* h_ms) {
  if (!ctx || !d_buf || !h_ms || iters < 1) return 1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  unsigned* sink = nullptr;
  CK(cudaMalloc(&sink, 4));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  const size_t n16 = bytes / 16;
  const int grid = ctx->num_sms * 8;
  auto pass = [&]() {
    if (mode == 0) stream_read_kernel<<<grid, 256>>>(reinterpret_cast<const uint4*>(d_buf), n16, sink);
    else stream_write_kernel<<<grid, 256>>>(reinterpret_cast<uint4*>(d_buf), n16, 0x3c003c00u);
  };
  pass();
  CK(cudaEventRecord(a));
  for (int i = 0; i < iters; ++i) pass();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  CK(cudaGetLastError());
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, a, b));
  *h_ms = ms / iters;
  ctx->launches += iters + 1;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(sink);
  return 0;
}

// ---- probe: how does tcgen05.mma read a 128B-swizzled K-major A operand whose start is NOT 1024-byte
// aligned (row offset j0 into a TMA-written slab)?  mode 0: plain descriptor; mode 1: base_offset field set
// to address bits [7,10).  D[128][64] = slab[j0 : j0+128][0:64] * B[64][64]^T
namespace {
__global__ void shift_mma_kernel(const __grid_constant__ CUtensorMap tmSlab, const __grid_constant__ CUtensorMap tmB,
                                 int j0, int mode, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;            // 256 rows x 128 B
  uint8_t* sB = smem + 32768;    // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 8192);
  uint64_t* done = bar + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tptr, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 32768 + 8192);
    tma_load_2d(&tmSlab, bar, sA, 0, 0);
    tma_load_2d(&tmSlab, bar, sA + 16384, 0, 128);
    tma_load_2d(&tmB, bar, sB, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sA) + j0 * 128;
    uint64_t adesc = umma_desc_sw128(a_addr);
    if (mode == 1) adesc |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
    const uint64_t bdesc = umma_desc_sw128(smem_u32(sB));
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
    for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0 ? 1u : 0u);
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  {
    const int row = warp * 32 + lane;
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[row * 64 + c * 32 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}
}  // namespace

extern "C" int frb_debug_shift_mma(frb_ctx* ctx, const void* d_slab_256x64, const void* d_B_64x64, int j0, int mode,
                                   float* d_out_128x64, void* stream) {
  if (!ctx) return 1;
  if (j0 < 0 || j0 > 128) return fail(ctx, "j0 out of range");
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  CUtensorMap ta, tb;
  if (make_tmap_2d(ctx, &ta, d_slab_256x64, 64, 256, 128)) return 1;
  if (make_tmap_2d(ctx, &tb, d_B_64x64, 64, 64, 64)) return 1;
  if (set_smem_attr(ctx, reinterpret_cast<const void*>(shift_mma_kernel), 32768 + 8192 + 64 + 1024)) return 1;
  shift_mma_kernel<<<1, 128, 32768 + 8192 + 64 + 1024, static_cast<cudaStream_t>(stream)>>>(ta, tb, j0, mode, d_out_128x64);
  CK(cudaGetLastError());
  ctx->launches++;
  return 0;
}

// ---- probe: issue rate of tcgen05.mma (cta_group::1, operands resident in smem, one accumulator):
// cycles per instruction as seen by the issuing thread (issue) and until the commit lands (complete).
namespace {
template <int N>
__global__ void mma_rate_kernel(int iters, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;            // 128 x 128 B
  uint8_t* sB = smem + 16384;    // 256 x 128 B
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tptr, 256);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tptr, 0);
  if (warp == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    // mode >= 2: A starts (mode - 2) rows (128 B each) into the buffer, i.e. not on a 1024-byte swizzle-atom boundary
    const uint64_t adesc = umma_desc_sw128(smem_u32(sA) + (mode >= 2 ? (mode - 2) * 128 : 0));
    const uint64_t bdesc = umma_desc_sw128(smem_u32(sB));
    long long t0 = clock64();
    if (mode != 1) {            // whole loop inside one elected lane
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) umma_bf16_ss(tmem, adesc + 2 * (i & 3), bdesc + 2 * (i & 3), idesc, i > 0 ? 1u : 0u);
        umma_commit(done);
      }
    } else {                     // converged loop, elect per issue (CUTLASS style)
      for (int i = 0; i < iters; ++i) {
        const uint64_t a = adesc + 2 * (i & 3), b = bdesc + 2 * (i & 3);
        if (elect_one()) umma_bf16_ss(tmem, a, b, idesc, i > 0 ? 1u : 0u);
      }
      if (elect_one()) umma_commit(done);
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(done, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}
}  // namespace

// ---- same for CTA pairs (tcgen05 cta_group::2, M = 256): mode = A row offset (0 = aligned)
namespace {
template <int N>
__global__ void __cluster_dims__(2, 1, 1) mma2_rate_kernel(int iters, int a_row_off, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;            // 256 x 128 B (slack for row offsets)
  uint8_t* sB = smem + 32768;    // 128 x 128 B (half of N <= 256)
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 32768 + 16384);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const bool leader = cluster_ctarank() == 0;
  for (int i = threadIdx.x; i < (32768 + 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem2_alloc(tptr, 256);
    tmem2_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tptr, 0);
  if (warp == 0 && leader) {
    constexpr uint32_t idesc = umma_idesc_bf16(256, N);
    const uint64_t adesc = umma_desc_sw128(smem_u32(sA) + a_row_off * 128);
    const uint64_t bdesc = umma_desc_sw128(smem_u32(sB));
    long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < iters; ++i) umma2_bf16_ss(tmem, adesc + 2 * (i & 3), bdesc + 2 * (i & 3), idesc, i > 0 ? 1u : 0u);
      umma2_commit_pair(done);
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(done, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  } else if (warp == 0) {
    mbar_wait(done, 0);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem2_dealloc(tmem, 256);
  }
}
}  // namespace

extern "C" int frb_debug_mma2_rate(frb_ctx* ctx, int N, int iters, int a_row_off, long long* h_out2) {
  if (!ctx) return 1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  long long* d = nullptr;
  CK(cudaMalloc(&d, 16));
  const int smem = 32768 + 16384 + 64 + 1024;
  if (N == 64) {
    CK(cudaFuncSetAttribute(mma2_rate_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mma2_rate_kernel<64><<<2, 64, smem>>>(iters, a_row_off, d);
  } else if (N == 128) {
    CK(cudaFuncSetAttribute(mma2_rate_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mma2_rate_kernel<128><<<2, 64, smem>>>(iters, a_row_off, d);
  } else {
    CK(cudaFuncSetAttribute(mma2_rate_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mma2_rate_kernel<256><<<2, 64, smem>>>(iters, a_row_off, d);
  }
  CK(cudaGetLastError());
  CK(cudaMemcpy(h_out2, d, 16, cudaMemcpyDeviceToHost));
  CK(cudaFree(d));
  return 0;
}

extern "C" int frb_debug_mma_rate(frb_ctx* ctx, int N, int iters, int mode, long long* h_out2) {
  if (!ctx) return 1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  long long* d = nullptr;
  CK(cudaMalloc(&d, 16));
  const int smem = 16384 + 32768 + 64 + 1024;
  if (N == 64) {
    CK(cudaFuncSetAttribute(mma_rate_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mma_rate_kernel<64><<<1, 64, smem>>>(iters, mode, d);
  } else if (N == 128) {
    CK(cudaFuncSetAttribute(mma_rate_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mma_rate_kernel<128><<<1, 64, smem>>>(iters, mode, d);
  } else {
    CK(cudaFuncSetAttribute(mma_rate_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mma_rate_kernel<256><<<1, 64, smem>>>(iters, mode, d);
  }
  CK(cudaGetLastError());
  CK(cudaMemcpy(h_out2, d, 16, cudaMemcpyDeviceToHost));
  CK(cudaFree(d));
  return 0;
}
