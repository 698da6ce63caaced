/* frb200.h — C ABI of libfrb200.so: the B200 (sm_100a) embed-and-match hot path behind the
 * FaceRecognitionPipeline API.
 *
 * The reference is pure Python and has no FFI; the seams this ABI replaces are the library calls
 * its hot path makes (all citations into the reference tree):
 *   - face_recognition.py:61-75   FaceAligner.align -> cv2.warpAffine          => frb_warp_normalize
 *   - face_embedder.py:93-110     FaceEmbedder.preprocess (resize/BGR/normalise) => frb_preprocess_u8
 *   - face_embedder.py:49-56      net.build_model + load_state_dict             => frb_backbone_load
 *   - face_embedder.py:119,157    self.model(x) -> (features, norm)             => frb_embed
 *   - face_embedder.py:132-133,178-180  e / (||e|| + 1e-8)                     => frb_embed(flags RENORM)
 *   - gallery_manager.py:177-187  get_gallery_embeddings (np.vstack per query)  => frb_gallery_upload (once)
 *   - gallery_manager.py:189-205  search: normalise q, np.dot, argsort top-k    => frb_match
 *   - face_matcher.py:205         score >= similarity_threshold                 => frb_match accept[]
 *   (new, no reference counterpart) identity-sharded gallery merge              => frb_topk_merge
 *
 * Conventions: every call returns 0 on success, non-zero on failure (message via
 * frb_last_error).  No exceptions, no Python objects, no torch types cross this boundary.
 * Pointers named d_* are CUDA device pointers owned by the caller; h_* are host pointers.
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  A ctx is bound to
 * one device and serialises its own calls with an internal mutex (the reference is called
 * concurrently under Flask threaded=True, face_recognition_server.py:1102).
 */
#ifndef FRB200_H
#define FRB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct frb_ctx frb_ctx;

#define FRB_OP_STEM 0 /* Conv3x3(3->64)+BN+PReLU, direct */
#define FRB_OP_CONV 1 /* implicit-GEMM conv (3x3 or 1x1) with fused epilogue */
#define FRB_OP_FC 2   /* BN2d-Flatten-Linear-BN1d folded into one GEMM + L2 norm */

#define FRB_EMBED_L2 1     /* model L2-normalises its output (AdaFace Backbone.forward) */
#define FRB_EMBED_RENORM 2 /* apply FaceEmbedder's extra e/(||e||+1e-8) (normalize=True) */
#define FRB_EMBED_FLIP 4   /* input holds 2B crops (orig, hflip): fuse pairs -> B embeddings */

/* One step of the backbone program.  The host side (weights.py) folds BatchNorm into the
 * weights/biases, orders K as (tap, cin) [+ shortcut cin] and assigns activation buffers. */
typedef struct frb_layer_desc {
  int32_t op;
  int32_t cin, cout;
  int32_t hin, win;           /* input spatial size */
  int32_t ksize, stride, pad; /* 3/1, 1|2, 1/0 */
  int32_t in_buf, out_buf;    /* activation buffer ids; in_buf = -1 means the network input */
  int32_t sc_buf, sc_cin, sc_hin, sc_win, sc_stride; /* fused 1x1 shortcut conv source; sc_buf < 0: none */
  int32_t res_buf, res_h, res_w, res_stride;         /* identity (MaxPool(1,stride)) shortcut; res_buf < 0: none */
  int32_t bias_cases;         /* 1, or 9 = border-position table (pre-activation BN under zero padding) */
  int32_t has_prelu;
  int32_t reserved[2];
  int64_t w_off, w_bytes;     /* bf16 [cout][ktot] (CONV/FC) or f32 [27][64] (STEM) inside the blob */
  int64_t bias_off;           /* f32 [bias_cases][cout] */
  int64_t prelu_off;          /* f32 [cout] (ignored unless has_prelu) */
} frb_layer_desc;

/* Per-face warp job: forward 2x3 matrix as cv2.estimateAffinePartial2D returns it. */
typedef struct frb_warp_job {
  uint64_t src_off; /* byte offset of this face's source image from d_src_base */
  int32_t H, W, pitch;
  int32_t _pad;
  double M[6];
} frb_warp_job;

int frb_ctx_create(int device, frb_ctx** out);
void frb_ctx_destroy(frb_ctx* ctx);
const char* frb_last_error(frb_ctx* ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
long long frb_launch_count(frb_ctx* ctx);

/* ---- preprocessing ---- */
/* d_in: [B][S][S][3] RGB u8, S in {112,224}; d_out: [B (2B if flip)][112][112][3] bf16 BGR normalised */
int frb_preprocess_u8(frb_ctx* ctx, const void* d_in, int B, int S, void* d_out, int flip, void* stream);
/* jobs on the HOST (copied inside); S = output size; either output may be NULL.
 * d_out_u8: [B][S][S][3] RGB u8 (== FaceAligner.align); d_out_bf16: [B][112][112][3] (S must be 112) */
int frb_warp_normalize(frb_ctx* ctx, const void* d_src_base, const frb_warp_job* h_jobs, int B, int S,
                       void* d_out_u8, void* d_out_bf16, void* stream);

/* ---- backbone ---- */
int frb_backbone_load(frb_ctx* ctx, const frb_layer_desc* layers, int n_layers, const void* h_blob,
                      size_t blob_bytes, int n_bufs);
/* d_in: [B'][112][112][3] bf16 (B' = 2B with FRB_EMBED_FLIP); outputs [B][512] f32, [B] f32, [B][512] bf16;
 * any output may be NULL. */
int frb_embed(frb_ctx* ctx, const void* d_in, int B, int flags, float* d_emb, float* d_norm,
              void* d_emb_bf16, void* stream);
double frb_backbone_flops_per_face(frb_ctx* ctx);
/* schedule of the last embed: out4 = {faces per pass of the layer program (large batches are chunked, default 256),
 * number of leading "front" layers run in sub-batches (stem + the 112/56-pixel Cout-64 layers), images per sub-batch
 * (default 32: their activations then stay in L2 between layers), sub-batches per pass}.  FRB_EMBED_CHUNK / FRB_FRONT_SUB. */
int frb_embed_schedule(frb_ctx* ctx, int* out4);
/* measurement aid (bench.py roofline): frb_embed with a CUDA event between consecutive layers; per layer the
 * duration (ms), the kernel that ran it (0 stem, 1/2/3 CTA-pair im2col conv with Cout tile 64/128/256,
 * 4/5/6 slab conv <64,1>/<128,1>/<128,2>, 7 FC + finalize) and its FLOPs for this batch. */
int frb_embed_profile(frb_ctx* ctx, const void* d_in, int B, int flags, float* d_emb, void* stream,
                      float* h_layer_ms, int* h_kernel_id, double* h_layer_flops, int max_layers, int* n_layers);

/* ---- gallery ---- */
/* g: [N][512] f32 rows (NOT re-normalised, as in search()); is_device selects the pointer kind.
 * first_global_id is added to local row numbers in every result (identity-sharded galleries). */
int frb_gallery_upload(frb_ctx* ctx, const float* g, long long N, long long first_global_id, int is_device);
long long frb_gallery_size(frb_ctx* ctx);
/* Residency generations: every frb_gallery_upload* / frb_backbone_load bumps the ctx's counter.  A wrapper that shares a
 * ctx remembers the value after ITS upload and re-uploads when it no longer matches (replaces Python-side id() tokens). */
long long frb_gallery_generation(frb_ctx* ctx);
long long frb_backbone_generation(frb_ctx* ctx);
/* ---- per-identity matching over a gallery of SAMPLES (evaluate_models_v2.ipynb cells 3-5: compute_all_similarities,
 * aggregate_max / aggregate_mean / aggregate_topk, identify_probe) ----
 * samples: [T][512] f32 rows (host or device); identity i owns rows [h_seg[i], h_seg[i+1]) (h_seg: S+1 int64 on the
 * HOST, h_seg[0] = 0, h_seg[S] = T, at most 64 samples per identity).  Replaces the resident gallery. */
int frb_gallery_upload_samples(frb_ctx* ctx, const float* samples, long long T, const long long* h_seg, long long S,
                               int is_device);
/* identity_scores of identify_probe for P probes: d_out [P][S] f32; agg 0 = max, 1 = mean, 2 = mean of the agg_k best;
 * an identity without samples scores -1.  Probes are normalised like search() when normalize != 0.
 * S_expected: the identity count the caller sized d_out for (fails when another gallery is resident; < 0 = unchecked). */
int frb_identity_scores(frb_ctx* ctx, const float* d_probes, int P, int normalize, int agg, int agg_k, float* d_out,
                        long long S_expected, void* stream);
/* top-k identities per probe ranked by (score desc, identity index asc): scores f32 [P][k], idx i64 [P][k] (identity
 * indices, -1 = fewer than k identities), accept u8 [P] (best score >= thr).  Exact by construction: tensor-core
 * filter over the samples, exact f64 aggregates for the candidate identities, proof, exact scan for unproven rows. */
int frb_match_identities(frb_ctx* ctx, const float* d_probes, int P, int k, float thr, int normalize, int agg, int agg_k,
                         float* d_scores, long long* d_idx, unsigned char* d_accept, void* stream);

/* Enrollment aggregation for S identities at once (GalleryManager._aggregate_embeddings incl. the quality filter,
 * gallery_manager.py:104-122,297-317): identity s owns rows [d_seg[s], d_seg[s+1]) of d_emb [T][512] f32 (device;
 * d_seg is S+1 int64 on the device; at most max_rows <= 64 rows each).  method: 0 mean, 1 median, 2 weighted_mean.
 * Outputs (device): d_templates [S][512] f32 (ready for frb_gallery_upload(is_device=1)), d_kept [S] rows that
 * survived the quality filter (may be NULL). */
int frb_aggregate_templates(frb_ctx* ctx, const float* d_emb, const long long* d_seg, int S, int max_rows, int method,
                            float min_similarity, float* d_templates, int* d_kept, void* stream);
/* ---- track-level consensus (FaceMatcher._aggregate_matches, face_matcher.py:321-363, and _get_best_candidate,
 * :365-385) for T tracks at once.  Frame f of the stream is described by its top-1 match as frb_match wrote it:
 * d_top_idx[f * k_stride] (gallery row, < 0 = the frame had no match and is skipped like face_matcher.py:180-181) and
 * d_top_score[f * k_stride]; track t owns frames [d_seg[t], d_seg[t+1]) (T+1 int64 on the device, at most max_frames
 * <= 1024 each).  Frames scoring >= min_quality (0.55) vote; >= min_frames (3) voters; the winner needs > 50 % of the
 * votes or > 40 % and twice the runner-up (ties: first occurrence, as Counter.most_common); its mean score (numpy's
 * float64 pairwise sum) must reach `threshold`. */
typedef struct frb_track_result {
  long long winner;               /* consensus gallery row, -1 = not recognised */
  double confidence;              /* 'confidence' */
  double consensus_strength;      /* 'consensus_strength' */
  int num_quality_frames;         /* 'num_quality_frames' */
  int total_frames_evaluated;     /* 'total_frames_evaluated' */
  long long candidate;            /* _get_best_candidate: 'student_id' row (-1: no matched frame) */
  double candidate_confidence;
  int candidate_num_quality_frames;
  int recognized;
} frb_track_result;
int frb_track_consensus(frb_ctx* ctx, const long long* d_top_idx, const float* d_top_score, int k_stride,
                        const long long* d_seg, int T, int max_frames, double min_quality, int min_frames,
                        double threshold, frb_track_result* d_out, void* stream);
/* ---- server best-frame selection (LiveRecognitionTracker.get_best_frame and the should_recognize gate,
 * face_recognition_server.py:39-85) for T tracks at once: track t owns frames [d_seg[t], d_seg[t+1]) (T+1 int64 on the
 * device); quality = det_score * min(blur_score / 100, 1) in f64; the FIRST frame with the largest quality wins (Python
 * max).  Outputs (device): best_idx i64 [T] (frame index inside the track, -1 = empty track), best_quality f64 [T]
 * (optional), ready u8 [T] (the best frame's det_score > min_det; the reference uses 0.6). */
int frb_best_frames(frb_ctx* ctx, const double* d_det, const double* d_blur, const long long* d_seg, int T,
                    double min_det, long long* d_best_idx, double* d_best_quality, unsigned char* d_ready, void* stream);
/* d_probes: [P][512] f32.  normalize != 0 applies q/(||q||+1e-8) first (search()).
 * Outputs (device): scores f32 [P][k], idx i64 [P][k] (global ids, -1 = fewer than k rows),
 * accept u8 [P] (top-1 score >= thr), scores64 f64 [P][k] (optional, for cross-rank merge).
 * Any k >= 1 (search() accepts any top_k, gallery_manager.py:197): k <= 32 on galleries of >= 4096 rows runs the
 * tensor-core filter + exact re-score + proof, everything else the dense exact scan.
 * Enqueue only: no host synchronisation, no device-to-host copy the caller has to wait for (CUDA-graph capturable
 * once the workspaces exist, i.e. from the second call with the same P / k on). */
int frb_match(frb_ctx* ctx, const float* d_probes, int P, int k, float thr, int normalize,
              float* d_scores, long long* d_idx, unsigned char* d_accept, double* d_scores64, void* stream);
/* measurement aid (bench.py per-kernel roofline): frb_match with CUDA events between prepare | filter kernel |
 * finalize + exact fix-up; h_ms3 receives the three durations (ms).  Synchronises. */
int frb_match_profile(frb_ctx* ctx, const float* d_probes, int P, int k, float thr, int normalize, float* d_scores,
                      long long* d_idx, unsigned char* d_accept, void* stream, float* h_ms3);
/* how many probes of the last frb_match needed the exact re-scan (filter proof failed); waits for that match */
int frb_match_last_flagged(frb_ctx* ctx);
/* frb_match with the result as 16-byte records [P][k] = (f64 score, i64 global id): the payload ONE all-gather moves */
int frb_match_packed(frb_ctx* ctx, const float* d_probes, int P, int k, float thr, int normalize, void* d_records,
                     void* stream);
/* merge G gathered per-rank top-k lists: in [G][P][k] -> out [P][k] (canonical order: score desc, id asc) */
int frb_topk_merge(frb_ctx* ctx, const double* d_in_scores64, const long long* d_in_idx, int G, int P, int k,
                   float thr, float* d_scores, long long* d_idx, unsigned char* d_accept, double* d_scores64,
                   void* stream);
int frb_topk_merge_packed(frb_ctx* ctx, const void* d_records, int G, int P, int k, float thr, float* d_scores,
                          long long* d_idx, unsigned char* d_accept, void* stream);

/* ---- identity-sharded gallery over peer memory (NVLink P2P; new work, the reference is single-device) ----
 * One process per GPU.  frb_xchg_create allocates this rank's exchange buffer (probes of all ranks, two parities of
 * result slots, flag words) and returns its 64-byte cudaIpc handle; the caller all-gathers the handles (any transport:
 * torch.distributed, files) and passes all of them to frb_xchg_connect.  world <= 8, max_k <= 32. */
int frb_xchg_create(frb_ctx* ctx, int world, int rank, int max_probes, int max_k, void* h_handle_out64);
int frb_xchg_connect(frb_ctx* ctx, const void* h_handles_all);
/* 0, or 1 + the rank a device-side wait gave up on (FRB_XCHG_TIMEOUT_MS, default 20 s; the kernel then traps) */
int frb_xchg_status(frb_ctx* ctx);
/* Collective: every rank passes its own rows [p_lo, p_lo + p_cnt) of the P_total probes (d_local_probes [p_cnt][512]
 * f32) and holds its shard (frb_gallery_upload with first_global_id = shard start).  Probes are stored into every
 * rank's buffer by the prepare kernel, each rank matches ALL probes against its shard, every finished row's k records
 * go straight into every rank's result slot from the finalize kernel, and a merge kernel that waits on the peers'
 * flags writes the global top-k of ALL P_total probes on every rank.  No NCCL and no host synchronisation on the data
 * path; results equal frb_match against the unsharded gallery bit for bit. */
int frb_match_sharded(frb_ctx* ctx, const float* d_local_probes, int p_lo, int p_cnt, int P_total, int k, float thr,
                      int normalize, float* d_scores, long long* d_idx, unsigned char* d_accept, void* stream);

/* ---- host-buffer entry points (what the Python drop-in calls; copies happen inside) ---- */
/* h_rgb: [B][S][S][3] u8 -> h_emb [B][512] f32 (+ h_norm [B] optional) */
int frb_embed_host(frb_ctx* ctx, const uint8_t* h_rgb, int B, int S, int flags, float* h_emb, float* h_norm);
/* h_probes [P][512] f32 -> h_scores [P][k], h_idx [P][k], h_accept [P] */
int frb_match_host(frb_ctx* ctx, const float* h_probes, int P, int k, float thr, int normalize,
                   float* h_scores, long long* h_idx, unsigned char* h_accept);
/* Streams of batches: start copying the crops of a LATER frb_embed_host / frb_embed_match_host call now (separate copy
 * stream), so the transfer overlaps the call issued in between.  The later call recognises the same h_rgb / B / S and
 * skips its own copy; the buffer must stay unchanged until then.  Two prefetches may be outstanding (the batch the next
 * call takes and the one after it); a third overwrites the older. */
int frb_prefetch_host(frb_ctx* ctx, const uint8_t* h_rgb, int B, int S);
/* whole path: aligned crops -> embed -> match against the uploaded gallery */
int frb_embed_match_host(frb_ctx* ctx, const uint8_t* h_rgb, int B, int S, int flags, int k, float thr,
                         float* h_emb, float* h_scores, long long* h_idx, unsigned char* h_accept);

/* ---- test hooks (used by tests/ only) ---- */
/* ranks living in one process (one ctx each) hand each other the raw exchange buffer instead of a cudaIpc handle */
void* frb_xchg_local_buffer(frb_ctx* ctx);
int frb_xchg_connect_local(frb_ctx* ctx, int peer_rank, void* d_peer_buffer);
/* C[M][N] f32 = A[M][K] * B[N][K]^T, bf16 inputs, K % 64 == 0, N % 64 == 0 */
int frb_debug_gemm(frb_ctx* ctx, const void* d_A, const void* d_B, int M, int N, int K, int splits,
                   float* d_C, void* stream);
/* single conv layer through the tcgen05 path (use_ref = 0) or the direct checker (use_ref = 1) */
int frb_debug_conv(frb_ctx* ctx, const frb_layer_desc* layer, int B, const void* d_in, const void* d_sc,
                   const void* d_res, const void* d_w, const float* d_bias, const float* d_prelu,
                   void* d_out, int use_ref, void* stream);
/* raw im2col-mode TMA tile (128 pixels x 64 channels, swizzled as landed) for inspection */
int frb_debug_im2col(frb_ctx* ctx, const void* d_in, int B, int H, int W, int C, int ksize, int stride, int pad,
                     int m0, int c0, int tap_r, int tap_s, void* d_out_16k, void* stream);

/* probe: read-only (mode 0) / write-only (mode 1) HBM stream over d_buf; h_ms = average pass duration.  The yardstick for
 * kernels that only read (match filter) or only write (stem); MEASURED_PEAKS.json holds the read+write copy figure. */
int frb_debug_stream_bw(frb_ctx* ctx, void* d_buf, size_t bytes, int mode, int iters, float* h_ms);

/* probe: tcgen05.mma over a 128B-swizzled A operand starting j0 rows into a TMA-written slab
 * (mode 1 sets the descriptor's base_offset field).  D[128][64] = slab[j0:j0+128] * B^T */
int frb_debug_shift_mma(frb_ctx* ctx, const void* d_slab_256x64, const void* d_B_64x64, int j0, int mode,
                        float* d_out_128x64, void* stream);

/* probe: tcgen05.mma issue / completion cycles for `iters` back-to-back 128xNx16 MMAs (h_out2[0] = issue
 * cycles, h_out2[1] = cycles until the commit lands); mode 0 = loop inside one lane, 1 = elect per issue */
int frb_debug_mma_rate(frb_ctx* ctx, int N, int iters, int mode, long long* h_out2);
/* same for CTA pairs (cta_group::2, 256 x N x 16); a_row_off = A start offset in 128-byte rows */
int frb_debug_mma2_rate(frb_ctx* ctx, int N, int iters, int a_row_off, long long* h_out2);

#ifdef __cplusplus
}
#endif
#endif /* FRB200_H */
