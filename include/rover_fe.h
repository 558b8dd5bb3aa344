/* rover_fe.h -- C ABI of the B200-native Rover-SLAM feature front end (SuperPoint extract + LightGlue match).
 *
 * This is the drop-in boundary for the reference's ONNXRuntime calls.  Each entry point names the
 * reference interface it replaces (paths relative to the Rover-SLAM repository):
 *
 *   rfe_create / rfe_destroy   <- SuperPointOnnxRunner::InitOrtEnv      src/Extractors/superpoint_onnx.cc:4-66
 *                                 LightGlueDecoupleOnnxRunner::InitOrtEnv src/Matchers/lightglue_onnx.cpp:4-98
 *   rfe_sp_extract_u8          <- NormalizeImage                         src/Matchers/transform.cpp:3-17
 *                                 + SuperPointOnnxRunner::Extractor_Inference   superpoint_onnx.cc:88-162
 *                                 + the tensor unpacking of Extractor_PostProcess superpoint_onnx.cc:165-255
 *   rfe_lg_match               <- LightGlueDecoupleOnnxRunner::Matcher_PreProcess lightglue_onnx.cpp:140-159
 *                                 + Matcher_Inference                    lightglue_onnx.cpp:162-240 / 241-330
 *                                 + the score>thresh filter of Matcher_PostProcess_fused lightglue_onnx.cpp:437-453
 *   rfe_get_timer_ms           <- SuperPointOnnxRunner::GetTimer         superpoint_onnx.cc:268-277
 *
 * Conventions (same as the reference's runners): return codes, never exceptions (EXIT_SUCCESS == RFE_OK);
 * plain pointers and sizes; caller owns every output buffer; one rfe_ctx per thread (a ctx is NOT
 * re-entrant, several ctxs may share a GPU).  All arithmetic runs on the GPU; there is no CPU fallback:
 * rfe_create fails when no sm_100 device is present.
 *
 * Outputs follow the ONNX graphs' conventions exactly: keypoints are integer pixel centres (x, y) in
 * row-major (y, then x) order; descriptors are unit-norm 256-vectors in keypoint order; matches are
 * (index into set 0, index into set 1) pairs ascending in the first index.
 */
#ifndef ROVER_FE_H_
#define ROVER_FE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RFE_OK 0
#define RFE_ERR_INVALID 1   /* bad argument (null pointer, size not a multiple of 8, batch too large ...) */
#define RFE_ERR_CUDA 2      /* CUDA runtime / driver error; see rfe_last_error() */
#define RFE_ERR_IO 3        /* weight blob missing or malformed */
#define RFE_ERR_CAPACITY 4  /* more keypoints than the ctx / caller capacity; outputs hold the first `cap` */
#define RFE_ERR_NO_DEVICE 5 /* no sm_100 GPU */

#define RFE_DESC_DIM 256

typedef struct rfe_ctx rfe_ctx;

typedef struct rfe_config {
  int device;               /* CUDA device ordinal */
  void* stream;             /* cudaStream_t to enqueue on; NULL = the ctx creates its own stream */
  const char* weights_path; /* RFW1 blob; NULL = $ROVER_FE_WEIGHTS, else "weights/rover_fe.rfw" */
  int max_batch;            /* images per rfe_sp_* call            (0 = 8) */
  int max_height;           /* largest image height, multiple of 8 (0 = 480) */
  int max_width;            /* largest image width, multiple of 8  (0 = 768) */
  int max_keypoints;        /* per-image keypoint capacity         (0 = 4096) */
  int flags;                /* RFE_FLAG_* (0 = a ctx that both extracts and matches) */
} rfe_config;
/* A runner that only ever extracts (SuperPointOnnxRunner) or only ever matches (LightGlueDecoupleOnnxRunner) skips the other
 * half's device buffers: the LightGlue state of an 8192-keypoint ctx is ~0.6 GB, the SuperPoint activations of a 1024x1280
 * ctx ~1.3 GB.  Calls into the disabled half return RFE_ERR_INVALID. */
#define RFE_FLAG_NO_MATCHER 1
#define RFE_FLAG_NO_EXTRACTOR 2

int rfe_create(const rfe_config* cfg, rfe_ctx** out);
void rfe_destroy(rfe_ctx* ctx);
/* Thread-local text of the last failure on this thread. */
const char* rfe_last_error(void);
/* Block until everything enqueued on the ctx stream has finished. */
int rfe_sync(rfe_ctx* ctx);

/* ---- SuperPoint ------------------------------------------------------------------------------- */
/* Host in / host out.  gray: batch images of h x w uint8 (CV_8UC1), row pitch stride_bytes, image pitch
 * h*stride_bytes.  For image b, counts[b] = N_b keypoints and the first min(N_b, cap) rows of
 * kpts_xy[b*cap ...], scores[b*cap ...], desc[b*cap*256 ...] are written.  scores/desc may be NULL.
 * h, w multiples of 8 (the graph reshapes to [h/8, w/8, 8, 8]).  Returns RFE_ERR_CAPACITY if any N_b > cap. */
int rfe_sp_extract_u8(rfe_ctx* ctx, const uint8_t* gray, int h, int w, int stride_bytes, int batch,
                      int32_t* kpts_xy, float* scores, float* desc, int32_t* counts, int cap);

/* Optional cap on the keypoints per image: the reference's SPextractor stores `nfeatures` and never uses
 * it (src/Extractors/SPextractor.cc:84-146; the ONNX graph has no top-K).  k > 0: every later extraction keeps the k
 * highest-scoring keypoints of each image, still in the graph's row-major order (among equal scores the earlier keypoint
 * stays); descriptors are sampled for the kept ones only.  k <= 0 (default): every keypoint, i.e. the reference's behaviour. */
int rfe_sp_set_topk(rfe_ctx* ctx, int k);

/* LABELLED FAST MODE -- not the parity path, never the default.  on != 0: SuperPoint's 3x3 convolutions (conv1b .. convPa/Da,
 * 98 % of the extractor's FLOPs) issue only the hi*hi product of the split-fp16 scheme, i.e. plain fp16 operands with fp32
 * accumulation: one tensor-core MMA per MAC instead of three.  Keypoints near the detection threshold and NMS near-ties
 * change (fp16 inputs lose about 1 % of the keypoints); bench.py reports the mode separately with the
 * keypoint and match overlap against the exact path.  LightGlue is not affected.  on == 0 (default): the reference's fp32
 * arithmetic. */
int rfe_set_fast_mode(rfe_ctx* ctx, int on);

/* The tensor-core kernels are persistent: one CTA per SM walks a static list of tiles, so a kernel that finds one SM taken by
 * somebody else's long-running kernel (an NCCL collective waiting for its peers) runs its last CTA after all the others and
 * takes twice as long.  A process that overlaps communication with the front end (bench.py's config-5 stream)
 * limits the front end to max_sms SMs and leaves the rest to the communication kernels.  0 = all SMs (default). */
int rfe_set_sm_limit(rfe_ctx* ctx, int max_sms);

/* Device in / device-resident out (asynchronous on the ctx stream).  d_gray: device pointer, layout as
 * above.  Features stay in the ctx ("slots" 0..batch-1) for rfe_lg_match_slots / rfe_sp_read_slot. */
int rfe_sp_extract_device(rfe_ctx* ctx, const uint8_t* d_gray, int h, int w, int stride_bytes, int batch);
/* Copy slot b's features to the host (synchronises).  Any output pointer may be NULL. */
int rfe_sp_read_slot(rfe_ctx* ctx, int slot, int32_t* kpts_xy, float* scores, float* desc, int32_t* count, int cap);

/* Upload features that are already on the host into slot b (synchronises): the inverse of rfe_sp_read_slot.  This is how a
 * stored KeyFrame's keypoints + descriptors re-enter the device-resident path: the reference copies
 * mDescriptors / mvKeysUn of BOTH frames host->device on every MatchingPoints_onnx call (src/Matchers/SPmatcher.cc:498-528,
 * `new float[]` + Ort tensors); here a KeyFrame is uploaded once and matched against many frames with rfe_lg_match_slots*.
 * slot < max_batch, n <= max_keypoints; scores may be NULL (stored as 0).  Slots below `slot` that were never written hold 0
 * keypoints.  The next rfe_sp_extract_* call overwrites slots 0..batch-1. */
int rfe_sp_write_slot(rfe_ctx* ctx, int slot, const int32_t* kpts_xy, const float* scores, const float* desc, int n);

/* Sign-binarised descriptors of slot b: what Frame::binarize_descriptors (src/Frame.cc:1034-1043) and
 * KeyFrame::binarize_descriptors (src/KeyFrame.cc:113-123) compute on the host with cv::threshold(row, 0, 1, THRESH_BINARY)
 * before every DBoW3 transform -- here written by the descriptor sampler while the values are in registers.
 * bin: [n][256] uchar, 1 where the descriptor element is > 0 (the CV_8UC1 layout of mDescriptors_bin). */
int rfe_sp_read_slot_bin(rfe_ctx* ctx, int slot, uint8_t* bin, int32_t* count, int cap);
/* The same operation for host descriptors that did not come from a slot.  bin [n][256] 0/1 and/or bits [n][8] uint32
 * (bit j%32 of word j/32 = element j); either may be NULL. */
int rfe_binarize_descriptors(rfe_ctx* ctx, const float* desc, int n, uint8_t* bin, uint32_t* bits);

/* ---- L2 projection matching -------------------------------------------------------- */
/* Best and second-best L2 distance of every query descriptor over ITS OWN candidate list: the inner loop of
 * SPmatcher::SearchByProjection / SearchByProjection1 / Fuse and Frame::ComputeStereoMatches (src/Matchers/SPmatcher.cc
 * 1225-1250, distance = cv::norm(a, b, NORM_L2), SPmatcher.cc:2184-2189).  q [nq][256], db [nd][256]; query i examines
 * db rows cand_idx[cand_off[i] .. cand_off[i+1]) in that order with the reference's update rule (strict <, so the first
 * of equal distances wins; both distances start at init_dist, 256 in the reference).  best_idx / second_idx are -1 when
 * no candidate beat init_dist.  second_dist / second_idx may be NULL. */
int rfe_l2_best2(rfe_ctx* ctx, const float* q, int nq, const float* db, int nd, const int32_t* cand_off,
                 const int32_t* cand_idx, float init_dist, float* best_dist, int32_t* best_idx, float* second_dist,
                 int32_t* second_idx);

/* The slot-resident form: the database is feature slot db_slot (left on the device by rfe_sp_extract_device /
 * rfe_pairs_submit, or uploaded once with rfe_sp_write_slot); the queries are either the first nq descriptors of feature
 * slot q_slot (q_host == NULL: frame against frame, e.g. left against right in Frame::ComputeStereoMatches,
 * src/Frame.cc:1159-1340) or nq host descriptors q_host [nq][256] (MapPoint descriptors against the current frame in
 * SPmatcher::SearchByProjection1, src/Matchers/SPmatcher.cc:1170-1352; q_slot is ignored).  The frame's descriptors never
 * cross PCIe again.  INTEGRATION.md shows SearchByProjection1 written on it. */
int rfe_l2_best2_slots(rfe_ctx* ctx, const float* q_host, int q_slot, int db_slot, int nq, const int32_t* cand_off,
                       const int32_t* cand_idx, float init_dist, float* best_dist, int32_t* best_idx, float* second_dist,
                       int32_t* second_idx);

/* ---- LightGlue -------------------------------------------------------------------------------- */
/* Host in / host out.  kpts*_px: [n][2] pixel coordinates (x, y); desc*: [n][256].  Keypoints are
 * normalised as the reference does, (kpt - (norm_w/2, norm_h/2)) / (max(norm_w, norm_h)/2).
 * Writes k pairs with mscore > match_thresh (the ONNX graph already drops mscore <= 0.1):
 * matches[2*i], matches[2*i+1], mscores[i]; capacity of both arrays: n0 entries. */
int rfe_lg_match(rfe_ctx* ctx, const float* kpts0_px, int n0, const float* kpts1_px, int n1, const float* desc0,
                 const float* desc1, int norm_h, int norm_w, float match_thresh, int32_t* matches, float* mscores,
                 int* k);

/* The same with keypoints the caller has ALREADY normalised -- exactly the tensors the reference feeds its session
 * (LightGlueDecoupleOnnxRunner::Matcher_Inference takes the output of Matcher_PreProcess / NormalizeKeypoints,
 * src/Matchers/lightglue_onnx.cpp:162-240 and :241-330; src/Matchers/transform.cpp:19-32).  The device applies no further
 * normalisation, so the host runner is stateless between Matcher_PreProcess and Matcher_Inference like the reference's. */
int rfe_lg_match_normalized(rfe_ctx* ctx, const float* kpts0_norm, int n0, const float* kpts1_norm, int n1,
                            const float* desc0, const float* desc1, float match_thresh, int32_t* matches, float* mscores,
                            int* k);

/* Match two feature slots left on the device by rfe_sp_extract_device (asynchronous); the result
 * stays on the device in result slot `rslot` (0 .. max_batch-1). */
int rfe_lg_match_slots(rfe_ctx* ctx, int slot0, int slot1, int norm_h, int norm_w, float match_thresh, int rslot);
/* Match n_pairs (<= max_batch) independent pairs of feature slots in ONE pass: every LightGlue linear layer runs as a
 * single GEMM over the rows of all pairs and attention as one fused launch over 2*n_pairs problems.  Pair i's result
 * goes to result slot i.  Asynchronous apart from one small device->host read of the keypoint counts. */
int rfe_lg_match_slots_batch(rfe_ctx* ctx, int n_pairs, const int* slot0, const int* slot1, int norm_h, int norm_w,
                             float match_thresh);
/* One feature slot against many: pair i = (slot, others[i]), result slot i -- LocalMapping's
 * CreateNewMapPoints matches ONE KeyFrame against up to ten covisible neighbours (src/LocalMapping.cc:522-634 ->
 * SPmatcher::SearchForTriangulation -> MatchingPoints_onnx per neighbour, src/Matchers/SPmatcher.cc:1355-1399), re-uploading
 * and re-projecting that KeyFrame every time.  Here everything that depends on an image alone -- positional encoding and the
 * whole first self-attention block of LightGlue -- is computed once per (slot contents, norm_h, norm_w) and cached on the
 * device; the cache entry is rebuilt when the slot is overwritten (extraction, rfe_sp_write_slot) or matched with another
 * normalisation size.  Results equal rfe_lg_match_slots_batch on the same pairs.  n_others <= max_batch. */
int rfe_lg_match_one_to_many(rfe_ctx* ctx, int slot, const int* others, int n_others, int norm_h, int norm_w,
                             float match_thresh);
/* How many slot states rfe_lg_match_one_to_many found cached / had to build so far. */
int rfe_lg_cache_stats(rfe_ctx* ctx, long long* hits, long long* builds);

/* The whole hot path in one call, host in / host out: n_pairs pairs of images (pair p = images 2p and 2p+1 of `gray`,
 * layout as rfe_sp_extract_u8), SuperPoint on all 2*n_pairs images, LightGlue on every pair (keypoints normalised with
 * the image size, i.e. the semantics of MatchingPoints_onnx(Frame&, Frame&, ...), SPmatcher.cc:457-542).  Descriptors
 * never leave the GPU.  Outputs: kp_counts[2*n_pairs], kpts_xy[2*n_pairs][cap][2] (may be NULL),
 * match_counts[n_pairs], matches[n_pairs][cap][2], mscores[n_pairs][cap] (may be NULL). */
int rfe_match_pairs_u8(rfe_ctx* ctx, const uint8_t* gray, int h, int w, int stride_bytes, int n_pairs, float match_thresh,
                       int32_t* kpts_xy, int32_t* kp_counts, int32_t* matches, float* mscores, int32_t* match_counts,
                       int cap);
/* Pipelined form of rfe_match_pairs_u8 for a STREAM of batches (Tracking feeds frames continuously).  submit() copies
 * the images to the device and enqueues SuperPoint; collect() returns the results of the OLDEST submitted batch: it waits
 * for that batch's keypoint counts, enqueues LightGlue, copies keypoints and matches back.  Up to two batches may be in
 * flight -- submit(A) submit(B) collect(A) submit(C) collect(B) ... -- so the GPU runs SuperPoint of the next batch while
 * the host lays out the matcher of the previous one.  `gray` must stay valid until the matching collect() returns (use
 * pinned memory for a truly asynchronous copy).  Output layout as rfe_match_pairs_u8; matches/mscores receive up to n0
 * entries per pair (the first match_counts[p] are valid). */
int rfe_pairs_submit(rfe_ctx* ctx, const uint8_t* gray, int h, int w, int stride_bytes, int n_pairs);
int rfe_pairs_collect(rfe_ctx* ctx, float match_thresh, int32_t* kpts_xy, int32_t* kp_counts, int32_t* matches,
                      float* mscores, int32_t* match_counts, int cap);
/* rfe_pairs_collect in two halves, so that the NEXT batch can be submitted before the host blocks: begin() enqueues the
 * matcher and the result copies (the output arrays must stay valid until end()), end() waits for exactly those copies.
 * Steady state of a stream:  collect_begin(i); submit(i+2); collect_end(i);  -- the GPU queue never drains. */
int rfe_pairs_collect_begin(rfe_ctx* ctx, float match_thresh, int32_t* kpts_xy, int32_t* matches, float* mscores, int cap);
int rfe_pairs_collect_end(rfe_ctx* ctx, int32_t* kp_counts, int32_t* match_counts);
/* collect_begin that also returns what SPextractor::operator() returns besides the keypoints -- scores[2*n_pairs][cap] and
 * the fp32 descriptors desc[2*n_pairs][cap][256] (Frame::mDescriptors, src/Frame.cc:544-559); either may be NULL.  Exactly
 * counts[b] rows per image are copied, on a second stream, overlapping the matcher; rfe_pairs_collect_end waits for them
 * too.  Use pinned memory (rfe_alloc_pinned) for all output arrays, or the copies serialise with the host. */
int rfe_pairs_collect_begin_full(rfe_ctx* ctx, float match_thresh, int32_t* kpts_xy, float* scores, float* desc,
                                 int32_t* matches, float* mscores, int cap);
/* Page-locked host memory for the input / output arrays of the pipelined calls (cudaMallocHost / cudaFreeHost). */
void* rfe_alloc_pinned(size_t bytes);
void rfe_free_pinned(void* p);
/* Device-to-device copy (asynchronous, ctx stream) of the first n_pairs match results -- matches [n_pairs][max_keypoints][2],
 * mscores [n_pairs][max_keypoints] (may be NULL), counts [n_pairs] -- into caller-owned device buffers: the hand-off to a
 * result gather across GPUs (bench.py, BASELINE config 5) without a host round trip. */
int rfe_lg_copy_results_device(rfe_ctx* ctx, int n_pairs, int32_t* d_matches, float* d_mscores, int32_t* d_counts);
/* Copy a match result to the host (synchronises). */
int rfe_lg_read_result(rfe_ctx* ctx, int rslot, int32_t* matches, float* mscores, int* k, int cap);

/* ---- timers / introspection --------------------------------------------------------------------- */
/* Accumulated GPU milliseconds (CUDA events) of "extractor" / "matcher" calls, like the reference's GetTimer. */
double rfe_get_timer_ms(rfe_ctx* ctx, const char* name);
/* Bytes the pipelined path (rfe_pairs_submit / rfe_pairs_collect, rfe_match_pairs_u8) has copied host->device and
 * device->host so far. */
int rfe_transfer_bytes(rfe_ctx* ctx, unsigned long long* h2d, unsigned long long* d2h);
/* Number of kernels the library has launched on this ctx so far. */
long long rfe_kernel_launches(rfe_ctx* ctx);
/* Per-kernel timing with CUDA events on the ctx stream (off by default).  Tags: "sp.conv1b", "sp.conv2a", ...,
 * "sp.convPb_softmax", "sp.convDb_l2norm", "lg.wqkv", "lg.attn_qk", "lg.attn_pv", "lg.ffn0", "lg.sim", ...
 * rfe_profile_read sums the launches whose tag starts with `prefix` (NULL = all); synchronises. */
int rfe_profile(rfe_ctx* ctx, int enable);
/* Restrict the recording to launches whose tag starts with `prefix` (NULL or "" = every launch): bench.py times only the
 * dominant kernel inside its timed region so that the event records do not perturb the throughput number. */
int rfe_profile_select(rfe_ctx* ctx, const char* prefix);
int rfe_profile_read(rfe_ctx* ctx, const char* prefix, double* total_ms, long long* launches, int reset);
/* Test hook: copy a named intermediate device tensor of the last call to the host (fp32 or raw bytes).
 * Returns the byte size of the tensor through *bytes; copies min(*bytes, capacity). */
int rfe_debug_read(rfe_ctx* ctx, const char* name, void* dst, size_t capacity, size_t* bytes);
/* Test hook: run one split-fp16 tensor-core GEMM D = A[M,K] * B[N,K]^T (+bias) on host fp32 data. */
int rfe_debug_gemm(rfe_ctx* ctx, const float* a, const float* b, const float* bias, float* d, int m, int n, int kdim);

#ifdef RFE_ENABLE_PROBES
/* Measurement scaffolding, only in a `make PROBES=1` build: tcgen05 hardware probes (csrc/probe_kernels.cu). */
int rfe_debug_probe(rfe_ctx* ctx, int which, const float* a, const float* b, float* out);
#endif

#ifdef __cplusplus
}
#endif
#endif /* ROVER_FE_H_ */
