// Launch wrappers of the non-tensor-core kernels (sp_kernels.cu, lg_kernels.cu).
#pragma once

#include "common.cuh"

namespace rfe {

// ---- SuperPoint ------------------------------------------------------------------------------------------
void launch_conv1a(cudaStream_t s, const uint8_t* img, int stride, int H, int W, int B, const float* w,
                   const float* bias, __half* out_hi, __half* out_lo);
int nms_prepare();
void launch_nms(cudaStream_t s, const float* heat, float* out, int B, int H, int W);
int nms2_prepare();
// bit-plane NMS that also produces the per-row keypoint counts of the ordered compaction
void launch_nms2(cudaStream_t s, const float* heat, float* out, int* row_cnt, int B, int H, int W, float thr);
void launch_select(cudaStream_t s, const float* nms, int B, int H, int W, float thr, int cap, int* row_cnt,
                   int* row_off, int* counts, int* kpts, float* scores, bool have_counts);
// optional top-K cap on the compacted keypoint lists (K <= 0: off = the reference's behaviour)
void launch_topk(cudaStream_t s, int B, int cap, int K, int* counts, int* kpts, float* scores);
// rowss != null: `dense` is un-normalised and rowss [B*h*w][4] holds four partial sums of squares per pixel
void launch_desc_sample(cudaStream_t s, const float* dense, const float* rowss, int h, int w, int B, const int* kpts,
                        const int* counts, int cap, float* desc, uint8_t* desc_bin);
void launch_dense_normalize(cudaStream_t s, const float* dense, const float* rowss, size_t npix, float* out);
void launch_binarize(cudaStream_t s, const float* desc, int n, uint8_t* out, uint32_t* bits);
void launch_l2_best2(cudaStream_t s, const float* q, int nq, const float* db, const int* cand_off, const int* cand_idx,
                     float init_dist, float* best_dist, int* best_idx, float* second_dist, int* second_idx);

// ---- LightGlue -------------------------------------------------------------------------------------------
constexpr int kLgMaxImages = 32;          // images (2 per pair) handled by one batched launch
struct LgImages {                         // inputs of one batched LightGlue pass, one entry per image
  const float* kpts_f[kLgMaxImages];      // pixel keypoints fp32 [n][2] (host API) or null
  const int* kpts_i[kLgMaxImages];        // pixel keypoints int32 [n][2] (device hand-off from SuperPoint) or null
  const float* desc[kLgMaxImages];        // descriptors fp32 [n][256]
  int n[kLgMaxImages];                    // keypoints
  int row0[kLgMaxImages];                 // first row of the image in the concatenated LightGlue state
  int count;
};
struct LgAssign {                         // the assignment stage of all pairs of a batch (blockIdx.y = pair)
  const float* sim[kLgMaxImages / 2];     // [n0][ld] similarity of pair p
  int n0[kLgMaxImages / 2], n1[kLgMaxImages / 2], ld[kLgMaxImages / 2];
  int off0[kLgMaxImages / 2], off1[kLgMaxImages / 2];   // row offsets of the two images (index the per-row vectors)
  int* matches[kLgMaxImages / 2];         // [cap][2] result slot of pair p
  float* mscores[kLgMaxImages / 2];
  int* count[kLgMaxImages / 2];
  int pairs, max_n0, max_n1;
};
struct LgCacheMove {                      // batch state <-> per-slot layer-0 cache, one entry per image
  float* x; __half* cat_hi; __half* cat_lo; float* cs; float* sn;            // batch state ([rows][256] / [rows][512] / [rows][32])
  float* cx; __half* ccat_hi; __half* ccat_lo; float* ccs; float* csn;       // cache ([slot][cap][256] / [256] / [32])
  int slot[kLgMaxImages], n[kLgMaxImages], row0[kLgMaxImages];
  int count, cap, to_cache;
};
// to_cache = 1: store rows [row0, row0 + n) of every image into its slot's cache entry; 0: load them back (and zero the
// image's padding rows, as lg_prepare does)
void launch_lg_cache_move(cudaStream_t s, const LgCacheMove& mv, int max_n);
// positional encoding + residual-stream initialisation (x = desc, cat[:, :256] = split(desc)) for every image
void launch_lg_prepare(cudaStream_t s, const LgImages& im, int max_n, int norm_h, int norm_w, const float* wr, float* cs,
                       float* sn, float* x, __half* cat_hi, __half* cat_lo);
// dual log-softmax statistics, both arg-maxes and the mutual-match compaction for all pairs
void launch_lg_assign(cudaStream_t s, const LgAssign& a, float* rmax, float* rlog, float* cmax, float* clog,
                      const float* ls, float* max0, int* m0, int* m1, float filter, float thresh, float* S_dbg);
// the same in two passes over sim (32-row bands; part_a / part_b: [pairs][part_bands][part_ld] scratch)
void launch_lg_assign_banded(cudaStream_t s, const LgAssign& a, float* rmax, float* rlog, float* cmax, float* clog,
                             const float* ls, float* max0, int* m0, int* m1, float filter, float thresh, float* S_dbg,
                             float* part_a, float* part_b, int part_ld, int part_bands);
void launch_posenc(cudaStream_t s, const float* kpts_px, int n, int norm_h, int norm_w, const float* wr, float* cs,
                   float* sn);
void launch_kpts_to_float(cudaStream_t s, const int* k, int n, float* o);
void launch_split_rows(cudaStream_t s, const float* src, int rows, int cols, int ld_src, float* dst, int ld_dst,
                       __half* hi, __half* lo, int ld_h);
void launch_ln_gelu_split(cudaStream_t s, const float* x, int rows, const float* g, const float* b, __half* hi,
                          __half* lo);
void launch_matchability(cudaStream_t s, const float* x, int rows, const float* w, const float* b, float* out);
void launch_lse(cudaStream_t s, const float* sim, int n0, int n1, int ld, float* rmax, float* rlog, float* cmax,
                float* clog);
void launch_argmax(cudaStream_t s, const float* sim, int n0, int n1, int ld, const float* rmax, const float* rlog,
                   const float* cmax, const float* clog, const float* ls0, const float* ls1, float* max0, int* m0,
                   int* m1, float* S_dbg);
void launch_match_compact(cudaStream_t s, const float* max0, const int* m0, const int* m1, int n0, float filter,
                          float thresh, int* matches, float* mscores, int* count);

}  // namespace rfe
