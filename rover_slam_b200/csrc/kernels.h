// Launch wrappers of the non-tensor-core kernels (sp_kernels.cu, lg_kernels.cu).
#pragma once

#include "common.cuh"

namespace rfe {

// ---- SuperPoint ------------------------------------------------------------------------------------------
void launch_conv1a(cudaStream_t s, const uint8_t* img, int stride, int H, int W, int B, const float* w,
                   const float* bias, __half* out_hi, __half* out_lo);
int nms_prepare();
void launch_nms(cudaStream_t s, const float* heat, float* out, int B, int H, int W);
void launch_select(cudaStream_t s, const float* nms, int B, int H, int W, float thr, int cap, int* row_cnt,
                   int* row_off, int* counts, int* kpts, float* scores);
void launch_desc_sample(cudaStream_t s, const float* dense, int h, int w, int B, const int* kpts, const int* counts,
                        int cap, float* desc);

// ---- LightGlue -------------------------------------------------------------------------------------------
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
