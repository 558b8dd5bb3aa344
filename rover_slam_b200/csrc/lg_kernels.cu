// LightGlue kernels that are not tensor-core contractions: keypoint normalisation + Fourier positional
// encoding, LayerNorm+GELU, matchability, dual log-softmax
// statistics, row/column arg-max, mutual check + ordered compaction.
// Reference: onnxmodel/lightglue_sim.onnx as run by src/Matchers/lightglue_onnx.cpp:210-214 (SURVEY.md
// Appendix B), keypoint normalisation from src/Matchers/transform.cpp:19-32.
#include "kernels.h"

namespace rfe {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- posenc: cs/sn [N][32] (nodes 0-17) ---------------------------------------------------------------
__global__ void posenc_kernel(const float* __restrict__ kpts_px, int n, float shift_x, float shift_y, float scale,
                              const float* __restrict__ wr /*[32][2]*/, float* __restrict__ cs,
                              float* __restrict__ sn) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 32) return;
  const int k = i >> 5, j = i & 31;
  const float x = (kpts_px[2 * k] - shift_x) / scale;
  const float y = (kpts_px[2 * k + 1] - shift_y) / scale;
  const float p = x * wr[2 * j] + y * wr[2 * j + 1];
  cs[i] = cosf(p);
  sn[i] = sinf(p);
}
void launch_posenc(cudaStream_t s, const float* kpts_px, int n, int norm_h, int norm_w, const float* wr, float* cs,
                   float* sn) {
  if (n == 0) return;
  const float sx = static_cast<float>(norm_w) / 2.0f, sy = static_cast<float>(norm_h) / 2.0f;
  const float sc = static_cast<float>(norm_w > norm_h ? norm_w : norm_h) / 2.0f;
  posenc_kernel<<<(n * 32 + 255) / 256, 256, 0, s>>>(kpts_px, n, sx, sy, sc, wr, cs, sn);
}

// ---- int keypoints -> float pixels (device-resident hand-off from SuperPoint) ---------------------------
__global__ void kpts_to_float_kernel(const int* __restrict__ k, int n, float* __restrict__ o) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * n) o[i] = static_cast<float>(k[i]);
}
void launch_kpts_to_float(cudaStream_t s, const int* k, int n, float* o) {
  if (n) kpts_to_float_kernel<<<(2 * n + 255) / 256, 256, 0, s>>>(k, n, o);
}

// ---- fp32 [rows][cols] -> fp32 copy + split-fp16 (row pitches may differ) -------------------------------
__global__ void split_rows_kernel(const float* __restrict__ src, int rows, int cols, int ld_src, float* __restrict__ dst,
                                  int ld_dst, __half* __restrict__ hi, __half* __restrict__ lo, int ld_h) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(rows) * cols) return;
  const int r = i / cols, c = i - static_cast<size_t>(r) * cols;
  const float v = src[static_cast<size_t>(r) * ld_src + c];
  if (dst) dst[static_cast<size_t>(r) * ld_dst + c] = v;
  __half h, l;
  split_f32(v, h, l);
  hi[static_cast<size_t>(r) * ld_h + c] = h;
  lo[static_cast<size_t>(r) * ld_h + c] = l;
}
void launch_split_rows(cudaStream_t s, const float* src, int rows, int cols, int ld_src, float* dst, int ld_dst,
                       __half* hi, __half* lo, int ld_h) {
  const size_t n = static_cast<size_t>(rows) * cols;
  if (n == 0) return;
  split_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(src, rows, cols, ld_src, dst, ld_dst, hi, lo,
                                                                          ld_h);
}

// GELU(y) = 0.5 y (1 + erf(y / sqrt 2))  (lightglue_sim.onnx nodes Div / Erf / Add / Mul / Mul of every FFN).
// erff() compiles to ~52 instructions with two divergent branches and bounded this kernel at 2x its HBM floor.  Here
// erfc(|x|) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-x^2), t = 1 / (1 + p |x|)  (Abramowitz & Stegun 7.1.26,
// |error| <= 1.5e-7), branch free: for y < 0, 1 + erf = erfc(|x|) is used directly, so the negative tail keeps its relative
// accuracy.  Measured against float64 over [-12, 12] and N(0, 2): max abs GELU error 4.2e-7, the same envelope as the
// erff() formulation evaluated in fp32 (4.4e-7); tests/test_split_numerics.py restates and checks it.
__device__ __forceinline__ float gelu_erf(float y) {
  const float x = fabsf(y) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, x, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -1.4426950408889634f));
  const float c = p * e;                              // erfc(|x|)
  return 0.5f * y * (y < 0.0f ? c : 2.0f - c);
}

// ---- LayerNorm(512, eps 1e-5) + GELU -> split-fp16 (one warp per row; nodes 62-67) --------------------
// Lane l owns columns 128 j + 4 l .. + 3 (j = 0..3): float4 loads, 8-byte stores per plane, packed split.
__global__ void __launch_bounds__(256) ln_gelu_split_kernel(const float* __restrict__ x, int rows,
                                                            const float* __restrict__ g, const float* __restrict__ b,
                                                            __half* __restrict__ hi, __half* __restrict__ lo) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= rows) return;
  const float4* r4 = reinterpret_cast<const float4*>(x + static_cast<size_t>(wid) * 512);
  float v[16];
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 t = r4[j * 32 + lane];
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    s += (t.x + t.y) + (t.z + t.w);
  }
  const float mean = warp_sum(s) * (1.0f / 512.0f);
  float q = 0.0f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float d = v[j] - mean;
    q += d * d;
  }
  const float var = warp_sum(q) * (1.0f / 512.0f);
  const float rstd = 1.0f / sqrtf(var + 1e-5f);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = j * 128 + lane * 4;
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g + c)), b4 = __ldg(reinterpret_cast<const float4*>(b + c));
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
    float ge[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float y = (v[4 * j + i] - mean) * rstd * gg[i] + bb[i];
      ge[i] = gelu_erf(y);
    }
    uint32_t h0, l0, h1, l1;
    split2(pk2(ge[0], ge[1]), h0, l0);
    split2(pk2(ge[2], ge[3]), h1, l1);
    *reinterpret_cast<uint2*>(hi + static_cast<size_t>(wid) * 512 + c) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(lo + static_cast<size_t>(wid) * 512 + c) = make_uint2(l0, l1);
  }
}
void launch_ln_gelu_split(cudaStream_t s, const float* x, int rows, const float* g, const float* b, __half* hi,
                          __half* lo) {
  if (rows == 0) return;
  ln_gelu_split_kernel<<<(rows * 32 + 255) / 256, 256, 0, s>>>(x, rows, g, b, hi, lo);
}

// ---- matchability: log(sigmoid(x . w + b)) (nodes 1488-1495), one warp per row --------------------------------
__global__ void __launch_bounds__(256) matchability_kernel(const float* __restrict__ x, int rows,
                                                           const float* __restrict__ w, const float* __restrict__ b,
                                                           float* __restrict__ out) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= rows) return;
  const float* r = x + static_cast<size_t>(wid) * 256;
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s = fmaf(r[j * 32 + lane], w[j * 32 + lane], s);
  s = warp_sum(s);
  if (lane == 0) {
    const float z = b[0] + s;
    out[wid] = logf(1.0f / (1.0f + expf(-z)));
  }
}
void launch_matchability(cudaStream_t s, const float* x, int rows, const float* w, const float* b, float* out) {
  if (rows == 0) return;
  matchability_kernel<<<(rows * 32 + 255) / 256, 256, 0, s>>>(x, rows, w, b, out);
}

// ---- dual log-softmax statistics ------------------------------------------------------------------------------
// row_stat[i] = (max_j sim[i][j], log sum_j exp(sim - max))   one warp per row
__global__ void __launch_bounds__(256) row_lse_kernel(const float* __restrict__ sim, int n0, int n1, int ld,
                                                      float* __restrict__ rmax, float* __restrict__ rlog) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= n0) return;
  const float* r = sim + static_cast<size_t>(wid) * ld;
  float mx = -INFINITY;
  for (int j = lane; j < n1; j += 32) mx = fmaxf(mx, r[j]);
  mx = warp_max(mx);
  float sum = 0.0f;
  for (int j = lane; j < n1; j += 32) sum += expf(r[j] - mx);
  sum = warp_sum(sum);
  if (lane == 0) {
    rmax[wid] = mx;
    rlog[wid] = logf(sum);
  }
}
// column statistics: block = 32 columns x 32 row-lanes; two passes over the rows (max, then sum)
__global__ void __launch_bounds__(1024) col_lse_kernel(const float* __restrict__ sim, int n0, int n1, int ld,
                                                       float* __restrict__ cmax, float* __restrict__ clog) {
  __shared__ float red[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  float mx = -INFINITY;
  if (j < n1)
    for (int i = ty; i < n0; i += 32) mx = fmaxf(mx, sim[static_cast<size_t>(i) * ld + j]);
  red[ty][tx] = mx;
  __syncthreads();
  if (ty == 0) {
    float m = red[0][tx];
    for (int k = 1; k < 32; ++k) m = fmaxf(m, red[k][tx]);
    red[0][tx] = m;
  }
  __syncthreads();
  mx = red[0][tx];
  __syncthreads();
  float sum = 0.0f;
  if (j < n1)
    for (int i = ty; i < n0; i += 32) sum += expf(sim[static_cast<size_t>(i) * ld + j] - mx);
  red[ty][tx] = sum;
  __syncthreads();
  if (ty == 0 && j < n1) {
    float sacc = 0.0f;
    for (int k = 0; k < 32; ++k) sacc += red[k][tx];
    cmax[j] = mx;
    clog[j] = logf(sacc);
  }
}
void launch_lse(cudaStream_t s, const float* sim, int n0, int n1, int ld, float* rmax, float* rlog, float* cmax,
                float* clog) {
  if (n0 == 0 || n1 == 0) return;
  row_lse_kernel<<<(n0 * 32 + 255) / 256, 256, 0, s>>>(sim, n0, n1, ld, rmax, rlog);
  col_lse_kernel<<<(n1 + 31) / 32, 1024, 0, s>>>(sim, n0, n1, ld, cmax, clog);
}

// S[i][j] = ((sim - rmax_i) - rlog_i) + ((sim - cmax_j) - clog_j) + (ls0_i + ls1_j)         (nodes 1497-1501)
__device__ __forceinline__ float assign_score(float v, float rm, float rl, float cm, float cl, float a0, float a1) {
  return (((v - rm) - rl) + ((v - cm) - cl)) + (a0 + a1);
}

// row arg-max (TopK k=1 axis 2; ties -> lowest index), one warp per row
__global__ void __launch_bounds__(256) row_argmax_kernel(const float* __restrict__ sim, int n0, int n1, int ld,
                                                         const float* __restrict__ rmax, const float* __restrict__ rlog,
                                                         const float* __restrict__ cmax, const float* __restrict__ clog,
                                                         const float* __restrict__ ls0, const float* __restrict__ ls1,
                                                         float* __restrict__ max0, int* __restrict__ m0,
                                                         float* __restrict__ S_dbg) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= n0) return;
  const float* r = sim + static_cast<size_t>(wid) * ld;
  const float rm = rmax[wid], rl = rlog[wid], a0 = ls0[wid];
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = lane; j < n1; j += 32) {
    const float sc = assign_score(r[j], rm, rl, cmax[j], clog[j], a0, ls1[j]);
    if (S_dbg) S_dbg[static_cast<size_t>(wid) * n1 + j] = sc;
    if (sc > best) {
      best = sc;
      bi = j;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) {
      best = ob;
      bi = oi;
    }
  }
  if (lane == 0) {
    max0[wid] = best;
    m0[wid] = bi;
  }
}
// column arg-max (TopK k=1 axis 1), block = 32 columns x 32 row-lanes
__global__ void __launch_bounds__(1024) col_argmax_kernel(const float* __restrict__ sim, int n0, int n1, int ld,
                                                          const float* __restrict__ rmax, const float* __restrict__ rlog,
                                                          const float* __restrict__ cmax, const float* __restrict__ clog,
                                                          const float* __restrict__ ls0, const float* __restrict__ ls1,
                                                          int* __restrict__ m1) {
  __shared__ float rb[32][33];
  __shared__ int ri[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  if (j < n1) {
    const float cm = cmax[j], cl = clog[j], a1 = ls1[j];
    for (int i = ty; i < n0; i += 32) {
      const float sc = assign_score(sim[static_cast<size_t>(i) * ld + j], rmax[i], rlog[i], cm, cl, ls0[i], a1);
      if (sc > best) {
        best = sc;
        bi = i;
      }
    }
  }
  rb[ty][tx] = best;
  ri[ty][tx] = bi;
  __syncthreads();
  if (ty == 0 && j < n1) {
    for (int k = 1; k < 32; ++k) {
      const float ob = rb[k][tx];
      const int oi = ri[k][tx];
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    m1[j] = bi;
  }
}
void launch_argmax(cudaStream_t s, const float* sim, int n0, int n1, int ld, const float* rmax, const float* rlog,
                   const float* cmax, const float* clog, const float* ls0, const float* ls1, float* max0, int* m0,
                   int* m1, float* S_dbg) {
  if (n0 == 0 || n1 == 0) return;
  row_argmax_kernel<<<(n0 * 32 + 255) / 256, 256, 0, s>>>(sim, n0, n1, ld, rmax, rlog, cmax, clog, ls0, ls1, max0, m0,
                                                           S_dbg);
  col_argmax_kernel<<<(n1 + 31) / 32, 1024, 0, s>>>(sim, n0, n1, ld, rmax, rlog, cmax, clog, ls0, ls1, m1);
}

// ---- mutual check + exp + filter + ordered compaction (nodes 1504-1525 + lightglue_onnx.cpp:437-453) -----------
__global__ void __launch_bounds__(1024) match_compact_kernel(const float* __restrict__ max0, const int* __restrict__ m0,
                                                             const int* __restrict__ m1, int n0, float filter,
                                                             float thresh, int* __restrict__ matches,
                                                             float* __restrict__ mscores, int* __restrict__ count) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n0; base += 1024) {
    const int i = base + threadIdx.x;
    float ms = 0.0f;
    int j = 0;
    if (i < n0) {
      j = m0[i];
      const bool mutual = (m1[j] == i);
      ms = mutual ? expf(max0[i]) : 0.0f;
    }
    const int keep = (i < n0 && ms > filter && ms > thresh) ? 1 : 0;
    int incl = keep;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
      int ws = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, ws, o);
        if (lane >= o) ws += t;
      }
      warp_sums[lane] = ws;
    }
    __syncthreads();
    const int pos = carry + (w ? warp_sums[w - 1] : 0) + incl - keep;
    if (keep) {
      matches[2 * pos] = i;
      matches[2 * pos + 1] = j;
      mscores[pos] = ms;
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry = pos + keep;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = carry;
}
void launch_match_compact(cudaStream_t s, const float* max0, const int* m0, const int* m1, int n0, float filter,
                          float thresh, int* matches, float* mscores, int* count) {
  match_compact_kernel<<<1, 1024, 0, s>>>(max0, m0, m1, n0, filter, thresh, matches, mscores, count);
}

// ================================================================================================
// Batched forms: one launch covers every image / every pair of a LightGlue batch
// ================================================================================================
// lg_prepare: one warp per keypoint row.  Keypoint normalisation (transform.cpp:19-32) + Fourier positional encoding
// (nodes 0-17) + x = desc, cat[:, 0:256] = split(desc).  grid = (ceil(max_n / 8), images).
__global__ void __launch_bounds__(256) lg_prepare_kernel(const LgImages im, float shift_x, float shift_y, float scale,
                                                         const float* __restrict__ wr, float* __restrict__ cs,
                                                         float* __restrict__ sn, float* __restrict__ x,
                                                         __half* __restrict__ cat_hi, __half* __restrict__ cat_lo) {
  const int img = blockIdx.y;
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int n = im.n[img];
  if (k >= ((n + 7) & ~7)) return;
  if (k >= n) {
    // Padding rows (every image starts at a multiple of 8 rows): every GEMM and FFN runs over them, and the last key tile
    // of the attention kernel loads their K / V^T columns (with P = 0).  Left uninitialised they would carry their state
    // from call to call through 18 more residual updates each time and eventually overflow the fp16 planes
    // (0 * inf = NaN for every query of the image): reset them to zero on every call.
    const size_t row = static_cast<size_t>(im.row0[img]) + k;
    cs[row * 32 + lane] = 0.0f;
    sn[row * 32 + lane] = 0.0f;
    const float4 z4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float4* x4 = reinterpret_cast<float4*>(x + row * 256);
    x4[lane] = z4;
    x4[32 + lane] = z4;
    *reinterpret_cast<uint2*>(cat_hi + row * 512 + 4 * lane) = make_uint2(0u, 0u);
    *reinterpret_cast<uint2*>(cat_lo + row * 512 + 4 * lane) = make_uint2(0u, 0u);
    *reinterpret_cast<uint2*>(cat_hi + row * 512 + 128 + 4 * lane) = make_uint2(0u, 0u);
    *reinterpret_cast<uint2*>(cat_lo + row * 512 + 128 + 4 * lane) = make_uint2(0u, 0u);
    return;
  }
  float px, py;
  if (im.kpts_i[img]) {
    px = static_cast<float>(im.kpts_i[img][2 * k]);
    py = static_cast<float>(im.kpts_i[img][2 * k + 1]);
  } else {
    px = im.kpts_f[img][2 * k];
    py = im.kpts_f[img][2 * k + 1];
  }
  const size_t row = static_cast<size_t>(im.row0[img]) + k;
  const float xn = (px - shift_x) / scale;
  const float yn = (py - shift_y) / scale;
  const float ph = xn * wr[2 * lane] + yn * wr[2 * lane + 1];
  cs[row * 32 + lane] = cosf(ph);
  sn[row * 32 + lane] = sinf(ph);
  const float4* d4 = reinterpret_cast<const float4*>(im.desc[img] + static_cast<size_t>(k) * 256);
  const float4 a = d4[lane], b = d4[32 + lane];          // columns 4*lane.., 128 + 4*lane..
  float4* x4 = reinterpret_cast<float4*>(x + row * 256);
  x4[lane] = a;
  x4[32 + lane] = b;
  uint32_t h0, l0, h1, l1;
  split2(pk2(a.x, a.y), h0, l0);
  split2(pk2(a.z, a.w), h1, l1);
  *reinterpret_cast<uint2*>(cat_hi + row * 512 + 4 * lane) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(cat_lo + row * 512 + 4 * lane) = make_uint2(l0, l1);
  split2(pk2(b.x, b.y), h0, l0);
  split2(pk2(b.z, b.w), h1, l1);
  *reinterpret_cast<uint2*>(cat_hi + row * 512 + 128 + 4 * lane) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(cat_lo + row * 512 + 128 + 4 * lane) = make_uint2(l0, l1);
}
void launch_lg_prepare(cudaStream_t s, const LgImages& im, int max_n, int norm_h, int norm_w, const float* wr, float* cs,
                       float* sn, float* x, __half* cat_hi, __half* cat_lo) {
  if (im.count == 0 || max_n == 0) return;
  // norm_h = norm_w = 0: the keypoints are ALREADY normalised by the caller (rfe_lg_match_normalized): (k - 0) / 1 is exact
  const bool ident = norm_h <= 0 || norm_w <= 0;
  const float sx = ident ? 0.0f : static_cast<float>(norm_w) / 2.0f, sy = ident ? 0.0f : static_cast<float>(norm_h) / 2.0f;
  const float sc = ident ? 1.0f : static_cast<float>(norm_w > norm_h ? norm_w : norm_h) / 2.0f;
  lg_prepare_kernel<<<dim3((max_n + 7) / 8, im.count), 256, 0, s>>>(im, sx, sy, sc, wr, cs, sn, x, cat_hi, cat_lo);
}

// Per-slot layer-0 cache (rover_fe.cu, lg_build_cache): one warp per keypoint row copies x (1 KB), the split copy of x
// (2 x 512 B) and cos / sin (2 x 128 B) between the batch state and the slot's cache entry.  grid = (ceil(max_n / 8), images).
__global__ void __launch_bounds__(256) lg_cache_move_kernel(const LgCacheMove mv) {
  const int img = blockIdx.y;
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int n = mv.n[img];
  if (k >= ((n + 7) & ~7)) return;
  const size_t row = static_cast<size_t>(mv.row0[img]) + k;
  const size_t crow = static_cast<size_t>(mv.slot[img]) * mv.cap + k;
  float4* bx = reinterpret_cast<float4*>(mv.x + row * 256);
  uint4* bh = reinterpret_cast<uint4*>(mv.cat_hi + row * 512);
  uint4* bl = reinterpret_cast<uint4*>(mv.cat_lo + row * 512);
  if (k >= n) {                          // padding rows of the batch state: zero, like lg_prepare (never stored)
    if (!mv.to_cache) {
      const float4 z4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      bx[lane] = z4; bx[32 + lane] = z4;
      bh[lane] = make_uint4(0u, 0u, 0u, 0u);
      bl[lane] = make_uint4(0u, 0u, 0u, 0u);
      mv.cs[row * 32 + lane] = 0.0f;
      mv.sn[row * 32 + lane] = 0.0f;
    }
    return;
  }
  float4* cx = reinterpret_cast<float4*>(mv.cx + crow * 256);
  uint4* ch = reinterpret_cast<uint4*>(mv.ccat_hi + crow * 256);
  uint4* cl = reinterpret_cast<uint4*>(mv.ccat_lo + crow * 256);
  if (mv.to_cache) {
    cx[lane] = bx[lane]; cx[32 + lane] = bx[32 + lane];
    ch[lane] = bh[lane];
    cl[lane] = bl[lane];
    mv.ccs[crow * 32 + lane] = mv.cs[row * 32 + lane];
    mv.csn[crow * 32 + lane] = mv.sn[row * 32 + lane];
  } else {
    bx[lane] = cx[lane]; bx[32 + lane] = cx[32 + lane];
    bh[lane] = ch[lane];
    bl[lane] = cl[lane];
    mv.cs[row * 32 + lane] = mv.ccs[crow * 32 + lane];
    mv.sn[row * 32 + lane] = mv.csn[crow * 32 + lane];
  }
}
void launch_lg_cache_move(cudaStream_t s, const LgCacheMove& mv, int max_n) {
  if (mv.count == 0 || max_n == 0) return;
  lg_cache_move_kernel<<<dim3((max_n + 7) / 8, mv.count), 256, 0, s>>>(mv);
}

// Row statistics / row arg-max: one warp per row, blockIdx.y = pair.  Column statistics / arg-max: block = 32 columns x
// 32 row-lanes.  Per-row vectors are indexed by the row's position in the concatenated LightGlue state.
__global__ void __launch_bounds__(256) row_lse_batch_kernel(const LgAssign a, float* __restrict__ rmax,
                                                            float* __restrict__ rlog) {
  const int pr = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int n0 = a.n0[pr], n1 = a.n1[pr];
  if (i >= n0) return;
  const float* r = a.sim[pr] + static_cast<size_t>(i) * a.ld[pr];
  float mx = -INFINITY;
#pragma unroll 8
  for (int j = lane; j < n1; j += 32) mx = fmaxf(mx, r[j]);
  mx = warp_max(mx);
  float sum = 0.0f;
#pragma unroll 8
  for (int j = lane; j < n1; j += 32) sum += expf(r[j] - mx);
  sum = warp_sum(sum);
  if (lane == 0) {
    rmax[a.off0[pr] + i] = mx;
    rlog[a.off0[pr] + i] = logf(sum);
  }
}
__global__ void __launch_bounds__(1024) col_lse_batch_kernel(const LgAssign a, float* __restrict__ cmax,
                                                             float* __restrict__ clog) {
  __shared__ float red[32][33];
  const int pr = blockIdx.y;
  const int n0 = a.n0[pr], n1 = a.n1[pr], ld = a.ld[pr];
  if (blockIdx.x * 32 >= n1) return;
  const float* sim = a.sim[pr];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  float mx = -INFINITY;
  if (j < n1) {
#pragma unroll 8
    for (int i = ty; i < n0; i += 32) mx = fmaxf(mx, sim[static_cast<size_t>(i) * ld + j]);
  }
  red[ty][tx] = mx;
  __syncthreads();
  if (ty == 0) {
    float m = red[0][tx];
    for (int k = 1; k < 32; ++k) m = fmaxf(m, red[k][tx]);
    red[0][tx] = m;
  }
  __syncthreads();
  mx = red[0][tx];
  __syncthreads();
  float sum = 0.0f;
  if (j < n1) {
#pragma unroll 8
    for (int i = ty; i < n0; i += 32) sum += expf(sim[static_cast<size_t>(i) * ld + j] - mx);
  }
  red[ty][tx] = sum;
  __syncthreads();
  if (ty == 0 && j < n1) {
    float sacc = 0.0f;
    for (int k = 0; k < 32; ++k) sacc += red[k][tx];
    cmax[a.off1[pr] + j] = mx;
    clog[a.off1[pr] + j] = logf(sacc);
  }
}
// Both arg-max kernels read sim as float4 (four columns per thread): with scalar loads they ran at 1.4-1.7 TB/s, bound by the
// number of loads in flight (profiles/r02_step_tail_full.csv).  A thread visits its candidates in ascending index order and
// replaces only on a strictly larger score, the cross-thread reductions prefer the lower index on equality: TopK's
// lowest-index tie rule, as before.
__global__ void __launch_bounds__(256) row_argmax_batch_kernel(const LgAssign a, const float* __restrict__ rmax,
                                                               const float* __restrict__ rlog,
                                                               const float* __restrict__ cmax,
                                                               const float* __restrict__ clog, const float* __restrict__ ls,
                                                               float* __restrict__ max0, int* __restrict__ m0,
                                                               float* __restrict__ S_dbg) {
  const int pr = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int n0 = a.n0[pr], n1 = a.n1[pr];
  if (i >= n0) return;
  const float* r = a.sim[pr] + static_cast<size_t>(i) * a.ld[pr];
  const int o0 = a.off0[pr], o1 = a.off1[pr];
  const float rm = rmax[o0 + i], rl = rlog[o0 + i], a0 = ls[o0 + i];
  // rows of sim and the per-column vectors start 32-byte aligned (ld and off1 are multiples of 8) and are padded to 8 columns
  const float4* r4 = reinterpret_cast<const float4*>(r);
  const float4* cm4 = reinterpret_cast<const float4*>(cmax + o1);
  const float4* cl4 = reinterpret_cast<const float4*>(clog + o1);
  const float4* ls4 = reinterpret_cast<const float4*>(ls + o1);
  float* dbg = (S_dbg && pr == a.pairs - 1) ? S_dbg + static_cast<size_t>(i) * n1 : nullptr;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  const int nv = (n1 + 3) >> 2;
#pragma unroll 4
  for (int v = lane; v < nv; v += 32) {
    const float4 sv = r4[v], cm = cm4[v], cl = cl4[v], l1 = ls4[v];
    const int j = 4 * v;
    const float sc[4] = {assign_score(sv.x, rm, rl, cm.x, cl.x, a0, l1.x), assign_score(sv.y, rm, rl, cm.y, cl.y, a0, l1.y),
                         assign_score(sv.z, rm, rl, cm.z, cl.z, a0, l1.z), assign_score(sv.w, rm, rl, cm.w, cl.w, a0, l1.w)};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (j + e < n1) {
        if (dbg) dbg[j + e] = sc[e];
        if (sc[e] > best) {
          best = sc[e];
          bi = j + e;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) {
      best = ob;
      bi = oi;
    }
  }
  if (lane == 0) {
    max0[o0 + i] = best;
    m0[o0 + i] = bi;
  }
}
// block = 64 columns (16 threads x float4) x 64 row-lanes
__global__ void __launch_bounds__(1024) col_argmax_batch_kernel(const LgAssign a, const float* __restrict__ rmax,
                                                                const float* __restrict__ rlog,
                                                                const float* __restrict__ cmax,
                                                                const float* __restrict__ clog, const float* __restrict__ ls,
                                                                int* __restrict__ m1) {
  __shared__ float rb[64][65];
  __shared__ int ri[64][65];
  const int pr = blockIdx.y;
  const int n0 = a.n0[pr], n1 = a.n1[pr], ld = a.ld[pr];
  if (blockIdx.x * 64 >= n1) return;
  const float* sim = a.sim[pr];
  const int o0 = a.off0[pr], o1 = a.off1[pr];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int j = blockIdx.x * 64 + 4 * tx;
  float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int bi[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
  if (j < n1) {                                       // j + 3 < ld: columns beyond n1 are computed and never stored
    const float4 cm = *reinterpret_cast<const float4*>(cmax + o1 + j), cl = *reinterpret_cast<const float4*>(clog + o1 + j),
                 a1 = *reinterpret_cast<const float4*>(ls + o1 + j);
#pragma unroll 4
    for (int i = ty; i < n0; i += 64) {
      const float4 sv = *reinterpret_cast<const float4*>(sim + static_cast<size_t>(i) * ld + j);
      const float rm = rmax[o0 + i], rl = rlog[o0 + i], l0 = ls[o0 + i];
      const float sc[4] = {assign_score(sv.x, rm, rl, cm.x, cl.x, l0, a1.x), assign_score(sv.y, rm, rl, cm.y, cl.y, l0, a1.y),
                           assign_score(sv.z, rm, rl, cm.z, cl.z, l0, a1.z), assign_score(sv.w, rm, rl, cm.w, cl.w, l0, a1.w)};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (sc[e] > best[e]) {
          best[e] = sc[e];
          bi[e] = i;
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    rb[ty][4 * tx + e] = best[e];
    ri[ty][4 * tx + e] = bi[e];
  }
  __syncthreads();
  if (threadIdx.x < 64 && blockIdx.x * 64 + threadIdx.x < n1) {
    const int c = threadIdx.x;
    float b = rb[0][c];
    int ix = ri[0][c];
    for (int k = 1; k < 64; ++k) {
      const float ob = rb[k][c];
      const int oi = ri[k][c];
      if (ob > b || (ob == b && oi < ix)) {
        b = ob;
        ix = oi;
      }
    }
    m1[o1 + blockIdx.x * 64 + c] = ix;
  }
}
__global__ void __launch_bounds__(1024) match_compact_batch_kernel(const LgAssign a, const float* __restrict__ max0,
                                                                   const int* __restrict__ m0, const int* __restrict__ m1,
                                                                   float filter, float thresh) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int pr = blockIdx.x;
  const int n0 = a.n0[pr], o0 = a.off0[pr], o1 = a.off1[pr];
  int* matches = a.matches[pr];
  float* mscores = a.mscores[pr];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n0; base += 1024) {
    const int i = base + threadIdx.x;
    float ms = 0.0f;
    int j = 0;
    if (i < n0) {
      j = m0[o0 + i];
      const bool mutual = (m1[o1 + j] == i);
      ms = mutual ? expf(max0[o0 + i]) : 0.0f;
    }
    const int keep = (i < n0 && ms > filter && ms > thresh) ? 1 : 0;
    int incl = keep;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
      int ws = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, ws, o);
        if (lane >= o) ws += t;
      }
      warp_sums[lane] = ws;
    }
    __syncthreads();
    const int pos = carry + (w ? warp_sums[w - 1] : 0) + incl - keep;
    if (keep) {
      matches[2 * pos] = i;
      matches[2 * pos + 1] = j;
      mscores[pos] = ms;
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry = pos + keep;
    __syncthreads();
  }
  if (threadIdx.x == 0) *a.count[pr] = carry;
}
// ---- assignment in two passes over sim (default) ------------------------------------------------------------------
// The five kernels above read the N0 x N1 similarity matrix six times (row statistics 2, column statistics 2, row arg-max 1,
// column arg-max 1; 576 MB of DRAM reads per 8 pairs).  Here a block owns a BAND of 32 rows of one pair and walks the
// columns in chunks of 256 (thread = column, 32 independent coalesced loads per chunk):
//   pass 1 (assign_stats_band): per-row (max, sum exp) AND the band's per-column partial (max, sum exp), both from one
//           32 x 256 shared-memory tile per chunk (column view: thread = column; row view: warp = 4 rows);
//           assign_col_merge folds the bands' partials into cmax / clog;
//   pass 2 (assign_argmax_band): S = log-assignment score, row arg-max the same way, per-column partial arg-max of the
//           band; assign_colarg_merge folds them into m1.
// Ties resolve to the lowest index exactly as TopK(k = 1) does (strict > in ascending index order, explicit index rule in the
// shuffles).  sim is read twice.
constexpr int kBandRows = 32;
constexpr int kBandCols = 256;
// A 32 x 256 tile of sim goes through shared memory once: the column view (thread = column, 32 conflict-free loads down
// the column) gives the band's column partials, the row view (warp = 4 rows, lane = 8 strided columns) the row statistics
// with ONE shuffle reduction per row and chunk.
__global__ void __launch_bounds__(256) assign_stats_band_kernel(const LgAssign a, float* __restrict__ rmax,
                                                                float* __restrict__ rlog, float* __restrict__ part_a,
                                                                float* __restrict__ part_b, int part_ld, int part_bands) {
  __shared__ float tile[kBandRows][kBandCols + 1];
  const int pr = blockIdx.y, band = blockIdx.x;
  const int n0 = a.n0[pr], n1 = a.n1[pr], ld = a.ld[pr];
  const int r0 = band * kBandRows;
  if (r0 >= n0) return;
  const int rows = min(kBandRows, n0 - r0);
  const float* sim = a.sim[pr] + static_cast<size_t>(r0) * ld;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* pa = part_a + (static_cast<size_t>(pr) * part_bands + band) * part_ld;
  float* pb = part_b + (static_cast<size_t>(pr) * part_bands + band) * part_ld;
  float run_m[4], run_s[4];                         // rows 4w .. 4w+3 of the band, this lane's columns
#pragma unroll
  for (int k = 0; k < 4; ++k) { run_m[k] = -INFINITY; run_s[k] = 0.0f; }
  for (int c0 = 0; c0 < n1; c0 += kBandCols) {
    const int j = c0 + threadIdx.x;
    const bool in = j < n1;
    __syncthreads();                                // the previous chunk's row view is done with the tile
#pragma unroll
    for (int i = 0; i < kBandRows; ++i) tile[i][threadIdx.x] = (in && i < rows) ? sim[static_cast<size_t>(i) * ld + j] : -INFINITY;
    __syncthreads();
    // column view: partial (max, sum exp) of my column over the band's rows
    {
      float cm = -INFINITY;
#pragma unroll
      for (int i = 0; i < kBandRows; ++i) cm = fmaxf(cm, tile[i][threadIdx.x]);
      float cs = 0.0f;
#pragma unroll
      for (int i = 0; i < kBandRows; ++i) cs += __expf(tile[i][threadIdx.x] - cm);     // rows beyond n0 hold -inf: exp = 0
      if (in) { pa[j] = cm; pb[j] = cs; }
    }
    // row view: online (max, sum exp) of rows 4w .. 4w+3 over columns lane, lane + 32, ...
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float* trow = tile[4 * w + k];
      float v[8];
      float m = -INFINITY;
#pragma unroll
      for (int q = 0; q < 8; ++q) { v[q] = trow[lane + 32 * q]; m = fmaxf(m, v[q]); }
      if (m > run_m[k]) { run_s[k] *= __expf(run_m[k] - m); run_m[k] = m; }       // exp(-inf) = 0 on the first chunk
      if (run_m[k] > -INFINITY) {
#pragma unroll
        for (int q = 0; q < 8; ++q) run_s[k] += __expf(v[q] - run_m[k]);
      }
    }
  }
  // merge the 32 lanes of every row (each lane holds its own reference maximum)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float m = warp_max(run_m[k]);
    const float sacc = warp_sum(run_m[k] > -INFINITY ? run_s[k] * __expf(run_m[k] - m) : 0.0f);
    const int i = 4 * w + k;
    if (lane == 0 && i < rows) {
      rmax[a.off0[pr] + r0 + i] = m;
      rlog[a.off0[pr] + r0 + i] = logf(sacc);
    }
  }
}
__global__ void __launch_bounds__(256) assign_col_merge_kernel(const LgAssign a, const float* __restrict__ part_a,
                                                               const float* __restrict__ part_b, int part_ld, int part_bands,
                                                               float* __restrict__ cmax, float* __restrict__ clog) {
  const int pr = blockIdx.y;
  const int j = blockIdx.x * 256 + threadIdx.x;
  const int n0 = a.n0[pr], n1 = a.n1[pr];
  if (j >= n1) return;
  const int bands = (n0 + kBandRows - 1) / kBandRows;
  const float* pa = part_a + static_cast<size_t>(pr) * part_bands * part_ld + j;
  const float* pb = part_b + static_cast<size_t>(pr) * part_bands * part_ld + j;
  float m = -INFINITY;
  for (int b = 0; b < bands; ++b) m = fmaxf(m, pa[static_cast<size_t>(b) * part_ld]);
  float sacc = 0.0f;
  for (int b = 0; b < bands; ++b) sacc += pb[static_cast<size_t>(b) * part_ld] * expf(pa[static_cast<size_t>(b) * part_ld] - m);
  cmax[a.off1[pr] + j] = m;
  clog[a.off1[pr] + j] = logf(sacc);
}
__global__ void __launch_bounds__(256) assign_argmax_band_kernel(const LgAssign a, const float* __restrict__ rmax,
                                                                 const float* __restrict__ rlog, const float* __restrict__ cmax,
                                                                 const float* __restrict__ clog, const float* __restrict__ ls,
                                                                 float* __restrict__ max0, int* __restrict__ m0,
                                                                 float* __restrict__ part_a, float* __restrict__ part_b,
                                                                 int part_ld, int part_bands, float* __restrict__ S_dbg) {
  __shared__ float tile[kBandRows][kBandCols + 1];       // log-assignment scores of the chunk
  __shared__ float s_rm[kBandRows], s_rl[kBandRows], s_a0[kBandRows];
  const int pr = blockIdx.y, band = blockIdx.x;
  const int n0 = a.n0[pr], n1 = a.n1[pr], ld = a.ld[pr];
  const int r0 = band * kBandRows;
  if (r0 >= n0) return;
  const int rows = min(kBandRows, n0 - r0);
  const int o0 = a.off0[pr], o1 = a.off1[pr];
  const float* sim = a.sim[pr] + static_cast<size_t>(r0) * ld;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x < kBandRows) {
    const bool ok = threadIdx.x < rows;
    s_rm[threadIdx.x] = ok ? rmax[o0 + r0 + threadIdx.x] : 0.0f;
    s_rl[threadIdx.x] = ok ? rlog[o0 + r0 + threadIdx.x] : 0.0f;
    s_a0[threadIdx.x] = ok ? ls[o0 + r0 + threadIdx.x] : 0.0f;
  }
  float* pa = part_a + (static_cast<size_t>(pr) * part_bands + band) * part_ld;
  int* pb = reinterpret_cast<int*>(part_b) + (static_cast<size_t>(pr) * part_bands + band) * part_ld;
  float run_b[4];                                   // rows 4w .. 4w+3: running best over this lane's columns (ascending)
  int run_i[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { run_b[k] = -INFINITY; run_i[k] = 0x7fffffff; }
  for (int c0 = 0; c0 < n1; c0 += kBandCols) {
    const int j = c0 + threadIdx.x;
    const bool in = j < n1;
    const float cm = in ? cmax[o1 + j] : 0.0f, cl = in ? clog[o1 + j] : 0.0f, a1 = in ? ls[o1 + j] : 0.0f;
    __syncthreads();                                // s_rm .. ready (first chunk) / previous row view done
    // column view while filling the tile: scores of my column, partial arg-max over the band's rows (lowest row on ties)
    float cb = -INFINITY;
    int ci = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < kBandRows; ++i) {
      float sc = -INFINITY;
      if (in && i < rows) {
        sc = assign_score(sim[static_cast<size_t>(i) * ld + j], s_rm[i], s_rl[i], cm, cl, s_a0[i], a1);
        if (S_dbg && pr == a.pairs - 1) S_dbg[static_cast<size_t>(r0 + i) * n1 + j] = sc;
        if (sc > cb) { cb = sc; ci = r0 + i; }
      }
      tile[i][threadIdx.x] = sc;
    }
    if (in) { pa[j] = cb; pb[j] = ci; }
    __syncthreads();
    // row view: columns lane, lane + 32, ... in ascending order: strict > keeps the lowest column among equals
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float* trow = tile[4 * w + k];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float sc = trow[lane + 32 * q];
        if (sc > run_b[k]) { run_b[k] = sc; run_i[k] = c0 + lane + 32 * q; }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float ob = run_b[k];
    int oi = run_i[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float tb = __shfl_xor_sync(0xffffffffu, ob, o);
      const int ti = __shfl_xor_sync(0xffffffffu, oi, o);
      if (tb > ob || (tb == ob && ti < oi)) { ob = tb; oi = ti; }
    }
    const int i = 4 * w + k;
    if (lane == 0 && i < rows) {
      max0[o0 + r0 + i] = ob;
      m0[o0 + r0 + i] = oi;
    }
  }
}
__global__ void __launch_bounds__(256) assign_colarg_merge_kernel(const LgAssign a, const float* __restrict__ part_a,
                                                                  const float* __restrict__ part_b, int part_ld, int part_bands,
                                                                  int* __restrict__ m1) {
  const int pr = blockIdx.y;
  const int j = blockIdx.x * 256 + threadIdx.x;
  const int n0 = a.n0[pr], n1 = a.n1[pr];
  if (j >= n1) return;
  const int bands = (n0 + kBandRows - 1) / kBandRows;
  const float* pa = part_a + static_cast<size_t>(pr) * part_bands * part_ld + j;
  const int* pb = reinterpret_cast<const int*>(part_b) + static_cast<size_t>(pr) * part_bands * part_ld + j;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int b = 0; b < bands; ++b) {                 // ascending bands = ascending rows: strict > keeps the lowest row
    const float ob = pa[static_cast<size_t>(b) * part_ld];
    if (ob > best) { best = ob; bi = pb[static_cast<size_t>(b) * part_ld]; }
  }
  m1[a.off1[pr] + j] = bi;
}
void launch_lg_assign_banded(cudaStream_t s, const LgAssign& a, float* rmax, float* rlog, float* cmax, float* clog,
                             const float* ls, float* max0, int* m0, int* m1, float filter, float thresh, float* S_dbg,
                             float* part_a, float* part_b, int part_ld, int part_bands) {
  if (a.pairs == 0) return;
  const dim3 gband((a.max_n0 + kBandRows - 1) / kBandRows, a.pairs), gcol((a.max_n1 + 255) / 256, a.pairs);
  assign_stats_band_kernel<<<gband, 256, 0, s>>>(a, rmax, rlog, part_a, part_b, part_ld, part_bands);
  assign_col_merge_kernel<<<gcol, 256, 0, s>>>(a, part_a, part_b, part_ld, part_bands, cmax, clog);
  assign_argmax_band_kernel<<<gband, 256, 0, s>>>(a, rmax, rlog, cmax, clog, ls, max0, m0, part_a, part_b, part_ld, part_bands, S_dbg);
  assign_colarg_merge_kernel<<<gcol, 256, 0, s>>>(a, part_a, part_b, part_ld, part_bands, m1);
  match_compact_batch_kernel<<<a.pairs, 1024, 0, s>>>(a, max0, m0, m1, filter, thresh);
}

void launch_lg_assign(cudaStream_t s, const LgAssign& a, float* rmax, float* rlog, float* cmax, float* clog,
                      const float* ls, float* max0, int* m0, int* m1, float filter, float thresh, float* S_dbg) {
  if (a.pairs == 0) return;
  const dim3 grow((a.max_n0 + 7) / 8, a.pairs), gcol((a.max_n1 + 31) / 32, a.pairs), gcol4((a.max_n1 + 63) / 64, a.pairs);
  row_lse_batch_kernel<<<grow, 256, 0, s>>>(a, rmax, rlog);
  col_lse_batch_kernel<<<gcol, 1024, 0, s>>>(a, cmax, clog);
  row_argmax_batch_kernel<<<grow, 256, 0, s>>>(a, rmax, rlog, cmax, clog, ls, max0, m0, S_dbg);
  col_argmax_batch_kernel<<<gcol4, 1024, 0, s>>>(a, rmax, rlog, cmax, clog, ls, m1);
  match_compact_batch_kernel<<<a.pairs, 1024, 0, s>>>(a, max0, m0, m1, filter, thresh);
}

}  // namespace rfe
