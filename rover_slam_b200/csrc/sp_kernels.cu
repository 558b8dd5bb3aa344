// SuperPoint kernels that are not tensor-core contractions:
//   conv1a (u8 -> 1/255 -> 3x3 conv 1->64 + ReLU, fp32 CUDA cores; K = 9 is too small for an MMA and the
//           weights reach +-197 so it is kept in exact fp32),
//   the 2-iteration 9x9 NMS + border + threshold (graph nodes 58-363),
//   the ordered (row-major) keypoint compaction (NonZero semantics, nodes 360-398),
//   the bilinear descriptor sampling + L2 normalisation (nodes 414-485).
// Reference: onnxmodel/superpoint.onnx as run by src/Extractors/superpoint_onnx.cc:133-136; SURVEY.md Appendix A.
#include "kernels.h"

namespace rfe {

// ------------------------------------------------------------------------------------------------
// conv1a: thread = (run of 8 pixels of one image row) x (8 output channels)
// ------------------------------------------------------------------------------------------------
// The 72 weights of the channel group live in registers as 36 channel PAIRS and every tap is one packed FFMA2
// (two IEEE fp32 FMAs per issue slot, same operation order as a scalar fmaf chain: bit-identical results).  The
// 3 x 10 input window of the run is loaded and scaled once and slides along x, so a pixel costs 36 FFMA2 + the
// bias / ReLU / split-fp16 epilogue instead of 72 FFMA + 9 guarded byte loads.  A warp covers 4 runs x 8 groups:
// each store instruction writes 4 full 128-byte lines per plane.  Bound: HBM writes (256 B out per 1 B in), so the number
// of warps in flight matters: __launch_bounds__(256, 2) caps the kernel at 128 registers (17 words spilled) for two blocks per
// SM: 577 -> 370 us per 16 frames.  (4 channels per thread at 80 registers / 3 blocks per SM: 456 us, more redundant loads.)
constexpr int kConv1aRun = 8;
__global__ void __launch_bounds__(256, 2) conv1a_kernel(const uint8_t* __restrict__ img, int stride, int H, int W, int B,
                                                     const float* __restrict__ w /*[64][9]*/,
                                                     const float* __restrict__ bias, __half* __restrict__ out_hi,
                                                     __half* __restrict__ out_lo) {
  const int cg = threadIdx.x & 7;
  f32x2 wr[4][9], br[4];
#pragma unroll
  for (int jp = 0; jp < 4; ++jp) {
    const int c = cg * 8 + 2 * jp;
    br[jp] = pk2(__ldg(bias + c), __ldg(bias + c + 1));
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[jp][t] = pk2(__ldg(w + c * 9 + t), __ldg(w + (c + 1) * 9 + t));
  }
  const unsigned runs_per_row = static_cast<unsigned>(W) / kConv1aRun;        // W is a multiple of 8
  const unsigned total = static_cast<unsigned>(B) * H * runs_per_row;
  const unsigned run = blockIdx.x * 32u + (threadIdx.x >> 3);
  if (run >= total) return;
  const unsigned rowi = run / runs_per_row;
  const int x0 = static_cast<int>(run - rowi * runs_per_row) * kConv1aRun;
  const int b = static_cast<int>(rowi / static_cast<unsigned>(H));
  const int y = static_cast<int>(rowi - static_cast<unsigned>(b) * H);
  const uint8_t* im = img + static_cast<size_t>(b) * H * stride;
  float in[3][kConv1aRun + 2];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int yy = y + dy - 1;
    const bool rv = yy >= 0 && yy < H;
#pragma unroll
    for (int i = 0; i < kConv1aRun + 2; ++i) {
      const int xx = x0 - 1 + i;
      float v = 0.0f;
      if (rv && xx >= 0 && xx < W)
        v = static_cast<float>(__ldg(im + static_cast<size_t>(yy) * stride + xx)) * 0.003921568859368563f;  // transform.cpp:8
      in[dy][i] = v;
    }
  }
  const size_t obase = (static_cast<size_t>(rowi) * W + x0) * 64 + cg * 8;
#pragma unroll
  for (int px = 0; px < kConv1aRun; ++px) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      f32x2 acc = pk2(0.0f, 0.0f);
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) acc = fma2(pk2(in[dy][px + dx], in[dy][px + dx]), wr[jp][dy * 3 + dx], acc);
      acc = add2(acc, br[jp]);
      float a0, a1;
      upk2(acc, a0, a1);
      split2(pk2(fmaxf(a0, 0.0f), fmaxf(a1, 0.0f)), hi[jp], lo[jp]);
    }
    *reinterpret_cast<uint4*>(out_hi + obase + static_cast<size_t>(px) * 64) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(out_lo + obase + static_cast<size_t>(px) * 64) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

void launch_conv1a(cudaStream_t s, const uint8_t* img, int stride, int H, int W, int B, const float* w,
                   const float* bias, __half* out_hi, __half* out_lo) {
  const size_t runs = static_cast<size_t>(B) * H * (W / kConv1aRun);
  conv1a_kernel<<<static_cast<unsigned>((runs + 31) / 32), 256, 0, s>>>(img, stride, H, W, B, w, bias, out_hi, out_lo);
}

// ------------------------------------------------------------------------------------------------
// NMS: fused tile kernel.  Tile 80x48 outputs, halo 20 = 5 chained 9x9 max-pools (211 KB of shared memory).
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int NT_W = 80, NT_H = 48, NHALO = 20, NR = 4;   // 640 = 8 x 80, 480 = 10 x 48: no ragged tiles; halo overhead 2.75x
constexpr int RW = NT_W + 2 * NHALO;  // 104
constexpr int RH = NT_H + 2 * NHALO;  // 72
constexpr int NMS_THREADS = 1024;

// out[i] = max(v[i .. i+8]) for i = 0..7 with 44 max operations (log-step doubling) instead of 64
__device__ __forceinline__ void max9_run8(const float (&v)[16], float (&out)[8]) {
  float a1[15], a2[13], a4[8];
#pragma unroll
  for (int i = 0; i < 15; ++i) a1[i] = fmaxf(v[i], v[i + 1]);
#pragma unroll
  for (int i = 0; i < 13; ++i) a2[i] = fmaxf(a1[i], a1[i + 2]);
#pragma unroll
  for (int i = 0; i < 8; ++i) a4[i] = fmaxf(a2[i], a2[i + 4]);
#pragma unroll
  for (int i = 0; i < 8; ++i) out[i] = fmaxf(a4[i], v[i + 8]);
}

// dst(y,x) = max over the 9x9 window of src, for (y,x) in the region shrunk by `e_out` from the full
// halo region; src must be valid on the region shrunk by e_out - 4.  tmp is scratch.  Separable; every thread
// produces a run of 8 outputs from 16 loads (2 shared-memory loads per output instead of 9).
__device__ __forceinline__ void pool9(const float* src, float* tmp, float* dst, int e_out) {
  const float NEG = -INFINITY;
  const int ex = e_out, ey = e_out - NR;  // horizontal pass rows: [ey, RH-ey), cols [ex, RW-ex)
  const int w = RW - 2 * ex, h = RH - 2 * ey;
  const int runs = (w + 7) >> 3;
  for (int i = threadIdx.x; i < runs * h; i += NMS_THREADS) {
    const int y = ey + i / runs, x0 = ex + (i % runs) * 8;
    const float* r = src + y * RW;
    float v[16], o[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int x = x0 - 4 + j;
      v[j] = x < RW ? r[x] : NEG;
    }
    max9_run8(v, o);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (x0 + j < RW - ex) tmp[y * RW + x0 + j] = o[j];
  }
  __syncthreads();
  const int h2 = RH - 2 * e_out;
  const int vruns = (h2 + 7) >> 3;
  for (int i = threadIdx.x; i < w * vruns; i += NMS_THREADS) {
    const int x = ex + i % w, y0 = e_out + (i / w) * 8;
    float v[16], o[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int y = y0 - 4 + j;
      v[j] = y < RH ? tmp[y * RW + x] : NEG;
    }
    max9_run8(v, o);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (y0 + j < RH - e_out) dst[(y0 + j) * RW + x] = o[j];
  }
  __syncthreads();
}
}  // namespace

__global__ void __launch_bounds__(NMS_THREADS, 1) nms_kernel(const float* __restrict__ heat, float* __restrict__ out,
                                                            int H, int W) {
  extern __shared__ float sm[];
  float* S = sm;                 // scores, -inf outside the image
  float* T = S + RW * RH;        // scratch (row pass)
  float* P = T + RW * RH;        // pool result
  float* MX = P + RW * RH;       // max_mask as 0/1 (0 outside the image)
  float* SS = MX + RW * RH;      // supp_scores (-inf outside the image)
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * NT_W - NHALO, y0 = blockIdx.y * NT_H - NHALO;
  const float* hb = heat + static_cast<size_t>(b) * H * W;
  const float NEG = -INFINITY;
  for (int i = threadIdx.x; i < RW * RH; i += NMS_THREADS) {
    const int y = y0 + i / RW, x = x0 + i % RW;
    const bool in = (y >= 0 && y < H && x >= 0 && x < W);
    S[i] = in ? hb[static_cast<size_t>(y) * W + x] : NEG;
    MX[i] = 0.0f;
  }
  __syncthreads();
  // max_mask = scores == max_pool(scores)                                   (nodes 59-62)
  pool9(S, T, P, 4);
  for (int i = threadIdx.x; i < RW * RH; i += NMS_THREADS) {
    const int ry = i / RW, rx = i % RW;
    if (ry >= 4 && ry < RH - 4 && rx >= 4 && rx < RW - 4) {
      const bool in = S[i] != NEG;
      MX[i] = (in && S[i] == P[i]) ? 1.0f : 0.0f;
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int it = 0; it < 2; ++it) {                                           // nodes 63-84
    const int e = 8 + it * 8;   // supp valid on shrink e; new mask valid on shrink e+4
    pool9(MX, T, P, e);         // P = max_pool(max_mask) ; supp_mask = P > 0
    for (int i = threadIdx.x; i < RW * RH; i += NMS_THREADS) {
      const int ry = i / RW, rx = i % RW;
      if (ry >= e && ry < RH - e && rx >= e && rx < RW - e) {
        const bool in = S[i] != NEG;
        SS[i] = in ? (P[i] > 0.0f ? 0.0f : S[i]) : NEG;      // where(supp, 0, scores)
        T[i] = P[i];                                           // keep supp for the update below
      }
    }
    __syncthreads();
    // stash supp (T gets clobbered by pool9's row pass) into P's place: reuse MX update in two steps
    // 1) copy supp flags into the sign of a side buffer: we use P after pooling SS, so save supp first.
    for (int i = threadIdx.x; i < RW * RH; i += NMS_THREADS) {
      const int ry = i / RW, rx = i % RW;
      if (ry >= e && ry < RH - e && rx >= e && rx < RW - e) {
        // encode supp into MX: 2.0 = (mask 0, supp), 3.0 = (mask 1, supp); mask value stays recoverable
        if (T[i] > 0.0f) MX[i] += 2.0f;
      }
    }
    __syncthreads();
    pool9(SS, T, P, e + 4);     // P = max_pool(supp_scores)
    for (int i = threadIdx.x; i < RW * RH; i += NMS_THREADS) {
      const int ry = i / RW, rx = i % RW;
      if (ry >= e && ry < RH - e && rx >= e && rx < RW - e) {
        float m = MX[i];
        const bool supp = m >= 2.0f;
        if (supp) m -= 2.0f;
        const bool inner = (ry >= e + 4 && ry < RH - e - 4 && rx >= e + 4 && rx < RW - e - 4);
        if (inner) {
          const bool in = S[i] != NEG;
          const bool new_max = in && (SS[i] == P[i]);
          if (new_max && !supp) m = 1.0f;
        }
        MX[i] = m;
      }
    }
    __syncthreads();
  }
  // scores = where(max_mask, scores, 0); borders -> -1                       (nodes 84-359)
  float* ob = out + static_cast<size_t>(b) * H * W;
  for (int i = threadIdx.x; i < NT_W * NT_H; i += NMS_THREADS) {
    const int ty = i / NT_W, tx = i % NT_W;
    const int y = blockIdx.y * NT_H + ty, x = blockIdx.x * NT_W + tx;
    if (y < H && x < W) {
      const int r = (ty + NHALO) * RW + tx + NHALO;
      float v = MX[r] == 1.0f ? S[r] : 0.0f;
      if (y < 4 || x < 4 || y >= H - 4 || x >= W - 4) v = -1.0f;
      ob[static_cast<size_t>(y) * W + x] = v;
    }
  }
}

int nms_smem_bytes() { return 5 * RW * RH * static_cast<int>(sizeof(float)); }

// ------------------------------------------------------------------------------------------------
// NMS, second form (default): the same graph nodes 58-359 on bit-planes.
//   * Only pixels above the detection threshold can become keypoints, and a pixel at or below it can neither suppress nor
//     out-score one above it (a maximum only ever suppresses pixels of lower-or-equal score, and `ss == maxpool(ss)` is only
//     lost to a higher unsuppressed score): the three max-pools therefore run on s_act = (s > thr ? s : 0) and the keypoints,
//     their scores and their order are exactly the graph's.  (The output map holds 0 instead of the score at sub-threshold
//     maxima, which the graph's own `s > 0.0005` selection never looks at.)
//   * max_mask / supp_mask are bit-planes (one word per 32 pixels): the two 9x9 pools of 0/1 planes become shifts and ORs;
//     only s_act and one row-pass plane are fp32 in shared memory (2 planes instead of 5).
//   * separable 9-wide maxima with FMNMX3: a1[i] = max3(v[i..i+2]), out[i] = max3(a1[i], a1[i+3], a1[i+6]): 22 operations
//     per 8 outputs; the row pass reads 16-byte vectors, the column pass one column per lane (conflict free) and packs its
//     verdicts with ballots.
//   * the per-row keypoint counts of the ordered compaction come out of this kernel (atomicAdd per tile row), so the separate
//     counting launch is gone.
// Tile 128 x 64 outputs + 20-pixel halo = 168 x 104 region, 512 threads, ~150 KB of shared memory.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int N2_W = 128, N2_H = 64, N2_HALO = 20;
constexpr int N2_RW = N2_W + 2 * N2_HALO;   // 168 = 21 runs of 8
constexpr int N2_RH = N2_H + 2 * N2_HALO;   // 104 = 13 runs of 8
constexpr int N2_PITCH = N2_RW + 8;         // 4 zero pad floats on either side of a row (16-byte aligned rows)
constexpr int N2_WORDS = 8;                 // bit-plane row: [0] = pad, [1..6] = 168 bits, [7] = pad
constexpr int N2_THREADS = 512;
constexpr int N2_SMEM = 2 * N2_RH * N2_PITCH * 4 + 4 * N2_RH * N2_WORDS * 4;

__device__ __forceinline__ float max3f(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
// out[i] = max(v[i .. i+8]), i = 0..7
__device__ __forceinline__ void max9_run8_v3(const float (&v)[16], float (&out)[8]) {
  float a1[14];
#pragma unroll
  for (int i = 0; i < 14; ++i) a1[i] = max3f(v[i], v[i + 1], v[i + 2]);
#pragma unroll
  for (int i = 0; i < 8; ++i) out[i] = max3f(a1[i], a1[i + 3], a1[i + 6]);
}
}  // namespace

__global__ void __launch_bounds__(N2_THREADS, 1) nms2_kernel(const float* __restrict__ heat, float* __restrict__ out,
                                                            int* __restrict__ row_cnt, int H, int W, float thr) {
  extern __shared__ __align__(16) float sm2[];
  float* S = sm2;                                   // s_act, [RH][PITCH], pixel (r, c) at r*PITCH + 4 + c
  float* T = S + N2_RH * N2_PITCH;                  // row-pass result
  uint32_t* ACT = reinterpret_cast<uint32_t*>(T + N2_RH * N2_PITCH);    // [RH][WORDS]
  uint32_t* MX = ACT + N2_RH * N2_WORDS;
  uint32_t* SUP = MX + N2_RH * N2_WORDS;
  uint32_t* HD = SUP + N2_RH * N2_WORDS;            // horizontal dilation scratch
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * N2_W - N2_HALO, y0 = blockIdx.y * N2_H - N2_HALO;
  const float* hb = heat + static_cast<size_t>(b) * H * W;

  // ---- load: s_act and the active bit-plane; pads and the other planes start at zero ----
  for (int i = tid; i < N2_RH * N2_WORDS; i += N2_THREADS) { MX[i] = 0; SUP[i] = 0; HD[i] = 0; ACT[i] = 0; }
  for (int i = tid; i < N2_RH * 8; i += N2_THREADS) {
    const int r = i >> 3, j = i & 7;
    S[r * N2_PITCH + (j < 4 ? j : N2_RW + j)] = 0.0f;
  }
  __syncthreads();
  // a warp takes two region rows per iteration: all twelve global loads are issued before the first is consumed (one load
  // in flight per warp made this loop 40 % of the kernel)
  for (int r2 = 2 * warp; r2 < N2_RH; r2 += 2 * (N2_THREADS / 32)) {
    float v[2][6];
#pragma unroll
    for (int dr = 0; dr < 2; ++dr)
#pragma unroll
      for (int w = 0; w < 6; ++w) {
        const int c = 32 * w + lane;
        const int y = y0 + r2 + dr, x = x0 + c;
        v[dr][w] = 0.0f;
        if (c < N2_RW && y >= 0 && y < H && x >= 0 && x < W) v[dr][w] = __ldg(hb + static_cast<size_t>(y) * W + x);
      }
#pragma unroll
    for (int dr = 0; dr < 2; ++dr)
#pragma unroll
      for (int w = 0; w < 6; ++w) {
        const int r = r2 + dr, c = 32 * w + lane;
        const float a = v[dr][w] > thr ? v[dr][w] : 0.0f;
        if (c < N2_RW) S[r * N2_PITCH + 4 + c] = a;
        const unsigned m = __ballot_sync(0xffffffffu, a > 0.0f);
        if (lane == 0) ACT[r * N2_WORDS + 1 + w] = m;
      }
  }
  __syncthreads();

#pragma unroll 1
  for (int stage = 0; stage < 3; ++stage) {
    if (stage > 0) {
      // ---- supp = 9x9 dilation of max_mask (nodes MaxPool(float(mask)) > 0): horizontal, then vertical ----
      for (int i = tid; i < N2_RH * 6; i += N2_THREADS) {
        const int r = i / 6, w = 1 + (i - r * 6);
        const uint32_t* row = MX + r * N2_WORDS;
        const uint32_t cur = row[w], L = row[w - 1], R = row[w + 1];
        uint32_t d = cur;
#pragma unroll
        for (int sft = 1; sft <= 4; ++sft) d |= (cur << sft) | (L >> (32 - sft)) | (cur >> sft) | (R << (32 - sft));
        HD[r * N2_WORDS + w] = d;
      }
      __syncthreads();
      for (int i = tid; i < N2_RH * 6; i += N2_THREADS) {
        const int r = i / 6, w = 1 + (i - r * 6);
        uint32_t d = 0;
#pragma unroll
        for (int dr = -4; dr <= 4; ++dr) {
          const int rr = r + dr;
          if (rr >= 0 && rr < N2_RH) d |= HD[rr * N2_WORDS + w];
        }
        SUP[r * N2_WORDS + w] = d;
      }
      __syncthreads();
    }
    // ---- row pass: T = 9-wide running maximum of ss = (supp ? 0 : s_act) ----
    for (int i = tid; i < N2_RH * 21; i += N2_THREADS) {
      const int r = i / 21, run = i - r * 21;
      const float4* src = reinterpret_cast<const float4*>(S + r * N2_PITCH + 8 * run);      // pixels 8run-4 .. 8run+11
      float v[16];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float4 t4 = src[q4];
        v[4 * q4] = t4.x; v[4 * q4 + 1] = t4.y; v[4 * q4 + 2] = t4.z; v[4 * q4 + 3] = t4.w;
      }
      if (stage > 0) {
        const int pos = 32 + 8 * run - 4;               // bit position of v[0] in the padded row (word 1 = pixels 0..31)
        const uint32_t* row = SUP + r * N2_WORDS;
        const uint32_t bits = __funnelshift_r(row[pos >> 5], row[(pos >> 5) + 1], pos & 31);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if ((bits >> j) & 1u) v[j] = 0.0f;
      }
      float o[8];
      max9_run8_v3(v, o);
      float4* dst = reinterpret_cast<float4*>(T + r * N2_PITCH + 4 + 8 * run);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
    __syncthreads();
    // ---- column pass + verdict: new maxima = active & ~supp & (ss == 9x9 max of ss) ----
    for (int t = warp; t < 13 * 6; t += N2_THREADS / 32) {
      const int rr = t / 6, w = t - rr * 6;
      const int c = 32 * w + lane, r0 = 8 * rr;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int r = r0 - 4 + j;
        v[j] = (c < N2_RW && r >= 0 && r < N2_RH) ? T[r * N2_PITCH + 4 + c] : 0.0f;
      }
      float o[8];
      max9_run8_v3(v, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = r0 + j;
        const uint32_t act = ACT[r * N2_WORDS + 1 + w], sup = SUP[r * N2_WORDS + 1 + w];
        const bool cand = ((act & ~sup) >> lane) & 1u;                 // active, not suppressed: ss = s_act there
        const bool is_max = cand && (c < N2_RW) && S[r * N2_PITCH + 4 + c] == o[j];
        const unsigned m = __ballot_sync(0xffffffffu, is_max);
        if (lane == 0 && m) MX[r * N2_WORDS + 1 + w] |= m;
      }
    }
    __syncthreads();
  }

  // ---- output: where(max_mask, scores, 0), borders -1, per-row keypoint counts ----
  float* ob = out + static_cast<size_t>(b) * H * W;
  for (int t = warp; t < N2_H * 4; t += N2_THREADS / 32) {
    const int ty = t >> 2, q = t & 3;
    const int tx = 32 * q + lane;
    const int y = blockIdx.y * N2_H + ty, x = blockIdx.x * N2_W + tx;
    const int r = ty + N2_HALO, c = tx + N2_HALO;
    bool kp = false;
    if (y < H && x < W) {
      const bool is_max = (MX[r * N2_WORDS + 1 + (c >> 5)] >> (c & 31)) & 1u;
      float v = is_max ? S[r * N2_PITCH + 4 + c] : 0.0f;
      if (y < 4 || x < 4 || y >= H - 4 || x >= W - 4) v = -1.0f;
      ob[static_cast<size_t>(y) * W + x] = v;
      kp = v > thr;
    }
    const unsigned m = __ballot_sync(0xffffffffu, kp);
    if (lane == 0 && m) atomicAdd(row_cnt + b * H + y, __popc(m));
  }
}

int nms2_prepare() {
  return cudaFuncSetAttribute(nms2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, N2_SMEM) == cudaSuccess ? 0 : 1;
}
// heat -> NMS map + per-row keypoint counts (row_cnt [B*H], zeroed here)
void launch_nms2(cudaStream_t s, const float* heat, float* out, int* row_cnt, int B, int H, int W, float thr) {
  cudaMemsetAsync(row_cnt, 0, sizeof(int) * static_cast<size_t>(B) * H, s);
  dim3 grid((W + N2_W - 1) / N2_W, (H + N2_H - 1) / N2_H, B);
  nms2_kernel<<<grid, N2_THREADS, N2_SMEM, s>>>(heat, out, row_cnt, H, W, thr);
}

void launch_nms(cudaStream_t s, const float* heat, float* out, int B, int H, int W) {
  dim3 grid((W + NT_W - 1) / NT_W, (H + NT_H - 1) / NT_H, B);
  nms_kernel<<<grid, NMS_THREADS, nms_smem_bytes(), s>>>(heat, out, H, W);
}

int nms_prepare() {
  return cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, nms_smem_bytes()) == cudaSuccess
             ? 0
             : 1;
}

// ------------------------------------------------------------------------------------------------
// Ordered compaction: idx = nonzero(s > 0.0005) in row-major order.
//   pass 1: one warp per image row counts; pass 2: one block per image scans the row counts;
//   pass 3: one warp per row writes (x, y), score at its offset.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kp_count_kernel(const float* __restrict__ nms, int H, int W, int B,
                                                       int* __restrict__ row_cnt, float thr) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= B * H) return;
  const float* r = nms + static_cast<size_t>(wid) * W;
  int c = 0;
  for (int x = lane; x < W; x += 32) c += (r[x] > thr) ? 1 : 0;
#pragma unroll
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) row_cnt[wid] = c;
}

__global__ void __launch_bounds__(1024) kp_scan_kernel(const int* __restrict__ row_cnt, int H, int* __restrict__ row_off,
                                                       int* __restrict__ counts) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < H; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < H ? row_cnt[b * H + i] : 0;
    int incl = v;
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
    const int prefix = carry + (w ? warp_sums[w - 1] : 0) + incl - v;
    if (i < H) row_off[b * H + i] = prefix;
    __syncthreads();
    if (threadIdx.x == 1023) carry = prefix + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[b] = carry;
}

__global__ void __launch_bounds__(256) kp_write_kernel(const float* __restrict__ nms, int H, int W, int B,
                                                       const int* __restrict__ row_off, float thr, int cap,
                                                       int* __restrict__ kpts, float* __restrict__ scores) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= B * H) return;
  const int b = wid / H, y = wid - b * H;
  const float* r = nms + static_cast<size_t>(wid) * W;
  int off = row_off[wid];
  for (int x0 = 0; x0 < W; x0 += 32) {
    const int x = x0 + lane;
    const float v = x < W ? r[x] : 0.0f;
    const bool keep = v > thr;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int idx = off + __popc(m & ((1u << lane) - 1));
      if (idx < cap) {
        kpts[(static_cast<size_t>(b) * cap + idx) * 2 + 0] = x;
        kpts[(static_cast<size_t>(b) * cap + idx) * 2 + 1] = y;
        scores[static_cast<size_t>(b) * cap + idx] = v;
      }
    }
    off += __popc(m);
  }
}

void launch_select(cudaStream_t s, const float* nms, int B, int H, int W, float thr, int cap, int* row_cnt,
                   int* row_off, int* counts, int* kpts, float* scores, bool have_counts) {
  const int warps = B * H;
  const int blocks = (warps * 32 + 255) / 256;
  if (!have_counts) kp_count_kernel<<<blocks, 256, 0, s>>>(nms, H, W, B, row_cnt, thr);    // nms2_kernel already counted
  kp_scan_kernel<<<B, 1024, 0, s>>>(row_cnt, H, row_off, counts);
  kp_write_kernel<<<blocks, 256, 0, s>>>(nms, H, W, B, row_off, thr, cap, kpts, scores);
}

// ------------------------------------------------------------------------------------------------
// Optional top-K cap (the `nfeatures` argument the reference stores and ignores, SPextractor.cc:84-146):
// keep the K highest-scoring keypoints of an image, in the reference's row-major order; among equal scores the earlier
// keypoint wins.  One block per image on the compacted list: a 4 x 8-bit radix select over the score bits finds the K-th
// largest score (scores are positive floats: their bit patterns order like the values), then an in-place stable compaction.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) kp_topk_kernel(int* __restrict__ counts, int* __restrict__ kpts,
                                                       float* __restrict__ scores, int cap, int K) {
  __shared__ unsigned hist[256];
  __shared__ unsigned sel_prefix, sel_remaining;
  __shared__ int warp_sums[2][32];
  __shared__ int carry[2];
  const int b = blockIdx.x;
  const int n = min(counts[b], cap);
  if (n <= K) return;
  int2* kp = reinterpret_cast<int2*>(kpts) + static_cast<size_t>(b) * cap;
  float* sc = scores + static_cast<size_t>(b) * cap;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  // ---- radix select: after the loop sel_prefix = bits of the K-th largest score, sel_remaining = how many keypoints with
  // exactly that score are kept
  if (tid == 0) { sel_prefix = 0; sel_remaining = static_cast<unsigned>(K); }
  __syncthreads();
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const unsigned prefix = sel_prefix;
    const unsigned himask = shift == 24 ? 0u : 0xFFFFFFFFu << (shift + 8);
    for (int i = tid; i < n; i += 1024) {
      const unsigned u = __float_as_uint(sc[i]);
      if ((u & himask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned rem = sel_remaining;
      int d = 255;
      for (; d > 0; --d) {                 // walk down from the largest digit
        if (hist[d] >= rem) break;
        rem -= hist[d];
      }
      sel_prefix = prefix | (static_cast<unsigned>(d) << shift);
      sel_remaining = rem;
    }
    __syncthreads();
  }
  const unsigned thr = sel_prefix;
  const int keep_eq = static_cast<int>(sel_remaining);
  // ---- stable in-place compaction, 1024 entries per round: keep (score > thr) and the first keep_eq with score == thr
  if (tid == 0) { carry[0] = 0; carry[1] = 0; }
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + tid;
    int2 k = make_int2(0, 0);
    float v = 0.0f;
    int eq = 0, gt = 0;
    if (i < n) {
      k = kp[i];
      v = sc[i];
      const unsigned u = __float_as_uint(v);
      gt = u > thr;
      eq = u == thr;
    }
    // two inclusive block scans at once: equals (to rank them) and kept-so-far is derived below
    int ie = eq, ig = gt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int te = __shfl_up_sync(0xffffffffu, ie, o), tg = __shfl_up_sync(0xffffffffu, ig, o);
      if (lane >= o) { ie += te; ig += tg; }
    }
    if (lane == 31) { warp_sums[0][w] = ie; warp_sums[1][w] = ig; }
    __syncthreads();
    if (w == 0) {
      int se = warp_sums[0][lane], sg = warp_sums[1][lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int te = __shfl_up_sync(0xffffffffu, se, o), tg = __shfl_up_sync(0xffffffffu, sg, o);
        if (lane >= o) { se += te; sg += tg; }
      }
      warp_sums[0][lane] = se;
      warp_sums[1][lane] = sg;
    }
    __syncthreads();
    const int eq_before = carry[0] + (w ? warp_sums[0][w - 1] : 0) + ie - eq;     // equals strictly before me
    const int gt_before = carry[1] + (w ? warp_sums[1][w - 1] : 0) + ig - gt;
    const bool keep = gt || (eq && eq_before < keep_eq);
    const int pos = gt_before + min(eq_before, keep_eq);                           // kept entries before me
    __syncthreads();                        // every read of this round is done before anything is overwritten
    if (keep) {
      kp[pos] = k;
      sc[pos] = v;
    }
    if (tid == 1023) { carry[0] = eq_before + eq; carry[1] = gt_before + gt; }
    __syncthreads();
  }
  if (tid == 0) counts[b] = K;
}
void launch_topk(cudaStream_t s, int B, int cap, int K, int* counts, int* kpts, float* scores) {
  if (K > 0 && B > 0) kp_topk_kernel<<<B, 1024, 0, s>>>(counts, kpts, scores, cap, K);
}

// ------------------------------------------------------------------------------------------------
// Descriptor sampling: one warp per keypoint, lane = 8 channels.
// grid = 2 * ((kp - 4 + 0.5) / (8*dim - 4 - 0.5)) - 1 ; grid_sample(bilinear, zeros, align_corners=True) ; L2 norm.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) desc_sample_kernel(const float* __restrict__ dense /*[B][h][w][256]*/,
                                                          const float* __restrict__ rowss /*[B*h*w][4] or null*/, int h,
                                                          int w, int B, const int* __restrict__ kpts,
                                                          const int* __restrict__ counts, int cap,
                                                          float* __restrict__ desc /*[B][cap][256]*/,
                                                          uint8_t* __restrict__ desc_bin /*[B][cap][256] or null*/) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= B * cap) return;
  const int b = wid / cap, i = wid - b * cap;
  const int n = min(counts[b], cap);
  if (i >= n) return;
  const float kx = static_cast<float>(kpts[(static_cast<size_t>(b) * cap + i) * 2 + 0]) - 4.0f + 0.5f;
  const float ky = static_cast<float>(kpts[(static_cast<size_t>(b) * cap + i) * 2 + 1]) - 4.0f + 0.5f;
  const float gx = (kx / (static_cast<float>(w * 8) - 4.0f - 0.5f)) * 2.0f - 1.0f;
  const float gy = (ky / (static_cast<float>(h * 8) - 4.0f - 0.5f)) * 2.0f - 1.0f;
  // align_corners=True un-normalisation: ((g + 1) / 2) * (size - 1)
  const float ix = ((gx + 1.0f) / 2.0f) * static_cast<float>(w - 1);
  const float iy = ((gy + 1.0f) / 2.0f) * static_cast<float>(h - 1);
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
  const int x1 = x0 + 1, y1 = y0 + 1;
  const float w_nw = (static_cast<float>(x1) - ix) * (static_cast<float>(y1) - iy);
  const float w_ne = (ix - static_cast<float>(x0)) * (static_cast<float>(y1) - iy);
  const float w_sw = (static_cast<float>(x1) - ix) * (iy - static_cast<float>(y0));
  const float w_se = (ix - static_cast<float>(x0)) * (iy - static_cast<float>(y0));
  const float* db = dense + static_cast<size_t>(b) * h * w * 256 + lane * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
  auto corner = [&](int yy, int xx, float wt) {
    if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
      const size_t pix = static_cast<size_t>(yy) * w + xx;
      const float4* p = reinterpret_cast<const float4*>(db + pix * 256);
      float4 a = __ldg(p), c = __ldg(p + 1);
      if (rowss) {
        // the descriptor head left the per-pixel L2 normalisation (nodes 402-413: d / max(||d||, 1e-12)) to the sampler:
        // the 1x1 conv stored d and four partial sums of squares per pixel
        const float4 q = __ldg(reinterpret_cast<const float4*>(rowss + (static_cast<size_t>(b) * h * w + pix) * 4));
        const float nrm = fmaxf(sqrtf(((q.x + q.y) + q.z) + q.w), 1e-12f);
        a.x /= nrm; a.y /= nrm; a.z /= nrm; a.w /= nrm;
        c.x /= nrm; c.y /= nrm; c.z /= nrm; c.w /= nrm;
      }
      acc[0] += a.x * wt; acc[1] += a.y * wt; acc[2] += a.z * wt; acc[3] += a.w * wt;
      acc[4] += c.x * wt; acc[5] += c.y * wt; acc[6] += c.z * wt; acc[7] += c.w * wt;
    }
  };
  corner(y0, x0, w_nw);
  corner(y0, x1, w_ne);
  corner(y1, x0, w_sw);
  corner(y1, x1, w_se);
  float ss = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) ss += acc[j] * acc[j];
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float nrm = fmaxf(sqrtf(ss), 1e-12f);
  float4* o = reinterpret_cast<float4*>(desc + (static_cast<size_t>(b) * cap + i) * 256 + lane * 8);
  const float4 d0 = make_float4(acc[0] / nrm, acc[1] / nrm, acc[2] / nrm, acc[3] / nrm);
  const float4 d1 = make_float4(acc[4] / nrm, acc[5] / nrm, acc[6] / nrm, acc[7] / nrm);
  o[0] = d0;
  o[1] = d1;
  if (desc_bin) {
    // sign binarisation for the DBoW3 feed (Frame::binarize_descriptors, Frame.cc:1034-1043: cv::threshold(row, 0, 1,
    // THRESH_BINARY) -> one uchar 0/1 per element), written while the descriptor is still in registers
    const uint32_t b0 = (d0.x > 0.0f ? 1u : 0u) | (d0.y > 0.0f ? 0x100u : 0u) | (d0.z > 0.0f ? 0x10000u : 0u) | (d0.w > 0.0f ? 0x1000000u : 0u);
    const uint32_t b1 = (d1.x > 0.0f ? 1u : 0u) | (d1.y > 0.0f ? 0x100u : 0u) | (d1.z > 0.0f ? 0x10000u : 0u) | (d1.w > 0.0f ? 0x1000000u : 0u);
    *reinterpret_cast<uint2*>(desc_bin + (static_cast<size_t>(b) * cap + i) * 256 + lane * 8) = make_uint2(b0, b1);
  }
}

void launch_desc_sample(cudaStream_t s, const float* dense, const float* rowss, int h, int w, int B, const int* kpts,
                        const int* counts, int cap, float* desc, uint8_t* desc_bin) {
  const int warps = B * cap;
  desc_sample_kernel<<<(warps * 32 + 255) / 256, 256, 0, s>>>(dense, rowss, h, w, B, kpts, counts, cap, desc, desc_bin);
}

// debug / tests: the normalised dense descriptor map when the normalisation is deferred to the sampler
__global__ void __launch_bounds__(256) dense_normalize_kernel(const float* __restrict__ dense, const float* __restrict__ rowss,
                                                              size_t npix, float* __restrict__ out) {
  const size_t wid = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= npix) return;
  const float4 q = __ldg(reinterpret_cast<const float4*>(rowss + wid * 4));
  const float nrm = fmaxf(sqrtf(((q.x + q.y) + q.z) + q.w), 1e-12f);
#pragma unroll
  for (int j = 0; j < 8; ++j) out[wid * 256 + j * 32 + lane] = dense[wid * 256 + j * 32 + lane] / nrm;
}
void launch_dense_normalize(cudaStream_t s, const float* dense, const float* rowss, size_t npix, float* out) {
  if (npix) dense_normalize_kernel<<<static_cast<unsigned>((npix * 32 + 255) / 256), 256, 0, s>>>(dense, rowss, npix, out);
}

// ------------------------------------------------------------------------------------------------
// Stand-alone sign binarisation of fp32 descriptors [n][256] -> u8 0/1 [n][256] (+ optional 256-bit packing [n][8] u32,
// bit j of word j/32 = element j): the same operation for descriptors that did not come out of desc_sample_kernel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) binarize_kernel(const float* __restrict__ desc, int n, uint8_t* __restrict__ out,
                                                       uint32_t* __restrict__ bits) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= n) return;
  const float* r = desc + static_cast<size_t>(wid) * 256;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const bool on = r[j * 32 + lane] > 0.0f;
    if (out) out[static_cast<size_t>(wid) * 256 + j * 32 + lane] = on ? 1 : 0;
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (bits && lane == 0) bits[static_cast<size_t>(wid) * 8 + j] = m;
  }
}
void launch_binarize(cudaStream_t s, const float* desc, int n, uint8_t* out, uint32_t* bits) {
  if (n > 0) binarize_kernel<<<(n * 32 + 255) / 256, 256, 0, s>>>(desc, n, out, bits);
}

// ------------------------------------------------------------------------------------------------
// Best / second-best L2 descriptor match over per-query candidate lists: the inner loop of SPmatcher::SearchByProjection /
// Fuse / Frame::ComputeStereoMatches (e.g. SPmatcher.cc:1225-1250 with DescriptorDistance_sp = cv::norm(a, b, NORM_L2),
// SPmatcher.cc:2184-2189).  One warp per query; candidates are visited in list order with the reference's update rule
// (strict <: the first of equal distances wins; both distances start at `init_dist`, 256 in the reference).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2_best2_kernel(const float* __restrict__ q, int nq, const float* __restrict__ db,
                                                       const int* __restrict__ cand_off, const int* __restrict__ cand_idx,
                                                       float init_dist, float* __restrict__ best_dist,
                                                       int* __restrict__ best_idx, float* __restrict__ second_dist,
                                                       int* __restrict__ second_idx) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= nq) return;
  const float4* q4 = reinterpret_cast<const float4*>(q + static_cast<size_t>(wid) * 256);
  const float4 qa = q4[lane], qb = q4[32 + lane];
  float b1 = init_dist, b2 = init_dist;
  int i1 = -1, i2 = -1;
  const int c0 = cand_off[wid], c1 = cand_off[wid + 1];
  for (int c = c0; c < c1; ++c) {
    const int idx = cand_idx[c];
    const float4* d4 = reinterpret_cast<const float4*>(db + static_cast<size_t>(idx) * 256);
    const float4 da = __ldg(d4 + lane), dbv = __ldg(d4 + 32 + lane);
    float t, acc = 0.0f;
    t = qa.x - da.x; acc = fmaf(t, t, acc);  t = qa.y - da.y; acc = fmaf(t, t, acc);
    t = qa.z - da.z; acc = fmaf(t, t, acc);  t = qa.w - da.w; acc = fmaf(t, t, acc);
    t = qb.x - dbv.x; acc = fmaf(t, t, acc); t = qb.y - dbv.y; acc = fmaf(t, t, acc);
    t = qb.z - dbv.z; acc = fmaf(t, t, acc); t = qb.w - dbv.w; acc = fmaf(t, t, acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const float dist = sqrtf(acc);
    if (dist < b1) {
      b2 = b1; i2 = i1;
      b1 = dist; i1 = idx;
    } else if (dist < b2) {
      b2 = dist; i2 = idx;
    }
  }
  if (lane == 0) {
    best_dist[wid] = b1;
    best_idx[wid] = i1;
    if (second_dist) second_dist[wid] = b2;
    if (second_idx) second_idx[wid] = i2;
  }
}
void launch_l2_best2(cudaStream_t s, const float* q, int nq, const float* db, const int* cand_off, const int* cand_idx,
                     float init_dist, float* best_dist, int* best_idx, float* second_dist, int* second_idx) {
  if (nq > 0)
    l2_best2_kernel<<<(nq * 32 + 255) / 256, 256, 0, s>>>(q, nq, db, cand_off, cand_idx, init_dist, best_dist, best_idx,
                                                          second_dist, second_idx);
}

}  // namespace rfe
