// rover_fe: context, weight upload, the SuperPoint / LightGlue launch sequences and the C ABI (include/rover_fe.h).
// Everything here is host orchestration; the arithmetic lives in umma_kernel.cuh, sp_kernels.cu, lg_kernels.cu.
#include "rover_fe.h"

#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "kernels.h"
#include "tensormap.h"
#include "attn_kernel.cuh"
#include "attn2_kernel.cuh"
#include "attn3_kernel.cuh"
#include "conv_strip.cuh"
#include "umma_kernel.cuh"
#include "weights.h"

using namespace rfe;

namespace {

constexpr float kAttnScale = 0.3535533845424652f;   // 64^-1/4 (lightglue_sim.onnx /inner_attn/Sqrt_1)
constexpr float kDetThreshold = 0.0005f;            // superpoint.onnx /Constant_116
constexpr float kFilterThreshold = 0.10000000149011612f;   // lightglue_sim.onnx /Constant_6
constexpr int kLayers = 9;
constexpr int kMaxPairs = rfe::kAttnMaxProblems / 2;

struct SplitBuf {   // device split-fp16 tensor
  __half* hi = nullptr;
  __half* lo = nullptr;
};
struct SplitW {     // device split-fp16 weight [n][k] (K contiguous) + fp32 bias
  SplitBuf w;
  float* bias = nullptr;
  int n = 0, k = 0;
};
struct LgLayer {
  SplitW wqkv, out_proj, s_ffn0, s_ffn3, to_qk, to_v, to_out, c_ffn0, c_ffn3;
  float *s_ln_w = nullptr, *s_ln_b = nullptr, *c_ln_w = nullptr, *c_ln_b = nullptr;
};

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

}  // namespace

struct rfe_ctx {
  int device = 0;
  int num_sms = 148;                     // SMs the persistent kernels spread over (rfe_set_sm_limit)
  int device_sms = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int max_batch = 8, max_h = 480, max_w = 768, cap = 4096;
  bool no_sp = false, no_lg = false;     // RFE_FLAG_NO_EXTRACTOR / RFE_FLAG_NO_MATCHER
  // layer-0 cache of the feature slots (rfe_lg_match_one_to_many): allocated on first use
  float* cache_x = nullptr;
  SplitBuf cache_cat;
  float *cache_cs = nullptr, *cache_sn = nullptr;
  std::vector<long long> slot_gen, cache_gen;   // contents generation of every slot / generation its cache entry was built from
  std::vector<int> cache_nh, cache_nw;          // normalisation size the entry was built with
  long long cache_hits = 0, cache_builds = 0;
  int fast = 0;                          // rfe_set_fast_mode: hi-only MMAs in SuperPoint's 3x3 convolutions (labelled, not parity)
  int topk = 0;                          // rfe_sp_set_topk: keep the K best keypoints per image (0 = all, the reference)
  std::vector<void*> allocs;
  long long launches = 0;
  double timer_extract_ms = 0.0, timer_match_ms = 0.0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;

  // ---- weights ----
  float *conv1a_w = nullptr, *conv1a_b = nullptr;
  SplitW c1b, c2a, c2b, c3a, c3b, c4a, c4b, cPa, cPb, cDa, cDb;
  float* posenc_w = nullptr;
  LgLayer layers[kLayers];
  SplitW final_proj;
  float *match_w = nullptr, *match_b = nullptr;

  // ---- SuperPoint buffers (sized for max_batch x max_h x max_w) ----
  uint8_t* img = nullptr;
  SplitBuf a1a, a1, a2a, a2, a3a, a3, a4a, feat, pa, da;
  float *heat = nullptr, *nmsmap = nullptr, *dense = nullptr;
  float* dense_ss = nullptr;    // [B*h/8*w/8][4] partial sums of squares of the un-normalised dense descriptors
  bool dense_deferred = false;  // the last extraction left `dense` un-normalised (rfe_debug_read normalises on the fly)
  int *row_cnt = nullptr, *row_off = nullptr;
  int* kp_counts = nullptr;     // [max_batch]
  int* kpts = nullptr;          // [max_batch][cap][2]
  float* kp_scores = nullptr;   // [max_batch][cap]
  float* desc = nullptr;        // [max_batch][cap][256]
  uint8_t* desc_bin = nullptr;  // [2*max_batch][cap][256] sign-binarised descriptors (0/1 per element)
  int last_batch = 0, last_h = 0, last_w = 0;
  int* h_counts = nullptr;      // pinned [max_batch]
  // pipelined pair matching (rfe_pairs_submit / rfe_pairs_collect): two feature-slot sets (set k = slots
  // [k*max_batch, (k+1)*max_batch)), up to two batches in flight
  struct Pending { int set, n_pairs, h, w; };
  Pending pending[2];
  int n_pending = 0, next_set = 0;
  cudaEvent_t ev_counts[2] = {nullptr, nullptr};
  // rfe_pairs_collect_begin_full: scores / descriptors of a batch go back on a SECOND stream, overlapping the matcher
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_feat[2] = {nullptr, nullptr};   // feature copies of slot set k have finished reading the slots
  bool feat_pending[2] = {false, false};
  int collecting_set = 0;
  cudaEvent_t ev_done = nullptr;     // results of the batch being collected have landed on the host
  int collecting_pairs = 0;          // > 0 between rfe_pairs_collect_begin and rfe_pairs_collect_end
  int collecting_rc = 0;
  int collecting_counts[2 * 16];
  int* h_counts2 = nullptr;     // pinned [2][max_batch] keypoint counts of the two sets
  int* h_mcounts = nullptr;     // pinned [max_batch] match counts
  unsigned long long bytes_h2d = 0, bytes_d2h = 0;   // host<->device traffic of the pipelined path (rfe_transfer_bytes)

  // ---- LightGlue buffers (sized for 2*cap rows) ----
  int lg_rows = 0, lg_ld = 0, lg_pairs = 1;   // row capacity, padded key count, pairs per batched match
  float *in_kpts = nullptr, *in_desc = nullptr;   // staging for the host API: [2*cap][2], [2*cap][256]
  float *cs = nullptr, *sn = nullptr;             // [rows][32]
  float* x = nullptr;                             // [rows][256]
  SplitBuf cat;                                   // [rows][512]
  SplitBuf q, k, vt;                              // q,k: [4][rows][64]; vt: [256][lg_ldv]
  int lg_ldv = 0;
  SplitBuf attn;                                  // [rows][256]
  float* hid = nullptr;                           // [rows][512]
  SplitBuf hs;                                    // [rows][512]
  SplitBuf md;                                    // [rows][256]
  float* sim = nullptr;                           // [cap][lg_ld]
  float *rmax = nullptr, *rlog = nullptr, *cmax = nullptr, *clog = nullptr, *ls = nullptr, *max0 = nullptr;
  int *m0 = nullptr, *m1 = nullptr;
  float *part_a = nullptr, *part_b = nullptr;   // banded assignment: per (pair, 32-row band, column) partials
  float* S_dbg = nullptr;
  const char* prof_tag = nullptr;            // $RFE_PROF_TAG: the launch tag whose UMMA role counters are recorded
  unsigned long long* attn_prof = nullptr;   // armed by rfe_debug_read("lg.attn_prof")
  float* attn_part_o = nullptr;              // key-range parts of the attention tail items (AttnParams::part_o / part_ml)
  float* attn_part_ml = nullptr;
  unsigned* attn_part_cnt = nullptr;
  int dbg_n0 = 0, dbg_n1 = 0, dbg_off0 = 0, dbg_off1 = 0, dbg_pair = 0;
  // match results: [max_batch] slots
  int* res_matches = nullptr;   // [slots][cap][2]
  float* res_scores = nullptr;  // [slots][cap]
  int* res_count = nullptr;     // [slots]

  // optional per-kernel CUDA-event profiling (rfe_profile)
  struct ProfRec { std::string tag; cudaEvent_t a, b; };
  bool use_strip_conv = true;   // RFE_CONV_STRIP=0 falls back to the 9-box implicit GEMM for the 64->64 layers
  // RFE_FUSE_CONV1A=1: conv1a computed inside conv1b's strip kernel (no activation round trip through HBM).  Correct and
  // bit-identical (tests), but measured SLOWER on B200: 1800 us per 16 frames against 381 + 904 us for the two kernels --
  // the two producer warps that fit beside the 160-register epilogue warps cannot deliver 2.25 rows per 8.3 K-cycle MMA
  // iteration (about 1650 instructions per thread and row).  Off by default; profiles/README.md has the numbers.
  bool fuse_conv1a = false;
  bool profiling = false;
  std::string prof_prefix;                   // rfe_profile_select
  std::vector<ProfRec> prof;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_pool;

  void* scr[2] = {nullptr, nullptr};         // grow-on-demand scratch of the host-in / host-out helpers
  size_t scr_bytes[2] = {0, 0};
  // debug scratch
  float* dbg = nullptr;
  size_t dbg_bytes = 0;
};

namespace {

// ------------------------------------------------------------------------------------------------
// allocation / upload helpers
// ------------------------------------------------------------------------------------------------
template <typename T>
int dev_alloc(rfe_ctx* c, T** p, size_t count) {
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, count * sizeof(T) + 256);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
    return RFE_ERR_CUDA;
  }
  cudaMemset(d, 0, count * sizeof(T) + 256);
  c->allocs.push_back(d);
  *p = static_cast<T*>(d);
  return RFE_OK;
}
int split_alloc(rfe_ctx* c, SplitBuf* b, size_t count) {
  int r = dev_alloc(c, &b->hi, count);
  if (r) return r;
  return dev_alloc(c, &b->lo, count);
}
int upload_f32(rfe_ctx* c, float** dst, const float* src, size_t n) {
  int r = dev_alloc(c, dst, n);
  if (r) return r;
  RFE_CUDA_CHECK(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice));
  return RFE_OK;
}
// rows of `src` ([n][k] fp32) are taken in the order perm[] (or identity) and stored split
int upload_split(rfe_ctx* c, SplitW* w, const float* src, const float* bias, int n, int k, const int* perm = nullptr) {
  std::vector<__half> hi(static_cast<size_t>(n) * k), lo(static_cast<size_t>(n) * k);
  std::vector<float> b(n);
  for (int i = 0; i < n; ++i) {
    const int si = perm ? perm[i] : i;
    for (int j = 0; j < k; ++j) split_f32(src[static_cast<size_t>(si) * k + j], hi[static_cast<size_t>(i) * k + j],
                                          lo[static_cast<size_t>(i) * k + j]);
    b[i] = bias ? bias[si] : 0.0f;
  }
  int r = split_alloc(c, &w->w, hi.size());
  if (r) return r;
  RFE_CUDA_CHECK(cudaMemcpy(w->w.hi, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice));
  RFE_CUDA_CHECK(cudaMemcpy(w->w.lo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice));
  if (bias) {
    r = upload_f32(c, &w->bias, b.data(), n);
    if (r) return r;
  }
  w->n = n;
  w->k = k;
  return RFE_OK;
}

int load_linear(rfe_ctx* c, const WeightBlob& blob, const std::string& name, SplitW* w, const int* perm = nullptr) {
  const HostTensor* t = blob.find(name + ".w");
  const HostTensor* b = blob.find(name + ".b");
  if (!t || t->dims.size() < 2) {
    set_error("weight '%s.w' missing from blob", name.c_str());
    return RFE_ERR_IO;
  }
  const int n = t->dims[0];
  const int k = static_cast<int>(t->size() / n);
  return upload_split(c, w, t->data, b ? b->data : nullptr, n, k, perm);
}
int load_vec(rfe_ctx* c, const WeightBlob& blob, const std::string& name, float** dst) {
  const HostTensor* t = blob.find(name);
  if (!t) {
    set_error("weight '%s' missing from blob", name.c_str());
    return RFE_ERR_IO;
  }
  return upload_f32(c, dst, t->data, t->size());
}

int load_weights(rfe_ctx* c, const char* path) {
  WeightBlob blob;
  if (blob.load(path)) return RFE_ERR_IO;
  int r;
  if (!c->no_sp) {     // a matcher-only ctx (RFE_FLAG_NO_EXTRACTOR) uploads no SuperPoint weights, and vice versa
  if ((r = load_vec(c, blob, "sp.conv1a.w", &c->conv1a_w))) return r;
  if ((r = load_vec(c, blob, "sp.conv1a.b", &c->conv1a_b))) return r;
  struct { const char* n; SplitW* w; } convs[] = {
      {"sp.conv1b", &c->c1b}, {"sp.conv2a", &c->c2a}, {"sp.conv2b", &c->c2b}, {"sp.conv3a", &c->c3a},
      {"sp.conv3b", &c->c3b}, {"sp.conv4a", &c->c4a}, {"sp.conv4b", &c->c4b}, {"sp.convPa", &c->cPa},
      {"sp.convPb", &c->cPb}, {"sp.convDa", &c->cDa}, {"sp.convDb", &c->cDb}};
  for (auto& e : convs)
    if ((r = load_linear(c, blob, e.n, e.w))) return r;   // OHWI flattens to [Cout][9*Cin]
  }
  if (c->no_lg) return RFE_OK;
  if ((r = load_vec(c, blob, "lg.posenc.w", &c->posenc_w))) return r;
  // Wqkv rows: ONNX column c = h*192 + d*3 + t  ->  ours t*256 + h*64 + d
  std::vector<int> perm(768);
  for (int t = 0; t < 3; ++t)
    for (int h = 0; h < 4; ++h)
      for (int d = 0; d < 64; ++d) perm[t * 256 + h * 64 + d] = h * 192 + d * 3 + t;
  for (int i = 0; i < kLayers; ++i) {
    LgLayer& L = c->layers[i];
    const std::string p = "lg.l" + std::to_string(i);
    if ((r = load_linear(c, blob, p + ".self.wqkv", &L.wqkv, perm.data()))) return r;
    if ((r = load_linear(c, blob, p + ".self.out_proj", &L.out_proj))) return r;
    if ((r = load_linear(c, blob, p + ".self.ffn0", &L.s_ffn0))) return r;
    if ((r = load_linear(c, blob, p + ".self.ffn3", &L.s_ffn3))) return r;
    if ((r = load_linear(c, blob, p + ".cross.to_qk", &L.to_qk))) return r;
    if ((r = load_linear(c, blob, p + ".cross.to_v", &L.to_v))) return r;
    if ((r = load_linear(c, blob, p + ".cross.to_out", &L.to_out))) return r;
    if ((r = load_linear(c, blob, p + ".cross.ffn0", &L.c_ffn0))) return r;
    if ((r = load_linear(c, blob, p + ".cross.ffn3", &L.c_ffn3))) return r;
    if ((r = load_vec(c, blob, p + ".self.ln.w", &L.s_ln_w))) return r;
    if ((r = load_vec(c, blob, p + ".self.ln.b", &L.s_ln_b))) return r;
    if ((r = load_vec(c, blob, p + ".cross.ln.w", &L.c_ln_w))) return r;
    if ((r = load_vec(c, blob, p + ".cross.ln.b", &L.c_ln_b))) return r;
  }
  if ((r = load_linear(c, blob, "lg.final_proj", &c->final_proj))) return r;
  if ((r = load_vec(c, blob, "lg.matchability.w", &c->match_w))) return r;
  if ((r = load_vec(c, blob, "lg.matchability.b", &c->match_b))) return r;
  return RFE_OK;
}

// ------------------------------------------------------------------------------------------------
// tensor-core launches
// ------------------------------------------------------------------------------------------------
struct ProfScope {   // records a CUDA event pair around one launch when profiling is on
  rfe_ctx* c;
  cudaEvent_t b = nullptr;
  ProfScope(rfe_ctx* ctx, const char* tag) : c(ctx) {
    if (!c->profiling || c->prof.size() >= 8192) return;
    if (!c->prof_prefix.empty() && strncmp(tag, c->prof_prefix.c_str(), c->prof_prefix.size()) != 0) return;
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    if (!c->prof_pool.empty()) {
      ev = c->prof_pool.back();
      c->prof_pool.pop_back();
    } else {
      cudaEventCreate(&ev.first);
      cudaEventCreate(&ev.second);
    }
    cudaEventRecord(ev.first, c->stream);
    b = ev.second;
    c->prof.push_back({tag, ev.first, ev.second});
  }
  ~ProfScope() {
    if (b) cudaEventRecord(b, c->stream);
  }
};

template <int BLOCK_N, int AMODE, int EPI, bool BRES = false, bool FAST = false>
int launch_umma(rfe_ctx* c, const char* tag, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                const CUtensorMap& b_lo, const UmmaParams& p, dim3 grid, const EpiMaps* em = nullptr) {
  static bool configured[64] = {};
  auto kern = umma_kernel<BLOCK_N, AMODE, EPI, BRES, FAST>;
  if (!configured[c->device & 63]) {
    RFE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, umma_smem_bytes(BLOCK_N, BRES)));
    configured[c->device & 63] = true;
  }
  // persistent launch: `grid` holds the tile counts (m, n, z); one CTA per SM walks the tiles round-robin
  UmmaParams pp = p;
  pp.tiles_m = static_cast<int>(grid.x);
  pp.tiles_n = static_cast<int>(grid.y);
  pp.num_tiles = static_cast<int>(grid.x * grid.y * grid.z);
  pp.prof = (c->attn_prof && c->prof_tag && !strcmp(tag, c->prof_tag)) ? c->attn_prof + 16 : nullptr;
  int ctas = pp.num_tiles < c->num_sms ? pp.num_tiles : c->num_sms;
  if (BRES) {   // B-resident: every CTA is bound to one n-tile, so launch a multiple of tiles_n (umma_kernel.cuh)
    if (grid.z != 1 || pp.num_k_steps > kBresMaxKSteps || pp.tiles_n > c->num_sms) {
      set_error("%s: B-resident GEMM needs an unbatched problem with K <= %d", tag, 64 * kBresMaxKSteps);
      return RFE_ERR_INVALID;
    }
    const int per_n = c->num_sms / pp.tiles_n < pp.tiles_m ? c->num_sms / pp.tiles_n : pp.tiles_m;
    ctas = per_n * pp.tiles_n;
  }
  ProfScope ps(c, tag);
  static const EpiMaps kNoMaps = {};
  kern<<<ctas, umma_threads(BLOCK_N, EPI), umma_smem_bytes(BLOCK_N, BRES), c->stream>>>(a_hi, a_lo, b_hi, b_lo, em ? *em : kNoMaps, pp);
  c->launches++;
  RFE_CUDA_CHECK(cudaGetLastError());
  return RFE_OK;
}

struct Operand {          // K-major split-fp16 matrix [batch][rows][k], row pitch ld, batch pitch bstride (elements)
  const __half* hi;
  const __half* lo;
  int rows, k;
  long long ld, bstride;
  int batch;
};

int make_operand_maps(const Operand& o, int box_rows, CUtensorMap* mh, CUtensorMap* ml) {
  const uint64_t dims[3] = {static_cast<uint64_t>(o.k), static_cast<uint64_t>(o.rows), static_cast<uint64_t>(o.batch)};
  const long long bs = o.batch > 1 ? o.bstride : static_cast<long long>(o.rows) * o.ld;
  const uint64_t strides[2] = {static_cast<uint64_t>(o.ld) * 2, static_cast<uint64_t>(bs) * 2};
  const uint32_t box[3] = {64, static_cast<uint32_t>(box_rows), 1};
  if (make_tmap_f16_sw128(mh, o.hi, 3, dims, strides, box)) return RFE_ERR_CUDA;
  if (make_tmap_f16_sw128(ml, o.lo, 3, dims, strides, box)) return RFE_ERR_CUDA;
  return RFE_OK;
}

// 3-D output map for the TMA-store epilogue: (cols, rows, batch) with a 32 x 32 box; fp32 tiles are staged with the
// 128-byte swizzle (32 x 4 B per row), fp16 planes with the 64-byte swizzle (32 x 2 B per row).
int make_out_map(CUtensorMap* m, const void* base, bool f32, int cols, int rows, long long ld, int batch, long long bstride) {
  const size_t es = f32 ? 4 : 2;
  const uint64_t dims[3] = {static_cast<uint64_t>(cols), static_cast<uint64_t>(rows), static_cast<uint64_t>(batch > 1 ? batch : 1)};
  const uint64_t strides[2] = {static_cast<uint64_t>(ld) * es,
                               static_cast<uint64_t>(batch > 1 ? bstride : static_cast<long long>(rows) * ld) * es};
  const uint32_t box[3] = {32, 32, 1};
  return make_tmap(m, f32 ? 1 : 0, f32 ? 3 : 2, base, 3, dims, strides, box) ? RFE_ERR_CUDA : RFE_OK;
}
// head-major planes [heads][rows][64]
int make_head_map(CUtensorMap* m, const void* base, int rows, long long head_stride, int heads) {
  const uint64_t dims[3] = {64, static_cast<uint64_t>(rows), static_cast<uint64_t>(heads)};
  const uint64_t strides[2] = {128, static_cast<uint64_t>(head_stride) * 2};
  const uint32_t box[3] = {32, 32, 1};
  return make_tmap(m, 0, 2, base, 3, dims, strides, box) ? RFE_ERR_CUDA : RFE_OK;
}

// transposed planes [cols][ld] (V^T): (rows contiguous, cols, batch) with a 32 x 32 box, 64-byte swizzle
int make_vt_map(CUtensorMap* m, const void* base, int rows, int cols, long long ld, int batch, long long bstride) {
  const uint64_t dims[3] = {static_cast<uint64_t>(rows), static_cast<uint64_t>(cols), static_cast<uint64_t>(batch > 1 ? batch : 1)};
  const uint64_t strides[2] = {static_cast<uint64_t>(ld) * 2,
                               static_cast<uint64_t>(batch > 1 ? bstride : static_cast<long long>(cols) * ld) * 2};
  const uint32_t box[3] = {32, 32, 1};
  return make_tmap(m, 0, 2, base, 3, dims, strides, box) ? RFE_ERR_CUDA : RFE_OK;
}

UmmaParams default_params() {
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.scale = 1.0f;
  return p;
}

// B-resident GEMM tiles (umma_kernel.cuh, BRES): RFE_BRES=0 in the environment falls back to the streamed-B kernel
// everywhere (A/B measurements), RFE_BRES=2 uses them even below one wave of tiles (tests), default: from one wave up.
static const int kBresMode = getenv("RFE_BRES") ? atoi(getenv("RFE_BRES")) : 1;
static bool use_bres(const rfe_ctx* c, unsigned tiles) {
  return kBresMode >= 2 || (kBresMode == 1 && tiles >= static_cast<unsigned>(c->num_sms));
}

// D = A * B^T with the LINEAR epilogue.  p carries the epilogue; M/N/K are filled here.
int gemm_linear(rfe_ctx* c, const char* tag, const Operand& A, const Operand& B, UmmaParams p, int block_n) {
  if (A.rows == 0 || B.rows == 0) return RFE_OK;
  CUtensorMap ah, al, bh, bl;
  int r;
  if ((r = make_operand_maps(A, kBlockM, &ah, &al))) return r;
  if ((r = make_operand_maps(B, block_n, &bh, &bl))) return r;
  p.num_k_steps = (A.k + 63) / 64;
  p.M = A.rows;
  p.N = B.rows;
  p.a_batched = A.batch > 1;
  p.b_batched = B.batch > 1;
  const int z = A.batch > B.batch ? A.batch : B.batch;
  dim3 grid((A.rows + kBlockM - 1) / kBlockM, (B.rows + block_n - 1) / block_n, z);
  EpiMaps em;
  memset(&em, 0, sizeof(em));
  if (p.out_f32 && (r = make_out_map(&em.f32, p.out_f32, true, p.N, p.M, p.ld_f32, z, p.bstride_f32))) return r;
  if (p.residual && (r = make_out_map(&em.res, p.residual, true, p.N, p.M, p.ld_res, z, p.bstride_res))) return r;
  if (p.out_hi && p.transpose_h) {
    if ((r = make_vt_map(&em.vt_hi, p.out_hi, p.M, p.N, p.ld_h, z, p.bstride_h))) return r;
    if ((r = make_vt_map(&em.vt_lo, p.out_lo, p.M, p.N, p.ld_h, z, p.bstride_h))) return r;
  }
  if (p.out_hi && !p.transpose_h) {
    if (p.head_major) {
      if ((r = make_head_map(&em.h_hi, p.out_hi, p.M, p.head_stride, p.N / 64))) return r;
      if ((r = make_head_map(&em.h_lo, p.out_lo, p.M, p.head_stride, p.N / 64))) return r;
    } else {
      if ((r = make_out_map(&em.h_hi, p.out_hi, false, p.N, p.M, p.ld_h, z, p.bstride_h))) return r;
      if ((r = make_out_map(&em.h_lo, p.out_lo, false, p.N, p.M, p.ld_h, z, p.bstride_h))) return r;
    }
  }
  if (block_n == 64) return launch_umma<64, A_GEMM, EPI_LINEAR>(c, tag, ah, al, bh, bl, p, grid, &em);
  // 128 x 256 tiles (K = 512 FFN GEMMs): one accumulator set (no MMA / epilogue overlap) but every A tile crosses L2 -> SM
  // once per 256 instead of once per 128 output columns -- these GEMMs are bound by operand ingest, not by the tensor pipe
  if (block_n == 256) return launch_umma<256, A_GEMM, EPI_LINEAR>(c, tag, ah, al, bh, bl, p, grid, &em);
  // K <= 256, one problem, enough m-tiles to amortise the weight load: keep the weights of one n-tile resident (BRES)
  if (z == 1 && p.num_k_steps <= kBresMaxKSteps && use_bres(c, grid.x * grid.y))
    return launch_umma<128, A_GEMM, EPI_LINEAR, true>(c, tag, ah, al, bh, bl, p, grid, &em);
  return launch_umma<128, A_GEMM, EPI_LINEAR>(c, tag, ah, al, bh, bl, p, grid, &em);
}

// 3x3 conv (pad 1) + bias + ReLU (+ 2x2 max-pool), NHWC split-fp16 in/out.
int conv3x3(rfe_ctx* c, const char* tag, const SplitBuf& in, int B, int H, int W, int Cin, const SplitW& w, const SplitBuf& out,
            bool pool) {
  const uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                            static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(Cin) * 2, static_cast<uint64_t>(W) * Cin * 2,
                               static_cast<uint64_t>(H) * W * Cin * 2};
  const uint32_t box[4] = {64, 16, 8, 1};
  CUtensorMap ah, al, bh, bl;
  if (make_tmap_f16_sw128(&ah, in.hi, 4, dims, strides, box)) return RFE_ERR_CUDA;
  if (make_tmap_f16_sw128(&al, in.lo, 4, dims, strides, box)) return RFE_ERR_CUDA;
  const int block_n = w.n == 64 ? 64 : 128;
  Operand Bop{w.w.hi, w.w.lo, w.n, w.k, w.k, 0, 1};
  int r;
  if ((r = make_operand_maps(Bop, block_n, &bh, &bl))) return r;
  UmmaParams p = default_params();
  p.num_k_steps = 9 * Cin / 64;
  p.cin_chunks = Cin / 64;
  p.N = w.n;
  p.H = H;
  p.W = W;
  p.tiles_x = (W + 15) / 16;
  p.tiles_y = (H + 7) / 8;
  p.bias = w.bias;
  p.out_hi = out.hi;
  p.out_lo = out.lo;
  p.pool = pool ? 1 : 0;
  dim3 grid(p.tiles_x * p.tiles_y * B, w.n / block_n, 1);
  if (c->fast) {          // labelled fast mode: separate instantiations, the exact kernels are not touched by it
    if (block_n == 64) return launch_umma<64, A_CONV3, EPI_CONV, false, true>(c, tag, ah, al, bh, bl, p, grid);
    return launch_umma<128, A_CONV3, EPI_CONV, false, true>(c, tag, ah, al, bh, bl, p, grid);
  }
  if (block_n == 64) return launch_umma<64, A_CONV3, EPI_CONV>(c, tag, ah, al, bh, bl, p, grid);
  return launch_umma<128, A_CONV3, EPI_CONV>(c, tag, ah, al, bh, bl, p, grid);
}

// 3x3 conv 64 -> 64 (+ReLU, + optional 2x2 max-pool) with the strip kernel (conv_strip.cuh).
// img != nullptr: conv1a is computed inside the kernel from the u8 image (conv_strip.cuh, FUSE1A) and `in` is not read.
int conv64_strip(rfe_ctx* c, const char* tag, const SplitBuf& in, int B, int H, int W, const SplitW& w, const SplitBuf& out,
                 bool pool, const uint8_t* img = nullptr, int img_stride = 0) {
  static bool configured[64] = {};
  if (!configured[c->device & 63]) {
    RFE_CUDA_CHECK(cudaFuncSetAttribute(conv64_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStripSmemBytes));
    RFE_CUDA_CHECK(cudaFuncSetAttribute(conv64_strip_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStripSmemBytes));
    RFE_CUDA_CHECK(cudaFuncSetAttribute(conv64_strip_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStripSmemBytes));
    configured[c->device & 63] = true;
  }
  const uint64_t dims[4] = {64, static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {128, static_cast<uint64_t>(W) * 128, static_cast<uint64_t>(H) * W * 128};
  const uint32_t box[4] = {64, 130, 1, 1};
  CUtensorMap ah, al, wh, wl;
  if (make_tmap_f16_sw128(&ah, in.hi, 4, dims, strides, box)) return RFE_ERR_CUDA;
  if (make_tmap_f16_sw128(&al, in.lo, 4, dims, strides, box)) return RFE_ERR_CUDA;
  Operand Wop{w.w.hi, w.w.lo, 64, 576, 576, 0, 1};
  int r;
  if ((r = make_operand_maps(Wop, 64, &wh, &wl))) return r;
  StripParams p;
  p.B = B;
  p.H = H;
  p.W = W;
  p.n_strips = (W + 127) / 128;
  p.seg_rows = 16;
  p.n_segs = (H + p.seg_rows - 1) / p.seg_rows;
  p.num_items = B * p.n_strips * p.n_segs;
  p.pool = pool ? 1 : 0;
  p.bias = w.bias;
  p.out_hi = out.hi;
  p.out_lo = out.lo;
  p.prof = (c->attn_prof && !strcmp(tag, "sp.conv1b")) ? c->attn_prof + 8 : nullptr;
  const int ctas = p.num_items < c->num_sms ? p.num_items : c->num_sms;
  ProfScope ps(c, tag);
  p.img = img;
  p.img_stride = img_stride;
  p.w1a = c->conv1a_w;
  p.b1a = c->conv1a_b;
  if (img) conv64_strip_fused_kernel<<<ctas, kStripThreadsFused, kStripSmemBytes, c->stream>>>(ah, al, wh, wl, p);
  else if (c->fast) conv64_strip_fast_kernel<<<ctas, kStripThreads, kStripSmemBytes, c->stream>>>(ah, al, wh, wl, p);
  else conv64_strip_kernel<<<ctas, kStripThreads, kStripSmemBytes, c->stream>>>(ah, al, wh, wl, p);
  c->launches++;
  RFE_CUDA_CHECK(cudaGetLastError());
  return RFE_OK;
}

// ------------------------------------------------------------------------------------------------
// SuperPoint: device in -> device-resident features
// ------------------------------------------------------------------------------------------------
int sp_run(rfe_ctx* c, const uint8_t* d_gray, int h, int w, int stride, int B, int slot_base = 0) {
  cudaStream_t s = c->stream;
  int r;
  if (c->no_sp) {
    set_error("this ctx was created with RFE_FLAG_NO_EXTRACTOR");
    return RFE_ERR_INVALID;
  }
  // conv1 group (K1): RFE_FUSE_CONV1A=1 computes conv1a inside conv1b's row producers, so its activation never touches
  // HBM; the default runs the stand-alone conv1a kernel first (faster as measured, see rfe_ctx::fuse_conv1a).
  const bool fuse1a = c->use_strip_conv && c->fuse_conv1a;
  if (!fuse1a) {
    ProfScope ps_(c, "sp.conv1a");
    launch_conv1a(s, d_gray, stride, h, w, B, c->conv1a_w, c->conv1a_b, c->a1a.hi, c->a1a.lo);
    c->launches++;
  }
  if (c->use_strip_conv) {
    if ((r = conv64_strip(c, "sp.conv1b", c->a1a, B, h, w, c->c1b, c->a1, true, fuse1a ? d_gray : nullptr, stride))) return r;
    if ((r = conv64_strip(c, "sp.conv2a", c->a1, B, h / 2, w / 2, c->c2a, c->a2a, false))) return r;
    if ((r = conv64_strip(c, "sp.conv2b", c->a2a, B, h / 2, w / 2, c->c2b, c->a2, true))) return r;
  } else {
    if ((r = conv3x3(c, "sp.conv1b", c->a1a, B, h, w, 64, c->c1b, c->a1, true))) return r;
    if ((r = conv3x3(c, "sp.conv2a", c->a1, B, h / 2, w / 2, 64, c->c2a, c->a2a, false))) return r;
    if ((r = conv3x3(c, "sp.conv2b", c->a2a, B, h / 2, w / 2, 64, c->c2b, c->a2, true))) return r;
  }
  if ((r = conv3x3(c, "sp.conv3a", c->a2, B, h / 4, w / 4, 64, c->c3a, c->a3a, false))) return r;
  if ((r = conv3x3(c, "sp.conv3b", c->a3a, B, h / 4, w / 4, 128, c->c3b, c->a3, true))) return r;
  const int hc = h / 8, wc = w / 8;
  if ((r = conv3x3(c, "sp.conv4a", c->a3, B, hc, wc, 128, c->c4a, c->a4a, false))) return r;
  if ((r = conv3x3(c, "sp.conv4b", c->a4a, B, hc, wc, 128, c->c4b, c->feat, false))) return r;
  if ((r = conv3x3(c, "sp.convPa", c->feat, B, hc, wc, 128, c->cPa, c->pa, false))) return r;
  if ((r = conv3x3(c, "sp.convDa", c->feat, B, hc, wc, 128, c->cDa, c->da, false))) return r;
  const int npix = B * hc * wc;
  {   // detector head: 1x1 conv 256->65 + softmax + depth-to-space
    Operand A{c->pa.hi, c->pa.lo, npix, 256, 256, 0, 1};
    Operand Bw{c->cPb.w.hi, c->cPb.w.lo, 65, 256, 256, 0, 1};
    CUtensorMap ah, al, bh, bl;
    if ((r = make_operand_maps(A, kBlockM, &ah, &al))) return r;
    if ((r = make_operand_maps(Bw, 80, &bh, &bl))) return r;
    UmmaParams p = default_params();
    p.num_k_steps = 4;
    p.M = npix;
    p.N = 65;
    p.H = hc;
    p.W = wc;
    p.bias = c->cPb.bias;
    p.out_f32 = c->heat;
    if ((r = launch_umma<80, A_GEMM, EPI_DET>(c, "sp.convPb_softmax", ah, al, bh, bl, p, dim3((npix + 127) / 128, 1, 1)))) return r;
  }
  // descriptor head.  Default: the 1x1 conv is a plain 128-wide double-buffered GEMM that stores d un-normalised plus four
  // partial sums of squares per pixel, and the sampler divides by the norm (the fused 256-wide epilogue had ONE accumulator
  // set and read it twice: 13 % tensor-pipe activity, 131 us per 16 frames).  RFE_CONVDB=1: the fused epilogue of round 1.
  static const int kConvDbMode = getenv("RFE_CONVDB") ? atoi(getenv("RFE_CONVDB")) : 2;
  if (kConvDbMode == 2) {
    Operand A{c->da.hi, c->da.lo, npix, 256, 256, 0, 1};
    Operand Bw{c->cDb.w.hi, c->cDb.w.lo, 256, 256, 256, 0, 1};
    UmmaParams p = default_params();
    p.bias = c->cDb.bias;
    p.out_f32 = c->dense;
    p.ld_f32 = 256;
    p.rowss = c->dense_ss;
    if ((r = gemm_linear(c, "sp.convDb", A, Bw, p, 128))) return r;
  } else {   // descriptor head: 1x1 conv 256->256 + L2 norm
    Operand A{c->da.hi, c->da.lo, npix, 256, 256, 0, 1};
    Operand Bw{c->cDb.w.hi, c->cDb.w.lo, 256, 256, 256, 0, 1};
    CUtensorMap ah, al, bh, bl;
    if ((r = make_operand_maps(A, kBlockM, &ah, &al))) return r;
    if ((r = make_operand_maps(Bw, 256, &bh, &bl))) return r;
    UmmaParams p = default_params();
    p.num_k_steps = 4;
    p.M = npix;
    p.N = 256;
    p.bias = c->cDb.bias;
    p.out_f32 = c->dense;
    p.ld_f32 = 256;
    if ((r = launch_umma<256, A_GEMM, EPI_DESC>(c, "sp.convDb_l2norm", ah, al, bh, bl, p, dim3((npix + 127) / 128, 1, 1)))) return r;
  }
  // RFE_NMS=1: the five-plane fp32 kernel of round 1 + a separate counting pass (A/B); default: bit-plane kernel with counts
  static const int kNmsMode = getenv("RFE_NMS") ? atoi(getenv("RFE_NMS")) : 2;
  {
    ProfScope ps_(c, "sp.nms");
    if (kNmsMode == 2) launch_nms2(s, c->heat, c->nmsmap, c->row_cnt, B, h, w, kDetThreshold);
    else launch_nms(s, c->heat, c->nmsmap, B, h, w);
  }
  int* kp_counts = c->kp_counts + slot_base;
  int* kpts = c->kpts + static_cast<size_t>(slot_base) * c->cap * 2;
  float* kp_scores = c->kp_scores + static_cast<size_t>(slot_base) * c->cap;
  float* desc = c->desc + static_cast<size_t>(slot_base) * c->cap * 256;
  { ProfScope ps_(c, "sp.select"); launch_select(s, c->nmsmap, B, h, w, kDetThreshold, c->cap, c->row_cnt, c->row_off, kp_counts, kpts,
                kp_scores, kNmsMode == 2); }
  if (c->topk > 0) {
    ProfScope ps_(c, "sp.topk");
    launch_topk(s, B, c->cap, c->topk, kp_counts, kpts, kp_scores);
    c->launches++;
  }
  { ProfScope ps_(c, "sp.desc_sample"); launch_desc_sample(s, c->dense, kConvDbMode == 2 ? c->dense_ss : nullptr, hc, wc, B, kpts, kp_counts, c->cap, desc,
                                                           c->desc_bin + static_cast<size_t>(slot_base) * c->cap * 256); }
  c->launches += kNmsMode == 2 ? 4 : 5;      // nms, (count), scan, write, desc_sample
  RFE_CUDA_CHECK(cudaGetLastError());
  c->dense_deferred = kConvDbMode == 2;
  for (int b = 0; b < B; ++b) c->slot_gen[slot_base + b]++;     // cached layer-0 state of these slots is stale now
  c->last_batch = B;
  c->last_h = h;
  c->last_w = w;
  return RFE_OK;
}

// ------------------------------------------------------------------------------------------------
// LightGlue: device-resident inputs (pixel keypoints fp32 [n][2], descriptors fp32 [n][256])
// Rows of image 0 live at [0, n0), rows of image 1 at [n0p, n0p + n1) with n0p = round_up(n0, 8).
// ------------------------------------------------------------------------------------------------
int ffn_block(rfe_ctx* c, int rows, const SplitW& ffn0, const float* ln_w, const float* ln_b, const SplitW& ffn3) {
  int r;
  {   // hid = cat @ ffn0^T + b
    Operand A{c->cat.hi, c->cat.lo, rows, 512, 512, 0, 1};
    Operand B{ffn0.w.hi, ffn0.w.lo, 512, 512, 512, 0, 1};
    UmmaParams p = default_params();
    p.bias = ffn0.bias;
    p.out_f32 = c->hid;
    p.ld_f32 = 512;
    static const int kFfn0Bn = getenv("RFE_FFN0_BN") ? atoi(getenv("RFE_FFN0_BN")) : 128;
    if ((r = gemm_linear(c, "lg.ffn0", A, B, p, kFfn0Bn))) return r;
  }
  { ProfScope ps_(c, "lg.ln_gelu"); launch_ln_gelu_split(c->stream, c->hid, rows, ln_w, ln_b, c->hs.hi, c->hs.lo); }
  c->launches++;
  {   // x = x + hs @ ffn3^T + b ; refresh split(x) in cat[:, 0:256]
    Operand A{c->hs.hi, c->hs.lo, rows, 512, 512, 0, 1};
    Operand B{ffn3.w.hi, ffn3.w.lo, 256, 512, 512, 0, 1};
    UmmaParams p = default_params();
    p.bias = ffn3.bias;
    p.residual = c->x;
    p.ld_res = 256;
    p.out_f32 = c->x;
    p.ld_f32 = 256;
    p.out_hi = c->cat.hi;
    p.out_lo = c->cat.lo;
    p.ld_h = 512;
    static const int kFfn3Bn = getenv("RFE_FFN3_BN") ? atoi(getenv("RFE_FFN3_BN")) : 128;
    if ((r = gemm_linear(c, "lg.ffn3", A, B, p, kFfn3Bn))) return r;
  }
  return RFE_OK;
}

// Fused attention over all problems of a block (self: every image against itself; cross: both directions of every pair).
int attention_fused(rfe_ctx* c, const char* tag, const SplitBuf& Q, const SplitBuf& K, int rows_total,
                    const AttnParams& problems, int nprob, int max_nq) {
  static bool configured[64] = {};
  // RFE_ATTN=1: the one-item-per-CTA two-pass kernel of round 1; 2 (default): its persistent form; 3: the persistent
  // single-pass (online-softmax) kernel -- correct, but measured slower (profiles/r02_attn3_online_softmax_prof.txt: the
  // softmax warps, not the tensor pipe, bound this kernel, and the per-tile maximum + pair exchange costs them more than the
  // separate maximum pass did).  Kept for A/B measurements.
  static const int kAttnMode = getenv("RFE_ATTN") ? atoi(getenv("RFE_ATTN")) : 2;
  if (!configured[c->device & 63]) {
    RFE_CUDA_CHECK(cudaFuncSetAttribute(attn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
    RFE_CUDA_CHECK(cudaFuncSetAttribute(attn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
    RFE_CUDA_CHECK(cudaFuncSetAttribute(attn2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttn2SmemBytes));
    RFE_CUDA_CHECK(cudaFuncSetAttribute(attn2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttn2SmemBytes));
    RFE_CUDA_CHECK(cudaFuncSetAttribute(attn3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kA3SmemBytes));
    RFE_CUDA_CHECK(cudaFuncSetAttribute(attn3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kA3SmemBytes));
    configured[c->device & 63] = true;
  }
  CUtensorMap qh, ql, kh, kl, vh, vl;
  const uint64_t dq[3] = {64, static_cast<uint64_t>(rows_total), 4};
  const uint64_t sq[2] = {128, static_cast<uint64_t>(rows_total) * 128};
  const uint32_t boxq[3] = {64, 128, 1}, boxk[3] = {64, 64, 1};
  if (make_tmap_f16_sw128(&qh, Q.hi, 3, dq, sq, boxq) || make_tmap_f16_sw128(&ql, Q.lo, 3, dq, sq, boxq) ||
      make_tmap_f16_sw128(&kh, K.hi, 3, dq, sq, boxk) || make_tmap_f16_sw128(&kl, K.lo, 3, dq, sq, boxk))
    return RFE_ERR_CUDA;
  const uint64_t dv[3] = {static_cast<uint64_t>(rows_total), 64, 4};
  const uint64_t sv[2] = {static_cast<uint64_t>(c->lg_ldv) * 2, 64ULL * c->lg_ldv * 2};
  if (make_tmap_f16_sw128(&vh, c->vt.hi, 3, dv, sv, boxk) || make_tmap_f16_sw128(&vl, c->vt.lo, 3, dv, sv, boxk))
    return RFE_ERR_CUDA;
  AttnParams p = problems;
  p.out_hi = c->attn.hi;
  p.out_lo = c->attn.lo;
  p.prof = c->attn_prof;
  p.prof_cta = getenv("RFE_ATTN_PROF_CTA") ? atoi(getenv("RFE_ATTN_PROF_CTA")) : 0;
  dim3 grid((max_nq + 127) / 128, 4, nprob);
  p.nprob = nprob;
  p.item_prefix[0] = 0;
  for (int z = 0; z < nprob; ++z) p.item_prefix[z + 1] = p.item_prefix[z] + 4 * ((p.nq[z] + 127) / 128);
  ProfScope ps(c, tag);
  if (kAttnMode == 3) {
    const int items = p.item_prefix[nprob];
    const int ctas = items < c->num_sms ? items : c->num_sms;
    if (p.prof) attn3_kernel<true><<<ctas, kAttnThreads, kA3SmemBytes, c->stream>>>(qh, ql, kh, kl, vh, vl, p);
    else attn3_kernel<false><<<ctas, kAttnThreads, kA3SmemBytes, c->stream>>>(qh, ql, kh, kl, vh, vl, p);
  } else if (kAttnMode == 2) {
    const int items = p.item_prefix[nprob];
    const int ctas = items < c->num_sms ? items : c->num_sms;
    // Tail balancing (RFE_ATTN_SPLIT=0 switches it off): with `items` = 7.35 x CTAs the last round keeps a third of the SMs
    // busy for a whole item; its items are cut into S = CTAs / tail parts along the keys instead (see AttnParams).
    static const bool kSplit = !(getenv("RFE_ATTN_SPLIT") && atoi(getenv("RFE_ATTN_SPLIT")) == 0);
    const int tail = items % ctas;
    int S = (kSplit && tail > 0) ? ctas / tail : 1;
    if (S > 4) S = 4;
    int min_nk = 1 << 30;
    for (int z = 0; z < nprob; ++z) min_nk = p.nk[z] < min_nk ? p.nk[z] : min_nk;
    while (S > 1 && min_nk < 256 * S) --S;                  // every part keeps at least two 128-key tiles
    if (S > 1 && (!c->attn_part_o || tail * S > kAttnPartSlots)) S = 1;
    p.split_s = S;
    p.split_first = S > 1 ? items - tail : items;
    p.n_items = S > 1 ? items - tail + tail * S : items;
    p.part_o = c->attn_part_o;
    p.part_ml = c->attn_part_ml;
    p.part_cnt = c->attn_part_cnt;
    // RFE_ATTN_CFG: softmax groups / ring depths (A/B; default 4): 0 = 2 groups of 8 warps, K 4 / V 3 / P 2 buffers; 1 = 2 groups, 3/2/3;
    // 2 = 4 groups of 4 warps, 3/2/3; 3 = 4 groups, 4/3/2; 4 = as 0 with two score buffers instead of three (the tensor pipe is one
    // queue: a score issuer that cannot run three tiles ahead delays the P V products less; +1.5 %)
    static const int kCfgEnv = getenv("RFE_ATTN_CFG") ? atoi(getenv("RFE_ATTN_CFG")) : 4;
    // the four-group forms need at least four pass-2 tiles and two pass-1 tiles per item (barrier discipline, attn2_kernel.cuh)
    const int kCfg = ((kCfgEnv == 2 || kCfgEnv == 3) && min_nk < 256) ? 0 : kCfgEnv;
    auto launch = [&](auto kern, int smem) -> cudaError_t {
      static const void* configured_fn[16][8] = {};      // [device][instantiation]: opt-in shared memory set once
      const void** slot = configured_fn[c->device & 15];
      int i = 0;
      while (i < 8 && slot[i] && slot[i] != reinterpret_cast<const void*>(kern)) ++i;
      if (i == 8 || !slot[i]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        if (i < 8) slot[i] = reinterpret_cast<const void*>(kern);
      }
      kern<<<ctas, kAttnThreads, smem, c->stream>>>(qh, ql, kh, kl, vh, vl, p);
      return cudaSuccess;
    };
    cudaError_t le;
    if (p.prof) {
      if (kCfg == 1) le = launch(attn2_kernel<true, 2, 3, 2, 3>, attn2_smem_bytes(3, 2, 3));
      else if (kCfg == 2) le = launch(attn2_kernel<true, 4, 3, 2, 3>, attn2_smem_bytes(3, 2, 3));
      else if (kCfg == 3) le = launch(attn2_kernel<true, 4, 4, 3, 2>, attn2_smem_bytes(4, 3, 2));
      else if (kCfg == 4) le = launch(attn2_kernel<true, 2, 4, 3, 2, 2>, attn2_smem_bytes(4, 3, 2));
      else le = launch(attn2_kernel<true>, kAttn2SmemBytes);
    } else {
      if (kCfg == 1) le = launch(attn2_kernel<false, 2, 3, 2, 3>, attn2_smem_bytes(3, 2, 3));
      else if (kCfg == 2) le = launch(attn2_kernel<false, 4, 3, 2, 3>, attn2_smem_bytes(3, 2, 3));
      else if (kCfg == 3) le = launch(attn2_kernel<false, 4, 4, 3, 2>, attn2_smem_bytes(4, 3, 2));
      else if (kCfg == 4) le = launch(attn2_kernel<false, 2, 4, 3, 2, 2>, attn2_smem_bytes(4, 3, 2));
      else le = launch(attn2_kernel<false>, kAttn2SmemBytes);
    }
    RFE_CUDA_CHECK(le);
  } else if (p.prof) attn_kernel<true><<<grid, kAttnThreads, kAttnSmemBytes, c->stream>>>(qh, ql, kh, kl, vh, vl, p);
  else attn_kernel<false><<<grid, kAttnThreads, kAttnSmemBytes, c->stream>>>(qh, ql, kh, kl, vh, vl, p);
  c->launches++;
  RFE_CUDA_CHECK(cudaGetLastError());
  return RFE_OK;
}

// One self-attention block (Wqkv + rotary, fused attention of every image against itself, out_proj, FFN) over `rows` rows.
int lg_self_block(rfe_ctx* c, const LgLayer& L, int rows, const AttnParams& self_p, int nprob, int max_n, int lin_bn) {
  int r;
  const long long hs = static_cast<long long>(rows) * 64;
  // ---------------- self attention ----------------
  {   // qkv = x Wqkv^T + b ; rotary on q,k ; head split ; V transposed -- all in the GEMM epilogue
    Operand A{c->cat.hi, c->cat.lo, rows, 256, 512, 0, 1};
    Operand B{L.wqkv.w.hi, L.wqkv.w.lo, 768, 256, 256, 0, 1};
    CUtensorMap ah, al, bh, bl;
    if ((r = make_operand_maps(A, kBlockM, &ah, &al))) return r;
    if ((r = make_operand_maps(B, 128, &bh, &bl))) return r;
    UmmaParams p = default_params();
    p.num_k_steps = 4;
    p.M = rows;
    p.N = 768;
    p.bias = L.wqkv.bias;
    p.scale = kAttnScale;
    p.cs = c->cs;
    p.sn = c->sn;
    p.out_hi = c->q.hi;
    p.out_lo = c->q.lo;
    p.k_hi = c->k.hi;
    p.k_lo = c->k.lo;
    p.vt_hi = c->vt.hi;
    p.vt_lo = c->vt.lo;
    p.ldv = c->lg_ldv;
    p.head_stride = hs;
    EpiMaps em;
    memset(&em, 0, sizeof(em));
    if ((r = make_head_map(&em.h_hi, c->q.hi, rows, hs, 4)) || (r = make_head_map(&em.h_lo, c->q.lo, rows, hs, 4)) ||
        (r = make_head_map(&em.k_hi, c->k.hi, rows, hs, 4)) || (r = make_head_map(&em.k_lo, c->k.lo, rows, hs, 4)) ||
        (r = make_vt_map(&em.vt_hi, c->vt.hi, rows, 256, c->lg_ldv, 1, 0)) || (r = make_vt_map(&em.vt_lo, c->vt.lo, rows, 256, c->lg_ldv, 1, 0)))
      return r;
    const dim3 tiles((rows + 127) / 128, 6, 1);
    if (use_bres(c, tiles.x * tiles.y))
      r = launch_umma<128, A_GEMM, EPI_QKV, true>(c, "lg.wqkv_rope", ah, al, bh, bl, p, tiles, &em);
    else
      r = launch_umma<128, A_GEMM, EPI_QKV>(c, "lg.wqkv_rope", ah, al, bh, bl, p, tiles, &em);
    if (r) return r;
  }
  if ((r = attention_fused(c, "lg.attn_self", c->q, c->k, rows, self_p, nprob, max_n))) return r;
  {
    Operand A{c->attn.hi, c->attn.lo, rows, 256, 256, 0, 1};
    Operand B{L.out_proj.w.hi, L.out_proj.w.lo, 256, 256, 256, 0, 1};
    UmmaParams p = default_params();
    p.bias = L.out_proj.bias;
    p.out_hi = c->cat.hi + 256;
    p.out_lo = c->cat.lo + 256;
    p.ld_h = 512;
    if ((r = gemm_linear(c, "lg.out_proj", A, B, p, lin_bn))) return r;
  }
  if ((r = ffn_block(c, rows, L.s_ffn0, L.s_ln_w, L.s_ln_b, L.s_ffn3))) return r;
  return RFE_OK;
}

// ---- per-slot layer-0 cache ----------------------------------------------------------------------
// LocalMapping matches ONE KeyFrame against up to ten neighbours (LocalMapping.cc:522-634 -> SearchForTriangulation ->
// MatchingPoints_onnx per neighbour), the reference re-uploading and re-projecting the same KeyFrame every time.  What
// depends on an image alone -- positional encoding, x = desc, and the whole first self-attention block -- is computed once
// per (slot contents, normalisation size) and kept: x after the block, its split copy, cos / sin.
int cache_alloc(rfe_ctx* c) {
  if (c->cache_x) return RFE_OK;
  const size_t slots = 2 * static_cast<size_t>(c->max_batch), cap = c->cap;
  int r;
  if ((r = dev_alloc(c, &c->cache_x, slots * cap * 256)) || (r = split_alloc(c, &c->cache_cat, slots * cap * 256)) ||
      (r = dev_alloc(c, &c->cache_cs, slots * cap * 32)) || (r = dev_alloc(c, &c->cache_sn, slots * cap * 32)))
    return r;
  c->cache_gen.assign(slots, -1);
  c->cache_nh.assign(slots, 0);
  c->cache_nw.assign(slots, 0);
  return RFE_OK;
}
void cache_move(rfe_ctx* c, LgCacheMove& mv, int max_n, int to_cache) {
  mv.x = c->x; mv.cat_hi = c->cat.hi; mv.cat_lo = c->cat.lo; mv.cs = c->cs; mv.sn = c->sn;
  mv.cx = c->cache_x; mv.ccat_hi = c->cache_cat.hi; mv.ccat_lo = c->cache_cat.lo; mv.ccs = c->cache_cs; mv.csn = c->cache_sn;
  mv.cap = c->cap;
  mv.to_cache = to_cache;
  ProfScope ps_(c, to_cache ? "lg.cache_store" : "lg.cache_load");
  launch_lg_cache_move(c->stream, mv, max_n);
  c->launches++;
}
// Build the cache entries of `count` slots (n[i] keypoints each) in one pass: prepare + layer-0 self block over all of them.
int lg_build_cache(rfe_ctx* c, const int* slots, const int* n, int count, int norm_h, int norm_w) {
  int r;
  if ((r = cache_alloc(c))) return r;
  LgImages im;
  LgCacheMove mv;
  AttnParams self_p;
  memset(&im, 0, sizeof(im));
  memset(&mv, 0, sizeof(mv));
  memset(&self_p, 0, sizeof(self_p));
  int rows = 0, mx = 0;
  for (int i = 0; i < count; ++i) {
    im.kpts_i[i] = c->kpts + static_cast<size_t>(slots[i]) * c->cap * 2;
    im.desc[i] = c->desc + static_cast<size_t>(slots[i]) * c->cap * 256;
    im.n[i] = n[i];
    im.row0[i] = rows;
    mv.slot[i] = slots[i]; mv.n[i] = n[i]; mv.row0[i] = rows;
    self_p.nq[i] = n[i]; self_p.nk[i] = n[i]; self_p.q_row0[i] = rows; self_p.k_row0[i] = rows;
    rows += round_up(n[i], 8);
    mx = n[i] > mx ? n[i] : mx;
  }
  im.count = mv.count = count;
  if (rows > c->lg_rows) {
    set_error("layer-0 cache build needs %d rows, ctx capacity %d", rows, c->lg_rows);
    return RFE_ERR_CAPACITY;
  }
  {
    ProfScope ps_(c, "lg.prepare");
    launch_lg_prepare(c->stream, im, mx, norm_h, norm_w, c->posenc_w, c->cs, c->sn, c->x, c->cat.hi, c->cat.lo);
    c->launches++;
  }
  const int lin_bn = use_bres(c, ((rows + 127) / 128) * 2) ? 128 : 64;
  if ((r = lg_self_block(c, c->layers[0], rows, self_p, count, mx, lin_bn))) return r;
  cache_move(c, mv, mx, /*to_cache=*/1);
  for (int i = 0; i < count; ++i) {
    c->cache_gen[slots[i]] = c->slot_gen[slots[i]];
    c->cache_nh[slots[i]] = norm_h;
    c->cache_nw[slots[i]] = norm_w;
  }
  return RFE_OK;
}

struct PairDesc {      // one LightGlue problem: device-resident pixel keypoints [n][2] and descriptors [n][256]
  const float* kpts0;  // fp32 keypoints (host API) ...
  const float* kpts1;
  const int* ikpts0;   // ... or int32 keypoints left on the device by SuperPoint (exactly one of the two forms is set)
  const int* ikpts1;
  const float* desc0;
  const float* desc1;
  int n0, n1;
  int rslot;
  int slot0 = -1, slot1 = -1;   // feature slots the two images came from (needed by the layer-0 cache only)
};

// Match `np` independent pairs in one pass.  All images' rows are concatenated (each image starts at a multiple of
// 8 rows), so every linear layer is ONE GEMM over all pairs; attention runs as 2*np problems of one fused launch.
int lg_run(rfe_ctx* c, const PairDesc* pairs_in, int np_in, int norm_h, int norm_w, float thresh, bool cache_slots = false) {
  cudaStream_t s = c->stream;
  int r;
  if (c->no_lg) {
    set_error("this ctx was created with RFE_FLAG_NO_MATCHER");
    return RFE_ERR_INVALID;
  }
  PairDesc pairs[kMaxPairs];
  int off0[kMaxPairs], off1[kMaxPairs];
  int np = 0, rows = 0;
  for (int i = 0; i < np_in; ++i) {
    const PairDesc& pd = pairs_in[i];
    if (pd.n0 == 0 || pd.n1 == 0) {       // the reference would run ORT on empty tensors; the answer is "no matches"
      RFE_CUDA_CHECK(cudaMemsetAsync(c->res_count + pd.rslot, 0, sizeof(int), s));
      continue;
    }
    pairs[np] = pd;
    off0[np] = rows;
    rows += round_up(pd.n0, 8);
    off1[np] = rows;
    rows += round_up(pd.n1, 8);
    ++np;
  }
  if (np == 0) return RFE_OK;
  if (rows > c->lg_rows) {
    set_error("LightGlue batch needs %d rows, ctx capacity %d", rows, c->lg_rows);
    return RFE_ERR_CAPACITY;
  }
  if (cache_slots) {   // every image's state after layer 0's self block comes out of its slot's cache: ONE launch
    LgCacheMove mv;
    memset(&mv, 0, sizeof(mv));
    int mx = 0;
    for (int i = 0; i < np; ++i) {
      mv.slot[2 * i] = pairs[i].slot0;      mv.n[2 * i] = pairs[i].n0;      mv.row0[2 * i] = off0[i];
      mv.slot[2 * i + 1] = pairs[i].slot1;  mv.n[2 * i + 1] = pairs[i].n1;  mv.row0[2 * i + 1] = off1[i];
      mx = pairs[i].n0 > mx ? pairs[i].n0 : mx;
      mx = pairs[i].n1 > mx ? pairs[i].n1 : mx;
    }
    mv.count = 2 * np;
    cache_move(c, mv, mx, /*to_cache=*/0);
  } else {   // positional encodings + residual-stream initialisation of every image: ONE launch
    LgImages im;
    memset(&im, 0, sizeof(im));
    int mx = 0;
    for (int i = 0; i < np; ++i) {
      const PairDesc& pd = pairs[i];
      im.kpts_f[2 * i] = pd.kpts0;      im.kpts_i[2 * i] = pd.ikpts0;      im.desc[2 * i] = pd.desc0;
      im.n[2 * i] = pd.n0;              im.row0[2 * i] = off0[i];
      im.kpts_f[2 * i + 1] = pd.kpts1;  im.kpts_i[2 * i + 1] = pd.ikpts1;  im.desc[2 * i + 1] = pd.desc1;
      im.n[2 * i + 1] = pd.n1;          im.row0[2 * i + 1] = off1[i];
      mx = pd.n0 > mx ? pd.n0 : mx;
      mx = pd.n1 > mx ? pd.n1 : mx;
    }
    im.count = 2 * np;
    ProfScope ps_(c, "lg.prepare");
    launch_lg_prepare(s, im, mx, norm_h, norm_w, c->posenc_w, c->cs, c->sn, c->x, c->cat.hi, c->cat.lo);
    c->launches++;
  }
  AttnParams self_p, cross_p;
  memset(&self_p, 0, sizeof(self_p));
  memset(&cross_p, 0, sizeof(cross_p));
  int max_n = 0;
  for (int i = 0; i < np; ++i) {
    const int n0 = pairs[i].n0, n1 = pairs[i].n1;
    max_n = n0 > max_n ? n0 : max_n;
    max_n = n1 > max_n ? n1 : max_n;
    self_p.nq[2 * i] = n0;      self_p.nk[2 * i] = n0;      self_p.q_row0[2 * i] = off0[i];      self_p.k_row0[2 * i] = off0[i];
    self_p.nq[2 * i + 1] = n1;  self_p.nk[2 * i + 1] = n1;  self_p.q_row0[2 * i + 1] = off1[i];  self_p.k_row0[2 * i + 1] = off1[i];
    cross_p.nq[2 * i] = n0;     cross_p.nk[2 * i] = n1;     cross_p.q_row0[2 * i] = off0[i];     cross_p.k_row0[2 * i] = off1[i];
    cross_p.nq[2 * i + 1] = n1; cross_p.nk[2 * i + 1] = n0; cross_p.q_row0[2 * i + 1] = off1[i]; cross_p.k_row0[2 * i + 1] = off0[i];
  }
  const long long hs = static_cast<long long>(rows) * 64;
  // the 256 -> 256 linears: 128-wide B-resident tiles once there are two waves of them, else 64-wide streamed tiles
  const int lin_bn = use_bres(c, ((rows + 127) / 128) * 2) ? 128 : 64;
  for (int i = 0; i < kLayers; ++i) {
    const LgLayer& L = c->layers[i];
    // ---------------- self attention ----------------
    // layer 0's self block depends on the image alone: with cached slots (rfe_lg_match_one_to_many) it was computed when
    // the cache was built and the gathered state already holds its output
    if (!(i == 0 && cache_slots) && (r = lg_self_block(c, L, rows, self_p, 2 * np, max_n, lin_bn))) return r;
    // ---------------- cross attention ----------------
    {
      Operand A{c->cat.hi, c->cat.lo, rows, 256, 512, 0, 1};
      Operand B{L.to_qk.w.hi, L.to_qk.w.lo, 256, 256, 256, 0, 1};
      UmmaParams p = default_params();
      p.bias = L.to_qk.bias;
      p.scale = kAttnScale;
      p.out_hi = c->q.hi;
      p.out_lo = c->q.lo;
      p.head_major = 1;
      p.head_stride = hs;
      if ((r = gemm_linear(c, "lg.to_qk", A, B, p, lin_bn))) return r;
    }
    {
      Operand A{c->cat.hi, c->cat.lo, rows, 256, 512, 0, 1};
      Operand B{L.to_v.w.hi, L.to_v.w.lo, 256, 256, 256, 0, 1};
      UmmaParams p = default_params();
      p.bias = L.to_v.bias;
      p.out_hi = c->vt.hi;
      p.out_lo = c->vt.lo;
      p.transpose_h = 1;
      p.ld_h = c->lg_ldv;
      if ((r = gemm_linear(c, "lg.to_v", A, B, p, lin_bn))) return r;
    }
    if ((r = attention_fused(c, "lg.attn_cross", c->q, c->q, rows, cross_p, 2 * np, max_n))) return r;
    {
      Operand A{c->attn.hi, c->attn.lo, rows, 256, 256, 0, 1};
      Operand B{L.to_out.w.hi, L.to_out.w.lo, 256, 256, 256, 0, 1};
      UmmaParams p = default_params();
      p.bias = L.to_out.bias;
      p.out_hi = c->cat.hi + 256;
      p.out_lo = c->cat.lo + 256;
      p.ld_h = 512;
      if ((r = gemm_linear(c, "lg.to_out", A, B, p, lin_bn))) return r;
    }
    if ((r = ffn_block(c, rows, L.c_ffn0, L.c_ln_w, L.c_ln_b, L.c_ffn3))) return r;
  }
  // ---------------- assignment ----------------
  {
    Operand A{c->cat.hi, c->cat.lo, rows, 256, 512, 0, 1};
    Operand B{c->final_proj.w.hi, c->final_proj.w.lo, 256, 256, 256, 0, 1};
    UmmaParams p = default_params();
    p.bias = c->final_proj.bias;
    p.scale = 0.25f;
    p.out_hi = c->md.hi;
    p.out_lo = c->md.lo;
    p.ld_h = 256;
    if ((r = gemm_linear(c, "lg.final_proj", A, B, p, lin_bn))) return r;
  }
  { ProfScope ps_(c, "lg.matchability"); launch_matchability(s, c->x, rows, c->match_w, c->match_b, c->ls); }
  c->launches++;
  LgAssign as;
  memset(&as, 0, sizeof(as));
  for (int i = 0; i < np; ++i) {
    const int n0 = pairs[i].n0, n1 = pairs[i].n1, rslot = pairs[i].rslot;
    const int ld = round_up(n1, 8);
    float* sim = c->sim + static_cast<size_t>(i) * c->cap * c->lg_ld;
    {
      Operand A{c->md.hi + static_cast<size_t>(off0[i]) * 256, c->md.lo + static_cast<size_t>(off0[i]) * 256, n0, 256, 256, 0, 1};
      Operand B{c->md.hi + static_cast<size_t>(off1[i]) * 256, c->md.lo + static_cast<size_t>(off1[i]) * 256, n1, 256, 256, 0, 1};
      UmmaParams p = default_params();
      p.out_f32 = sim;
      p.ld_f32 = ld;
      if ((r = gemm_linear(c, "lg.sim", A, B, p, 128))) return r;
    }
    as.sim[i] = sim;
    as.n0[i] = n0;
    as.n1[i] = n1;
    as.ld[i] = ld;
    as.off0[i] = off0[i];
    as.off1[i] = off1[i];
    as.matches[i] = c->res_matches + static_cast<size_t>(rslot) * c->cap * 2;
    as.mscores[i] = c->res_scores + static_cast<size_t>(rslot) * c->cap;
    as.count[i] = c->res_count + rslot;
    as.max_n0 = n0 > as.max_n0 ? n0 : as.max_n0;
    as.max_n1 = n1 > as.max_n1 ? n1 : as.max_n1;
    c->dbg_n0 = n0;
    c->dbg_n1 = n1;
    c->dbg_off0 = off0[i];
    c->dbg_off1 = off1[i];
    c->dbg_pair = i;
  }
  as.pairs = np;
  {   // dual log-softmax statistics, both arg-maxes, mutual check + compaction of ALL pairs: five launches
    ProfScope ps_(c, "lg.assign");
    // default: the five one-purpose kernels (six reads of sim, 294 us per 8 pairs).  RFE_ASSIGN=2: 32-row bands through
    // shared-memory tiles, two reads of sim -- correct (all GPU tests) but measured slower (404 us): kept for A/B only.
    static const int kAssignMode = getenv("RFE_ASSIGN") ? atoi(getenv("RFE_ASSIGN")) : 1;
    if (kAssignMode == 2)
      launch_lg_assign_banded(s, as, c->rmax, c->rlog, c->cmax, c->clog, c->ls, c->max0, c->m0, c->m1, kFilterThreshold, thresh,
                              c->S_dbg, c->part_a, c->part_b, c->lg_ld, (c->cap + 31) / 32);
    else
      launch_lg_assign(s, as, c->rmax, c->rlog, c->cmax, c->clog, c->ls, c->max0, c->m0, c->m1, kFilterThreshold, thresh,
                       c->S_dbg);
    c->launches += 5;
  }
  RFE_CUDA_CHECK(cudaGetLastError());
  return RFE_OK;
}

}  // namespace
#ifdef RFE_ENABLE_PROBES
namespace rfe {
int launch_probe_shift(cudaStream_t s, const __half* a, const __half* b, float* out);
int launch_probe_mma_rate(cudaStream_t s, float* out, int reps);
int launch_probe_ts(cudaStream_t s, const __half* a, const __half* b, float* out, int reps);
int launch_probe_softmax_role(cudaStream_t s, float* out, int reps);
}
#endif
namespace {

int check_ctx(rfe_ctx* c) {
  if (!c) {
    set_error("null ctx");
    return RFE_ERR_INVALID;
  }
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) {
    set_error("cudaSetDevice(%d): %s", c->device, cudaGetErrorString(e));
    return RFE_ERR_CUDA;
  }
  return RFE_OK;
}

// combine a split tensor into fp32 on the device (debug only)
__global__ void combine_split_kernel(const __half* hi, const __half* lo, float* out, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __half2float(hi[i]) + __half2float(lo[i]) * RFE_SPLIT_INV;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

const char* rfe_last_error(void) { return get_error(); }

int rfe_create(const rfe_config* cfg, rfe_ctx** out) {
  if (!cfg || !out) {
    set_error("rfe_create: null argument");
    return RFE_ERR_INVALID;
  }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: the rover_fe front end has no CPU fallback");
    return RFE_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  RFE_CUDA_CHECK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; this library contains sm_100a code only", cfg->device, prop.major, prop.minor);
    return RFE_ERR_NO_DEVICE;
  }
  RFE_CUDA_CHECK(cudaSetDevice(cfg->device));
  rfe_ctx* c = new rfe_ctx();
  c->device = cfg->device;
  c->num_sms = c->device_sms = prop.multiProcessorCount;
  c->max_batch = cfg->max_batch > 0 ? cfg->max_batch : 8;
  c->max_h = cfg->max_height > 0 ? cfg->max_height : 480;
  c->max_w = cfg->max_width > 0 ? cfg->max_width : 768;
  c->cap = cfg->max_keypoints > 0 ? cfg->max_keypoints : 4096;
  c->slot_gen.assign(2 * static_cast<size_t>(c->max_batch), 0);
  c->no_sp = (cfg->flags & RFE_FLAG_NO_EXTRACTOR) != 0;
  c->no_lg = (cfg->flags & RFE_FLAG_NO_MATCHER) != 0;
  if (c->no_sp) c->max_h = c->max_w = 8;      // the activation buffers shrink to nothing; the feature slots stay (rfe_sp_write_slot)
  if (c->max_h % 8 || c->max_w % 8) {
    set_error("max_height/max_width must be multiples of 8");
    delete c;
    return RFE_ERR_INVALID;
  }
  // every failure below releases what was built so far
#define C_(expr)                                                                                  \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));            \
      rfe_destroy(c);                                                                             \
      return RFE_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)
  if (cfg->stream) {
    c->stream = static_cast<cudaStream_t>(cfg->stream);
  } else {
    C_(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  C_(cudaEventCreate(&c->ev0));
  C_(cudaEventCreate(&c->ev1));
  c->prof_tag = getenv("RFE_PROF_TAG");
  if (getenv("RFE_CONV_STRIP")) c->use_strip_conv = atoi(getenv("RFE_CONV_STRIP")) != 0;
  if (getenv("RFE_FUSE_CONV1A")) c->fuse_conv1a = atoi(getenv("RFE_FUSE_CONV1A")) != 0;
  const char* path = cfg->weights_path;
  if (!path) path = getenv("ROVER_FE_WEIGHTS");
  if (!path) path = "weights/rover_fe.rfw";
  int r = load_weights(c, path);
  if (r) {
    rfe_destroy(c);
    return r;
  }
  if (nms_prepare() || nms2_prepare()) {
    set_error("cudaFuncSetAttribute(nms) failed");
    rfe_destroy(c);
    return RFE_ERR_CUDA;
  }
  // ---- SuperPoint buffers ----
  const size_t B = c->max_batch, H = c->max_h, W = c->max_w, cap = c->cap;
  const size_t full = B * H * W, half_ = full / 4, quarter = full / 16, coarse = full / 64;
#define A_(expr) if ((r = (expr))) { rfe_destroy(c); return r; }
  A_(dev_alloc(c, &c->img, full));
  if (!(c->use_strip_conv && c->fuse_conv1a)) A_(split_alloc(c, &c->a1a, full * 64));   // fused: conv1a's activation never exists
  A_(split_alloc(c, &c->a1, half_ * 64));
  A_(split_alloc(c, &c->a2a, half_ * 64));
  A_(split_alloc(c, &c->a2, quarter * 64));
  A_(split_alloc(c, &c->a3a, quarter * 128));
  A_(split_alloc(c, &c->a3, coarse * 128));
  A_(split_alloc(c, &c->a4a, coarse * 128));
  A_(split_alloc(c, &c->feat, coarse * 128));
  A_(split_alloc(c, &c->pa, coarse * 256));
  A_(split_alloc(c, &c->da, coarse * 256));
  A_(dev_alloc(c, &c->heat, full));
  A_(dev_alloc(c, &c->nmsmap, full));
  A_(dev_alloc(c, &c->dense, coarse * 256));
  A_(dev_alloc(c, &c->dense_ss, coarse * 4));
  A_(dev_alloc(c, &c->row_cnt, B * H));
  A_(dev_alloc(c, &c->row_off, B * H));
  A_(dev_alloc(c, &c->kp_counts, 2 * B));            // two feature-slot sets (see rfe_pairs_submit)
  A_(dev_alloc(c, &c->kpts, 2 * B * cap * 2));
  A_(dev_alloc(c, &c->kp_scores, 2 * B * cap));
  A_(dev_alloc(c, &c->desc, 2 * B * cap * 256));
  A_(dev_alloc(c, &c->desc_bin, 2 * B * cap * 256));
  C_(cudaMallocHost(&c->h_counts, sizeof(int) * (2 * B + 2)));
  C_(cudaMallocHost(&c->h_counts2, sizeof(int) * 2 * B));
  C_(cudaMallocHost(&c->h_mcounts, sizeof(int) * B));
  C_(cudaEventCreateWithFlags(&c->ev_counts[0], cudaEventDisableTiming));
  C_(cudaEventCreateWithFlags(&c->ev_counts[1], cudaEventDisableTiming));
  C_(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
  C_(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  C_(cudaEventCreateWithFlags(&c->ev_feat[0], cudaEventDisableTiming));
  C_(cudaEventCreateWithFlags(&c->ev_feat[1], cudaEventDisableTiming));
  // ---- LightGlue buffers ----
  c->lg_pairs = c->max_batch < kMaxPairs ? c->max_batch : kMaxPairs;   // pairs per rfe_lg_match_slots_batch
  if (c->no_lg) c->lg_pairs = 0;
  c->lg_rows = 2 * c->lg_pairs * (c->cap + 8);
  c->lg_ld = round_up(c->cap, 8);
  c->lg_ldv = round_up(c->lg_rows, 8);
  const size_t R = c->lg_rows, LD = c->lg_ld;
  if (!c->no_lg) {
  A_(dev_alloc(c, &c->in_kpts, R * 2));
  A_(dev_alloc(c, &c->in_desc, R * 256));
  A_(dev_alloc(c, &c->cs, R * 32));
  A_(dev_alloc(c, &c->sn, R * 32));
  A_(dev_alloc(c, &c->x, R * 256));
  A_(split_alloc(c, &c->cat, R * 512));
  A_(split_alloc(c, &c->q, R * 256));
  A_(split_alloc(c, &c->k, R * 256));
  A_(split_alloc(c, &c->vt, 256 * static_cast<size_t>(c->lg_ldv)));
  A_(split_alloc(c, &c->attn, R * 256));
  A_(dev_alloc(c, &c->attn_part_o, static_cast<size_t>(kAttnPartSlots) * 128 * 64));
  A_(dev_alloc(c, &c->attn_part_ml, static_cast<size_t>(kAttnPartSlots) * 2 * 128));
  A_(dev_alloc(c, &c->attn_part_cnt, static_cast<size_t>(kAttnPartSlots)));
  C_(cudaMemset(c->attn_part_cnt, 0, kAttnPartSlots * sizeof(unsigned)));
  A_(dev_alloc(c, &c->hid, R * 512));
  A_(split_alloc(c, &c->hs, R * 512));
  A_(split_alloc(c, &c->md, R * 256));
  A_(dev_alloc(c, &c->sim, static_cast<size_t>(c->lg_pairs) * cap * LD));   // one similarity matrix per pair of a batch
  A_(dev_alloc(c, &c->rmax, R));      // per-row vectors of the assignment stage, indexed like the LightGlue rows
  A_(dev_alloc(c, &c->rlog, R));
  A_(dev_alloc(c, &c->cmax, R));
  A_(dev_alloc(c, &c->clog, R));
  A_(dev_alloc(c, &c->ls, R));
  A_(dev_alloc(c, &c->max0, R));
  A_(dev_alloc(c, &c->m0, R));
  A_(dev_alloc(c, &c->m1, R));
  A_(dev_alloc(c, &c->part_a, static_cast<size_t>(c->lg_pairs) * ((cap + 31) / 32) * LD));
  A_(dev_alloc(c, &c->part_b, static_cast<size_t>(c->lg_pairs) * ((cap + 31) / 32) * LD));
  }
  A_(dev_alloc(c, &c->res_matches, B * cap * 2));
  A_(dev_alloc(c, &c->res_scores, B * cap));
  A_(dev_alloc(c, &c->res_count, B));
#undef A_
  C_(cudaDeviceSynchronize());
#undef C_
  *out = c;
  return RFE_OK;
}

void rfe_destroy(rfe_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (void* p : c->allocs) cudaFree(p);
  if (c->h_counts) cudaFreeHost(c->h_counts);
  if (c->h_counts2) cudaFreeHost(c->h_counts2);
  if (c->h_mcounts) cudaFreeHost(c->h_mcounts);
  for (int i = 0; i < 2; ++i)
    if (c->ev_counts[i]) cudaEventDestroy(c->ev_counts[i]);
  if (c->ev_done) cudaEventDestroy(c->ev_done);
  for (int i = 0; i < 2; ++i)
    if (c->ev_feat[i]) cudaEventDestroy(c->ev_feat[i]);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int rfe_sync(rfe_ctx* c) {
  int r = check_ctx(c);
  if (r) return r;
  RFE_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  return RFE_OK;
}

static int check_image_args(rfe_ctx* c, int h, int w, int stride, int batch) {
  if (h <= 0 || w <= 0 || h % 8 || w % 8 || h > c->max_h || w > c->max_w || stride < w || batch <= 0 ||
      batch > c->max_batch || static_cast<size_t>(batch) * h * stride > static_cast<size_t>(c->max_batch) * c->max_h * c->max_w) {
    set_error("invalid image arguments h=%d w=%d stride=%d batch=%d (limits %dx%d x%d; h,w multiples of 8)", h, w,
              stride, batch, c->max_h, c->max_w, c->max_batch);
    return RFE_ERR_INVALID;
  }
  return RFE_OK;
}

int rfe_sp_extract_device(rfe_ctx* c, const uint8_t* d_gray, int h, int w, int stride, int batch) {
  int r = check_ctx(c);
  if (r) return r;
  if (!d_gray) {
    set_error("null image");
    return RFE_ERR_INVALID;
  }
  if ((r = check_image_args(c, h, w, stride, batch))) return r;
  return sp_run(c, d_gray, h, w, stride, batch);
}

int rfe_sp_set_topk(rfe_ctx* c, int k) {
  int r = check_ctx(c);
  if (r) return r;
  if (k > c->cap) {
    set_error("rfe_sp_set_topk: %d exceeds the ctx keypoint capacity %d", k, c->cap);
    return RFE_ERR_INVALID;
  }
  c->topk = k > 0 ? k : 0;
  return RFE_OK;
}

int rfe_set_fast_mode(rfe_ctx* c, int on) {
  int r = check_ctx(c);
  if (r) return r;
  c->fast = on ? 1 : 0;
  return RFE_OK;
}

int rfe_set_sm_limit(rfe_ctx* c, int max_sms) {
  int r = check_ctx(c);
  if (r) return r;
  if (max_sms < 0 || max_sms > c->device_sms) {
    set_error("rfe_set_sm_limit: %d outside 0..%d", max_sms, c->device_sms);
    return RFE_ERR_INVALID;
  }
  c->num_sms = max_sms > 0 ? max_sms : c->device_sms;
  return RFE_OK;
}

int rfe_sp_read_slot(rfe_ctx* c, int slot, int32_t* kpts_xy, float* scores, float* desc, int32_t* count, int cap) {
  int r = check_ctx(c);
  if (r) return r;
  if (slot < 0 || slot >= c->last_batch) {
    set_error("slot %d out of range (last batch %d)", slot, c->last_batch);
    return RFE_ERR_INVALID;
  }
  RFE_CUDA_CHECK(cudaMemcpyAsync(c->h_counts, c->kp_counts + slot, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  RFE_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  const int n = c->h_counts[0];
  if (count) *count = n;
  int m = n < c->cap ? n : c->cap;
  if (cap < m) m = cap;
  if (m > 0) {
    if (kpts_xy) RFE_CUDA_CHECK(cudaMemcpyAsync(kpts_xy, c->kpts + static_cast<size_t>(slot) * c->cap * 2, sizeof(int) * 2 * m, cudaMemcpyDeviceToHost, c->stream));
    if (scores) RFE_CUDA_CHECK(cudaMemcpyAsync(scores, c->kp_scores + static_cast<size_t>(slot) * c->cap, sizeof(float) * m, cudaMemcpyDeviceToHost, c->stream));
    if (desc) RFE_CUDA_CHECK(cudaMemcpyAsync(desc, c->desc + static_cast<size_t>(slot) * c->cap * 256, sizeof(float) * 256 * m, cudaMemcpyDeviceToHost, c->stream));
    RFE_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
  if (n > c->cap || n > cap) {
    set_error("image slot %d has %d keypoints, capacity %d", slot, n, cap < c->cap ? cap : c->cap);
    return RFE_ERR_CAPACITY;
  }
  return RFE_OK;
}

int rfe_sp_write_slot(rfe_ctx* c, int slot, const int32_t* kpts_xy, const float* scores, const float* desc, int n) {
  int r = check_ctx(c);
  if (r) return r;
  if (slot < 0 || slot >= c->max_batch || n < 0 || n > c->cap || (n > 0 && (!kpts_xy || !desc))) {
    set_error("rfe_sp_write_slot: invalid argument (slot %d of %d, n %d, capacity %d)", slot, c->max_batch, n, c->cap);
    return RFE_ERR_INVALID;
  }
  if (n > 0) {
    const size_t so = static_cast<size_t>(slot) * c->cap;
    RFE_CUDA_CHECK(cudaMemcpyAsync(c->kpts + so * 2, kpts_xy, sizeof(int) * 2 * n, cudaMemcpyHostToDevice, c->stream));
    if (scores) RFE_CUDA_CHECK(cudaMemcpyAsync(c->kp_scores + so, scores, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    else RFE_CUDA_CHECK(cudaMemsetAsync(c->kp_scores + so, 0, sizeof(float) * n, c->stream));
    RFE_CUDA_CHECK(cudaMemcpyAsync(c->desc + so * 256, desc, sizeof(float) * 256 * n, cudaMemcpyHostToDevice, c->stream));
    // keep the slot's sign-binarised copy (rfe_sp_read_slot_bin) consistent with the uploaded descriptors
    launch_binarize(c->stream, c->desc + so * 256, n, c->desc_bin + so * 256, nullptr);
    c->launches++;
    c->bytes_h2d += static_cast<unsigned long long>(n) * (8 + (scores ? 4 : 0) + 1024);
  }
  c->slot_gen[slot]++;
  c->h_counts[0] = n;
  RFE_CUDA_CHECK(cudaMemcpyAsync(c->kp_counts + slot, c->h_counts, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  RFE_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (slot >= c->last_batch) {
    // slots between the old batch end and this one have never been written: give them zero keypoints
    if (slot > c->last_batch)
      RFE_CUDA_CHECK(cudaMemsetAsync(c->kp_counts + c->last_batch, 0, sizeof(int) * (slot - c->last_batch), c->stream));
    c->last_batch = slot + 1;
  }
  return RFE_OK;
}

int rfe_sp_read_slot_bin(rfe_ctx* c, int slot, uint8_t* bin, int32_t* count, int cap) {
  int r = check_ctx(c);
  if (r) return r;
  if (slot < 0 || slot >= c->last_batch || !bin || cap <= 0) {
    set_error("rfe_sp_read_slot_bin: invalid argument (slot %d, last batch %d)", slot, c->last_batch);
    return RFE_ERR_INVALID;
  }
  RFE_CUDA_CHECK(cudaMemcpyAsync(c->h_counts, c->kp_counts + slot, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  RFE_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  const int n = c->h_counts[0];
  if (count) *count = n;
  int m = n < c->cap ? n : c->cap;
  if (cap < m) m = cap;
  if (m > 0) {
    RFE_CUDA_CHECK(cudaMemcpyAsync(bin, c->desc_bin + static_cast<size_t>(slot) * c->cap * 256, static_cast<size_t>(m) * 256,
                                   cudaMemcpyDeviceToHost, c->stream));
    RFE_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
  if (n > c->cap || n > cap) {
    set_error("image slot %d has %d keypoints, capacity %d", slot, n, cap < c->cap ? cap : c->cap);
    return RFE_ERR_CAPACITY;
  }
  return RFE_OK;
}

// scratch device buffer of the ctx, grown on demand (host-in / host-out helpers below)
static int scratch(rfe_ctx* c, void** p, size_t* have, size_t need) {
  if (*have >= need) return RFE_OK;
  // per-frame callers (rfe_l2_best2, rfe_binarize_descriptors) see sizes that drift upward with the local map: grow
  // geometrically and release the old buffer (it may still be read by work in flight on the stream: synchronise first)
  const size_t want = need > 2 * *have ? need : 2 * *have;
  void* d = nullptr;
  if (cudaMalloc(&d, want) != cudaSuccess) {
    set_error("cudaMalloc(%zu bytes) failed", want);
    return RFE_ERR_CUDA;
  }
  if (*p) {
    cudaStreamSynchronize(c->stream);
    for (size_t i = 0; i < c->allocs.size(); ++i)
      if (c->allocs[i] == *p) {
        c->allocs.erase(c->allocs.begin() + i);
        break;
      }
    cudaFree(*p);
  }
  c->allocs.push_back(d);
  *p = d;
  *have = want;
  return RFE_OK;
}

int rfe_binarize_descriptors(rfe_ctx* c, const float* desc, int n, uint8_t* bin, uint32_t* bits) {
  int r = check_ctx(c);
  if (r) return r;
  if (n < 0 || (n > 0 && (!desc || (!bin && !bits)))) {
    set_error("rfe_binarize_descriptors: null/invalid argument");
    return RFE_ERR_INVALID;
  }
  if (n == 0) return RFE_OK;
  const size_t nb = static_cast<size_t>(n) * 256;
  if ((r = scratch(c, &c->scr[0], &c->scr_bytes[0], nb * 4))) return r;
  if ((r = scratch(c, &c->scr[1], &c->scr_bytes[1], nb + static_cast<size_t>(n) * 32))) return r;
  float* d_desc = static_cast<float*>(c->scr[0]);
  uint8_t* d_bin = static_cast<uint8_t*>(c->scr[1]);
  uint32_t* d_bits = reinterpret_cast<uint32_t*>(d_bin + nb);
  RFE_CUDA_CHECK(cudaMemcpyAsync(d_desc, desc, nb * 4, cudaMemcpyHostToDevice, c->stream));
  launch_binarize(c->stream, d_desc, n, d_bin, d_bits);
  c->launches++;
  if (bin) RFE_CUDA_CHECK(cudaMemcpyAsync(bin, d_bin, nb, cudaMemcpyDeviceToHost, c->stream));
  if (bits) RFE_CUDA_CHECK(cudaMemcpyAsync(bits, d_bits, static_cast<size_t>(n) * 32, cudaMemcpyDeviceToHost, c->stream));
  RFE_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  return RFE_OK;
}

int rfe_l2_best2(rfe_ctx* c, const float* q, int nq, const float* db, int nd, const int32_t* cand_off, const int32_t* cand_idx,
                 float init_dist, float* best_dist, int32_t* best_idx, float* second_dist, int32_t* second_idx) {
  int r = check_ctx(c);
  if (r) return r;
  if (nq < 0 || nd < 0 || (nq > 0 && (!q || !cand_off || !best_dist || !best_idx))) {
    set_error("rfe_l2_best2: null/invalid argument");
    return RFE_ERR_INVALID;
  }
  if (nq == 0) return RFE_OK;
  const int total = cand_off[nq];
  if (cand_off[0] != 0 || total < 0 || (total > 0 && (!cand_idx || !db))) {
    set_error("rfe_l2_best2: malformed candidate lists");
    return RFE_ERR_INVALID;
  }
  for (int i = 0; i < nq; ++i)
    if (cand_off[i + 1] < cand_off[i]) {
      set_error("rfe_l2_best2: candidate offsets must be non-decreasing");
      return RFE_ERR_INVALID;
    }
  for (int i = 0; i < total; ++i)
    if (cand_idx[i] < 0 || cand_idx[i] >= nd) {
      set_error("rfe_l2_best2: candidate index %d out of range [0, %d)", cand_idx[i], nd);
      return RFE_ERR_INVALID;
    }
  const size_t bq = static_cast<size_t>(nq) * 1024, bd = static_cast<size_t>(nd > 0 ? nd : 1) * 1024;
  const size_t bo = static_cast<size_t>(nq + 1) * 4, bc = static_cast<size_t>(total > 0 ? total : 1) * 4, br = static_cast<size_t>(nq) * 16;
  if ((r = scratch(c, &c->scr[0], &c->scr_bytes[0], bq + bd))) return r;
  if ((r = scratch(c, &c->scr[1], &c->scr_bytes[1], bo + bc + br + 64))) return r;
  float* d_q = static_cast<float*>(c->scr[0]);
  float* d_db = d_q + static_cast<size_t>(nq) * 256;
  int* d_off = static_cast<int*>(c->scr[1]);
  int* d_idx = d_off + (nq + 1);
  float* d_b1 = reinterpret_cast<float*>(d_idx + (total > 0 ? total : 1));
  int* d_i1 = reinterpret_cast<int*>(d_b1 + nq);
  float* d_b2 = reinterpret_cast<float*>(d_i1 + nq);
  int* d_i2 = reinterpret_cast<int*>(d_b2 + nq);
  cudaStream_t s = c->stream;
  RFE_CUDA_CHECK(cudaMemcpyAsync(d_q, q, bq, cudaMemcpyHostToDevice, s));
  if (nd > 0) RFE_CUDA_CHECK(cudaMemcpyAsync(d_db, db, static_cast<size_t>(nd) * 1024, cudaMemcpyHostToDevice, s));
  RFE_CUDA_CHECK(cudaMemcpyAsync(d_off, cand_off, bo, cudaMemcpyHostToDevice, s));
  if (total > 0) RFE_CUDA_CHECK(cudaMemcpyAsync(d_idx, cand_idx, static_cast<size_t>(total) * 4, cudaMemcpyHostToDevice, s));
  launch_l2_best2(s, d_q, nq, d_db, d_off, d_idx, init_dist, d_b1, d_i1, d_b2, d_i2);
  c->launches++;
  RFE_CUDA_CHECK(cudaMemcpyAsync(best_dist, d_b1, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, s));
  RFE_CUDA_CHECK(cudaMemcpyAsync(best_idx, d_i1, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, s));
  if (second_dist) RFE_CUDA_CHECK(cudaMemcpyAsync(second_dist, d_b2, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, s));
  if (second_idx) RFE_CUDA_CHECK(cudaMemcpyAsync(second_idx, d_i2, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, s));
  RFE_CUDA_CHECK(cudaStreamSynchronize(s));
  return RFE_OK;
}

int rfe_l2_best2_slots(rfe_ctx* c, const float* q_host, int q_slot, int db_slot, int nq, const int32_t* cand_off,
                       const int32_t* cand_idx, float init_dist, float* best_dist, int32_t* best_idx, float* second_dist,
                       int32_t* second_idx) {
  int r = check_ctx(c);
  if (r) return r;
  if (q_host) q_slot = db_slot;          // queries come from the host: only the database slot is looked at
  if (q_slot < 0 || db_slot < 0 || q_slot >= c->last_batch || db_slot >= c->last_batch || nq < 0 ||
      (nq > 0 && (!cand_off || !best_dist || !best_idx))) {
    set_error("rfe_l2_best2_slots: null/invalid argument (slots %d, %d of %d)", q_slot, db_slot, c->last_batch);
    return RFE_ERR_INVALID;
  }
  if (nq == 0) return RFE_OK;
  cudaStream_t s = c->stream;
  RFE_CUDA_CHECK(cudaMemcpyAsync(c->h_counts, c->kp_counts, sizeof(int) * c->last_batch, cudaMemcpyDeviceToHost, s));
  RFE_CUDA_CHECK(cudaStreamSynchronize(s));
  const int have_q = q_host ? nq : (c->h_counts[q_slot] < c->cap ? c->h_counts[q_slot] : c->cap);
  const int nd = c->h_counts[db_slot] < c->cap ? c->h_counts[db_slot] : c->cap;
  const int total = cand_off[nq];
  if (nq > have_q || cand_off[0] != 0 || total < 0 || (total > 0 && !cand_idx)) {
    set_error("rfe_l2_best2_slots: %d queries but slot %d holds %d descriptors, or malformed candidate lists", nq, q_slot, have_q);
    return RFE_ERR_INVALID;
  }
  for (int i = 0; i < nq; ++i)
    if (cand_off[i + 1] < cand_off[i]) {
      set_error("rfe_l2_best2_slots: candidate offsets must be non-decreasing");
      return RFE_ERR_INVALID;
    }
  for (int i = 0; i < total; ++i)
    if (cand_idx[i] < 0 || cand_idx[i] >= nd) {
      set_error("rfe_l2_best2_slots: candidate index %d out of range [0, %d)", cand_idx[i], nd);
      return RFE_ERR_INVALID;
    }
  // only the candidate lists travel: the descriptors of both sets are already in the slots
  const size_t bo = static_cast<size_t>(nq + 1) * 4, bc = static_cast<size_t>(total > 0 ? total : 1) * 4, br = static_cast<size_t>(nq) * 16;
  if ((r = scratch(c, &c->scr[1], &c->scr_bytes[1], bo + bc + br + 64))) return r;
  int* d_off = static_cast<int*>(c->scr[1]);
  int* d_idx = d_off + (nq + 1);
  float* d_b1 = reinterpret_cast<float*>(d_idx + (total > 0 ? total : 1));
  int* d_i1 = reinterpret_cast<int*>(d_b1 + nq);
  float* d_b2 = reinterpret_cast<float*>(d_i1 + nq);
  int* d_i2 = reinterpret_cast<int*>(d_b2 + nq);
  RFE_CUDA_CHECK(cudaMemcpyAsync(d_off, cand_off, bo, cudaMemcpyHostToDevice, s));
  if (total > 0) RFE_CUDA_CHECK(cudaMemcpyAsync(d_idx, cand_idx, static_cast<size_t>(total) * 4, cudaMemcpyHostToDevice, s));
  const float* d_q = c->desc + static_cast<size_t>(q_slot) * c->cap * 256;
  if (q_host) {          // e.g. MapPoint descriptors (SearchByProjection1): uploaded; the frame's descriptors stay in their slot
    if ((r = scratch(c, &c->scr[0], &c->scr_bytes[0], static_cast<size_t>(nq) * 1024))) return r;
    RFE_CUDA_CHECK(cudaMemcpyAsync(c->scr[0], q_host, static_cast<size_t>(nq) * 1024, cudaMemcpyHostToDevice, s));
    d_q = static_cast<const float*>(c->scr[0]);
  }
  launch_l2_best2(s, d_q, nq, c->desc + static_cast<size_t>(db_slot) * c->cap * 256, d_off, d_idx, init_dist, d_b1, d_i1, d_b2, d_i2);
  c->launches++;
  RFE_CUDA_CHECK(cudaMemcpyAsync(best_dist, d_b1, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, s));
  RFE_CUDA_CHECK(cudaMemcpyAsync(best_idx, d_i1, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, s));
  if (second_dist) RFE_CUDA_CHECK(cudaMemcpyAsync(second_dist, d_b2, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, s));
  if (second_idx) RFE_CUDA_CHECK(cudaMemcpyAsync(second_idx, d_i2, static_cast<size_t>(nq) * 4, cudaMemcpyDeviceToHost, s));
  RFE_CUDA_CHECK(cudaStreamSynchronize(s));
  return RFE_OK;
}

int rfe_sp_extract_u8(rfe_ctx* c, const uint8_t* gray, int h, int w, int stride, int batch, int32_t* kpts_xy,
                      float* scores, float* desc, int32_t* counts, int cap) {
  int r = check_ctx(c);
  if (r) return r;
  if (!gray || !kpts_xy || !counts || cap <= 0) {
    set_error("rfe_sp_extract_u8: null/invalid argument");
    return RFE_ERR_INVALID;
  }
  if ((r = check_image_args(c, h, w, stride, batch))) return r;
  cudaStream_t s = c->stream;
  RFE_CUDA_CHECK(cudaEventRecord(c->ev0, s));
  RFE_CUDA_CHECK(cudaMemcpyAsync(c->img, gray, static_cast<size_t>(batch) * h * stride, cudaMemcpyHostToDevice, s));
  if ((r = sp_run(c, c->img, h, w, stride, batch))) return r;
  RFE_CUDA_CHECK(cudaMemcpyAsync(c->h_counts, c->kp_counts, sizeof(int) * batch, cudaMemcpyDeviceToHost, s));
  RFE_CUDA_CHECK(cudaStreamSynchronize(s));
  int rc = RFE_OK;
  for (int b = 0; b < batch; ++b) {
    const int n = c->h_counts[b];
    counts[b] = n;
    int m = n < c->cap ? n : c->cap;
    if (cap < m) m = cap;
    if (n > c->cap || n > cap) {
      set_error("image %d has %d keypoints, capacity %d", b, n, cap < c->cap ? cap : c->cap);
      rc = RFE_ERR_CAPACITY;
    }
    if (m == 0) continue;
    RFE_CUDA_CHECK(cudaMemcpyAsync(kpts_xy + static_cast<size_t>(b) * cap * 2, c->kpts + static_cast<size_t>(b) * c->cap * 2, sizeof(int) * 2 * m, cudaMemcpyDeviceToHost, s));
    if (scores) RFE_CUDA_CHECK(cudaMemcpyAsync(scores + static_cast<size_t>(b) * cap, c->kp_scores + static_cast<size_t>(b) * c->cap, sizeof(float) * m, cudaMemcpyDeviceToHost, s));
    if (desc) RFE_CUDA_CHECK(cudaMemcpyAsync(desc + static_cast<size_t>(b) * cap * 256, c->desc + static_cast<size_t>(b) * c->cap * 256, sizeof(float) * 256 * m, cudaMemcpyDeviceToHost, s));
  }
  RFE_CUDA_CHECK(cudaEventRecord(c->ev1, s));
  RFE_CUDA_CHECK(cudaStreamSynchronize(s));
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, c->ev0, c->ev1);
  c->timer_extract_ms += ms;
  return rc;
}

int rfe_lg_match_slots_batch(rfe_ctx* c, int n_pairs, const int* slot0, const int* slot1, int norm_h, int norm_w,
                             float thresh) {
  int r = check_ctx(c);
  if (r) return r;
  if (n_pairs <= 0 || n_pairs > c->lg_pairs || !slot0 || !slot1 || norm_h <= 0 || norm_w <= 0) {
    set_error("rfe_lg_match_slots_batch: invalid argument (at most %d pairs per call)", c->lg_pairs);
    return RFE_ERR_INVALID;
  }
  for (int i = 0; i < n_pairs; ++i)
    if (slot0[i] < 0 || slot1[i] < 0 || slot0[i] >= c->last_batch || slot1[i] >= c->last_batch) {
      set_error("rfe_lg_match_slots_batch: slot out of range");
      return RFE_ERR_INVALID;
    }
  // the keypoint counts are needed on the host to lay out the rows and size the launches
  RFE_CUDA_CHECK(cudaMemcpyAsync(c->h_counts, c->kp_counts, sizeof(int) * c->last_batch, cudaMemcpyDeviceToHost, c->stream));
  RFE_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  PairDesc pd[kMaxPairs];
  for (int i = 0; i < n_pairs; ++i) {
    const int s0 = slot0[i], s1 = slot1[i];
    const int n0 = c->h_counts[s0] < c->cap ? c->h_counts[s0] : c->cap;
    const int n1 = c->h_counts[s1] < c->cap ? c->h_counts[s1] : c->cap;
    pd[i] = PairDesc{nullptr, nullptr, c->kpts + static_cast<size_t>(s0) * c->cap * 2, c->kpts + static_cast<size_t>(s1) * c->cap * 2,
                     c->desc + static_cast<size_t>(s0) * c->cap * 256, c->desc + static_cast<size_t>(s1) * c->cap * 256, n0, n1, i};
  }
  return lg_run(c, pd, n_pairs, norm_h, norm_w, thresh);
}

int rfe_lg_match_slots(rfe_ctx* c, int slot0, int slot1, int norm_h, int norm_w, float thresh, int rslot) {
  int r = check_ctx(c);
  if (r) return r;
  if (slot0 < 0 || slot1 < 0 || slot0 >= c->last_batch || slot1 >= c->last_batch || rslot < 0 || rslot >= c->max_batch) {
    set_error("rfe_lg_match_slots: slot out of range");
    return RFE_ERR_INVALID;
  }
  RFE_CUDA_CHECK(cudaMemcpyAsync(c->h_counts, c->kp_counts, sizeof(int) * c->last_batch, cudaMemcpyDeviceToHost, c->stream));
  RFE_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  const int n0 = c->h_counts[slot0] < c->cap ? c->h_counts[slot0] : c->cap;
  const int n1 = c->h_counts[slot1] < c->cap ? c->h_counts[slot1] : c->cap;
  PairDesc pd{nullptr, nullptr, c->kpts + static_cast<size_t>(slot0) * c->cap * 2, c->kpts + static_cast<size_t>(slot1) * c->cap * 2,
              c->desc + static_cast<size_t>(slot0) * c->cap * 256, c->desc + static_cast<size_t>(slot1) * c->cap * 256, n0, n1, rslot};
  return lg_run(c, &pd, 1, norm_h, norm_w, thresh);
}

int rfe_lg_match_one_to_many(rfe_ctx* c, int slot, const int* others, int n_others, int norm_h, int norm_w, float thresh) {
  int r = check_ctx(c);
  if (r) return r;
  if (n_others <= 0 || n_others > c->lg_pairs || !others || slot < 0 || slot >= c->last_batch || norm_h <= 0 || norm_w <= 0) {
    set_error("rfe_lg_match_one_to_many: invalid argument (at most %d partners per call)", c->lg_pairs);
    return RFE_ERR_INVALID;
  }
  for (int i = 0; i < n_others; ++i)
    if (others[i] < 0 || others[i] >= c->last_batch) {
      set_error("rfe_lg_match_one_to_many: slot out of range");
      return RFE_ERR_INVALID;
    }
  if ((r = cache_alloc(c))) return r;
  RFE_CUDA_CHECK(cudaMemcpyAsync(c->h_counts, c->kp_counts, sizeof(int) * c->last_batch, cudaMemcpyDeviceToHost, c->stream));
  RFE_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  auto count_of = [&](int sl) { return c->h_counts[sl] < c->cap ? c->h_counts[sl] : c->cap; };
  // cache entries that are missing or stale (slot rewritten, other normalisation size): built in ONE pass
  int need[kMaxPairs + 1], need_n[kMaxPairs + 1], n_need = 0;
  auto want = [&](int sl) {
    if (count_of(sl) == 0) return;
    if (c->cache_gen[sl] == c->slot_gen[sl] && c->cache_nh[sl] == norm_h && c->cache_nw[sl] == norm_w) {
      c->cache_hits++;
      return;
    }
    for (int k = 0; k < n_need; ++k)
      if (need[k] == sl) return;
    need[n_need] = sl;
    need_n[n_need++] = count_of(sl);
  };
  want(slot);
  for (int i = 0; i < n_others; ++i) want(others[i]);
  if (n_need) {
    c->cache_builds += n_need;
    if ((r = lg_build_cache(c, need, need_n, n_need, norm_h, norm_w))) return r;
  }
  PairDesc pd[kMaxPairs];
  for (int i = 0; i < n_others; ++i) {
    const int s0 = slot, s1 = others[i];
    pd[i] = PairDesc{nullptr, nullptr, c->kpts + static_cast<size_t>(s0) * c->cap * 2, c->kpts + static_cast<size_t>(s1) * c->cap * 2,
                     c->desc + static_cast<size_t>(s0) * c->cap * 256, c->desc + static_cast<size_t>(s1) * c->cap * 256,
                     count_of(s0), count_of(s1), i, s0, s1};
  }
  return lg_run(c, pd, n_others, norm_h, norm_w, thresh, /*cache_slots=*/true);
}

int rfe_lg_cache_stats(rfe_ctx* c, long long* hits, long long* builds) {
  if (!c) return RFE_ERR_INVALID;
  if (hits) *hits = c->cache_hits;
  if (builds) *builds = c->cache_builds;
  return RFE_OK;
}

int rfe_pairs_submit(rfe_ctx* c, const uint8_t* gray, int h, int w, int stride, int n_pairs) {
  int r = check_ctx(c);
  if (r) return r;
  if (!gray || n_pairs <= 0 || n_pairs > c->lg_pairs) {
    set_error("rfe_pairs_submit: null/invalid argument (at most %d pairs per batch)", c->lg_pairs);
    return RFE_ERR_INVALID;
  }
  if (c->n_pending >= 2) {
    set_error("rfe_pairs_submit: two batches are already waiting; call rfe_pairs_collect first");
    return RFE_ERR_INVALID;
  }
  const int B = 2 * n_pairs;
  if ((r = check_image_args(c, h, w, stride, B))) return r;
  cudaStream_t s = c->stream;
  const int set = c->next_set;
  // one image staging buffer is enough: the copy of batch k+1 is ordered after SuperPoint of batch k on the stream
  RFE_CUDA_CHECK(cudaMemcpyAsync(c->img, gray, static_cast<size_t>(B) * h * stride, cudaMemcpyHostToDevice, s));
  c->bytes_h2d += static_cast<size_t>(B) * h * stride;
  if (c->feat_pending[set]) {     // this slot set is still being copied to the host on the copy stream
    RFE_CUDA_CHECK(cudaStreamWaitEvent(s, c->ev_feat[set], 0));
    c->feat_pending[set] = false;
  }
  if ((r = sp_run(c, c->img, h, w, stride, B, set * c->max_batch))) return r;
  RFE_CUDA_CHECK(cudaMemcpyAsync(c->h_counts2 + set * c->max_batch, c->kp_counts + set * c->max_batch, sizeof(int) * B,
                                 cudaMemcpyDeviceToHost, s));
  RFE_CUDA_CHECK(cudaEventRecord(c->ev_counts[set], s));
  c->pending[c->n_pending++] = rfe_ctx::Pending{set, n_pairs, h, w};
  c->next_set ^= 1;
  return RFE_OK;
}

int rfe_pairs_collect_begin(rfe_ctx* c, float thresh, int32_t* kpts_xy, int32_t* matches, float* mscores, int cap) {
  return rfe_pairs_collect_begin_full(c, thresh, kpts_xy, nullptr, nullptr, matches, mscores, cap);
}

int rfe_pairs_collect_begin_full(rfe_ctx* c, float thresh, int32_t* kpts_xy, float* scores, float* desc, int32_t* matches,
                                 float* mscores, int cap) {
  int r = check_ctx(c);
  if (r) return r;
  if (!matches || cap <= 0) {
    set_error("rfe_pairs_collect_begin: null/invalid argument");
    return RFE_ERR_INVALID;
  }
  if (c->n_pending == 0 || c->collecting_pairs) {
    set_error("rfe_pairs_collect_begin: nothing was submitted, or the previous collect was not ended");
    return RFE_ERR_INVALID;
  }
  const rfe_ctx::Pending pd0 = c->pending[0];     // dequeued below, once the matcher has been enqueued successfully
  cudaStream_t s = c->stream;
  const int B = 2 * pd0.n_pairs, base = pd0.set * c->max_batch;
  // the keypoint counts size the LightGlue launches: wait for THIS batch's SuperPoint only -- a batch submitted after it
  // keeps the GPU busy while the host lays out and enqueues the matcher below
  RFE_CUDA_CHECK(cudaEventSynchronize(c->ev_counts[pd0.set]));
  const int* hc = c->h_counts2 + base;
  PairDesc pd[kMaxPairs];
  int rc = RFE_OK;
  for (int i = 0; i < pd0.n_pairs; ++i) {
    const int s0 = base + 2 * i, s1 = s0 + 1;
    const int n0 = hc[2 * i] < c->cap ? hc[2 * i] : c->cap;
    const int n1 = hc[2 * i + 1] < c->cap ? hc[2 * i + 1] : c->cap;
    pd[i] = PairDesc{nullptr, nullptr, c->kpts + static_cast<size_t>(s0) * c->cap * 2, c->kpts + static_cast<size_t>(s1) * c->cap * 2,
                     c->desc + static_cast<size_t>(s0) * c->cap * 256, c->desc + static_cast<size_t>(s1) * c->cap * 256, n0, n1, i};
  }
  if (scores || desc) {
    // What SPextractor::operator() hands back besides the keypoints (Frame::mDescriptors, Frame.cc:544-559): exactly n rows per
    // image, on the copy stream -- the features were complete when ev_counts fired, so these 1-2 MB per image overlap the
    // matcher that is enqueued on the main stream right below
    RFE_CUDA_CHECK(cudaStreamWaitEvent(c->copy_stream, c->ev_counts[pd0.set], 0));
    for (int b = 0; b < B; ++b) {
      int m = hc[b] < c->cap ? hc[b] : c->cap;
      if (cap < m) m = cap;
      if (m <= 0) continue;
      if (scores) {
        RFE_CUDA_CHECK(cudaMemcpyAsync(scores + static_cast<size_t>(b) * cap, c->kp_scores + static_cast<size_t>(base + b) * c->cap,
                                       sizeof(float) * m, cudaMemcpyDeviceToHost, c->copy_stream));
        c->bytes_d2h += sizeof(float) * m;
      }
      if (desc) {
        RFE_CUDA_CHECK(cudaMemcpyAsync(desc + static_cast<size_t>(b) * cap * 256, c->desc + static_cast<size_t>(base + b) * c->cap * 256,
                                       sizeof(float) * 256 * m, cudaMemcpyDeviceToHost, c->copy_stream));
        c->bytes_d2h += sizeof(float) * 256 * m;
      }
    }
    RFE_CUDA_CHECK(cudaEventRecord(c->ev_feat[pd0.set], c->copy_stream));
    c->feat_pending[pd0.set] = true;
  }
  c->collecting_set = pd0.set;
  if ((r = lg_run(c, pd, pd0.n_pairs, pd0.h, pd0.w, thresh))) return r;   // the batch stays queued: collect can be retried
  c->pending[0] = c->pending[1];
  --c->n_pending;
  // Results go back in as few copies as possible (a small device-to-host copy costs ~10 us of stream time, and a batch
  // has 3 arrays per pair): when the caller's capacity equals the ctx capacity the slot arrays are copied whole.
  const bool bulk = (cap == c->cap);
  for (int b = 0; b < B; ++b) {
    const int n = hc[b];
    c->collecting_counts[b] = n;
    int m = n < c->cap ? n : c->cap;
    if (cap < m) m = cap;
    if (n > c->cap || n > cap) {
      set_error("image %d has %d keypoints, capacity %d", b, n, cap < c->cap ? cap : c->cap);
      rc = RFE_ERR_CAPACITY;
    }
    if (kpts_xy && m > 0 && !bulk) {
      RFE_CUDA_CHECK(cudaMemcpyAsync(kpts_xy + static_cast<size_t>(b) * cap * 2, c->kpts + static_cast<size_t>(base + b) * c->cap * 2,
                                     sizeof(int) * 2 * m, cudaMemcpyDeviceToHost, s));
      c->bytes_d2h += sizeof(int) * 2 * m;
    }
  }
  if (kpts_xy && bulk) {
    RFE_CUDA_CHECK(cudaMemcpyAsync(kpts_xy, c->kpts + static_cast<size_t>(base) * c->cap * 2, sizeof(int) * 2 * c->cap * B,
                                   cudaMemcpyDeviceToHost, s));
    c->bytes_d2h += sizeof(int) * 2 * c->cap * B;
  }
  RFE_CUDA_CHECK(cudaMemcpyAsync(c->h_mcounts, c->res_count, sizeof(int) * pd0.n_pairs, cudaMemcpyDeviceToHost, s));
  c->bytes_d2h += sizeof(int) * (pd0.n_pairs + B);
  if (bulk) {
    RFE_CUDA_CHECK(cudaMemcpyAsync(matches, c->res_matches, sizeof(int) * 2 * c->cap * pd0.n_pairs, cudaMemcpyDeviceToHost, s));
    c->bytes_d2h += sizeof(int) * 2 * c->cap * pd0.n_pairs;
    if (mscores) {
      RFE_CUDA_CHECK(cudaMemcpyAsync(mscores, c->res_scores, sizeof(float) * c->cap * pd0.n_pairs, cudaMemcpyDeviceToHost, s));
      c->bytes_d2h += sizeof(float) * c->cap * pd0.n_pairs;
    }
  } else {
    // a pair has at most n0 matches: copy that upper bound now instead of synchronising once more for the exact count
    for (int i = 0; i < pd0.n_pairs; ++i) {
      int m = pd[i].n0 < cap ? pd[i].n0 : cap;
      if (pd[i].n0 == 0 || pd[i].n1 == 0) m = 0;
      if (m > 0) {
        RFE_CUDA_CHECK(cudaMemcpyAsync(matches + static_cast<size_t>(i) * cap * 2, c->res_matches + static_cast<size_t>(i) * c->cap * 2,
                                       sizeof(int) * 2 * m, cudaMemcpyDeviceToHost, s));
        c->bytes_d2h += sizeof(int) * 2 * m;
        if (mscores) {
          RFE_CUDA_CHECK(cudaMemcpyAsync(mscores + static_cast<size_t>(i) * cap, c->res_scores + static_cast<size_t>(i) * c->cap,
                                         sizeof(float) * m, cudaMemcpyDeviceToHost, s));
          c->bytes_d2h += sizeof(float) * m;
        }
      }
    }
  }
  RFE_CUDA_CHECK(cudaEventRecord(c->ev_done, s));
  c->collecting_pairs = pd0.n_pairs;
  c->collecting_rc = rc;
  c->last_batch = B;   // slot queries (rfe_sp_read_slot) refer to set 0 only; the pipelined path is self-contained
  return RFE_OK;
}

int rfe_pairs_collect_end(rfe_ctx* c, int32_t* kp_counts, int32_t* match_counts) {
  int r = check_ctx(c);
  if (r) return r;
  if (!kp_counts || !match_counts || !c->collecting_pairs) {
    set_error("rfe_pairs_collect_end: null argument or no collect in progress");
    return RFE_ERR_INVALID;
  }
  // wait for THIS batch's results only: work submitted after rfe_pairs_collect_begin keeps running behind the event
  RFE_CUDA_CHECK(cudaEventSynchronize(c->ev_done));
  if (c->feat_pending[c->collecting_set]) RFE_CUDA_CHECK(cudaEventSynchronize(c->ev_feat[c->collecting_set]));
  for (int b = 0; b < 2 * c->collecting_pairs; ++b) kp_counts[b] = c->collecting_counts[b];
  for (int i = 0; i < c->collecting_pairs; ++i) match_counts[i] = c->h_mcounts[i];
  c->collecting_pairs = 0;
  if (c->collecting_rc == RFE_ERR_CAPACITY) set_error("an image of the batch has more keypoints than the capacity");
  return c->collecting_rc;
}

int rfe_pairs_collect(rfe_ctx* c, float thresh, int32_t* kpts_xy, int32_t* kp_counts, int32_t* matches, float* mscores,
                      int32_t* match_counts, int cap) {
  if (!kp_counts || !match_counts) {
    set_error("rfe_pairs_collect: null argument");
    return RFE_ERR_INVALID;
  }
  int r = rfe_pairs_collect_begin(c, thresh, kpts_xy, matches, mscores, cap);
  if (r) return r;
  return rfe_pairs_collect_end(c, kp_counts, match_counts);
}

int rfe_match_pairs_u8(rfe_ctx* c, const uint8_t* gray, int h, int w, int stride, int n_pairs, float thresh,
                       int32_t* kpts_xy, int32_t* kp_counts, int32_t* matches, float* mscores, int32_t* match_counts,
                       int cap) {
  int r = check_ctx(c);
  if (r) return r;
  if (!gray || !matches || !match_counts || !kp_counts || cap <= 0 || n_pairs <= 0 || n_pairs > c->lg_pairs) {
    set_error("rfe_match_pairs_u8: null/invalid argument (at most %d pairs per call)", c->lg_pairs);
    return RFE_ERR_INVALID;
  }
  if (c->n_pending) {
    set_error("rfe_match_pairs_u8: batches submitted with rfe_pairs_submit are still in flight");
    return RFE_ERR_INVALID;
  }
  RFE_CUDA_CHECK(cudaEventRecord(c->ev0, c->stream));
  if ((r = rfe_pairs_submit(c, gray, h, w, stride, n_pairs))) return r;
  r = rfe_pairs_collect(c, thresh, kpts_xy, kp_counts, matches, mscores, match_counts, cap);
  cudaEventRecord(c->ev1, c->stream);
  cudaStreamSynchronize(c->stream);
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, c->ev0, c->ev1);
  c->timer_extract_ms += ms;
  return r;
}

int rfe_lg_read_result(rfe_ctx* c, int rslot, int32_t* matches, float* mscores, int* k, int cap) {
  int r = check_ctx(c);
  if (r) return r;
  if (rslot < 0 || rslot >= c->max_batch || !k) {
    set_error("rfe_lg_read_result: invalid argument");
    return RFE_ERR_INVALID;
  }
  RFE_CUDA_CHECK(cudaMemcpyAsync(c->h_counts, c->res_count + rslot, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  RFE_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  const int n = c->h_counts[0];
  *k = n;
  const int m = n < cap ? n : cap;
  if (m > 0) {
    if (matches) RFE_CUDA_CHECK(cudaMemcpyAsync(matches, c->res_matches + static_cast<size_t>(rslot) * c->cap * 2, sizeof(int) * 2 * m, cudaMemcpyDeviceToHost, c->stream));
    if (mscores) RFE_CUDA_CHECK(cudaMemcpyAsync(mscores, c->res_scores + static_cast<size_t>(rslot) * c->cap, sizeof(float) * m, cudaMemcpyDeviceToHost, c->stream));
    RFE_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  }
  if (n > cap) {
    set_error("%d matches, capacity %d", n, cap);
    return RFE_ERR_CAPACITY;
  }
  return RFE_OK;
}

static int lg_match_host(rfe_ctx* c, const char* who, const float* kpts0, int n0, const float* kpts1, int n1,
                         const float* desc0, const float* desc1, int norm_h, int norm_w, float thresh, int32_t* matches,
                         float* mscores, int* k) {
  int r = check_ctx(c);
  if (r) return r;
  if (!k || n0 < 0 || n1 < 0 || (n0 > 0 && (!kpts0 || !desc0)) || (n1 > 0 && (!kpts1 || !desc1))) {
    set_error("%s: null/invalid argument", who);
    return RFE_ERR_INVALID;
  }
  if (n0 > c->cap || n1 > c->cap) {
    set_error("%s: %d/%d keypoints exceed the ctx capacity %d", who, n0, n1, c->cap);
    return RFE_ERR_CAPACITY;
  }
  *k = 0;
  if (n0 == 0 || n1 == 0) return RFE_OK;
  cudaStream_t s = c->stream;
  const int n0p = round_up(n0, 8);
  RFE_CUDA_CHECK(cudaEventRecord(c->ev0, s));
  RFE_CUDA_CHECK(cudaMemcpyAsync(c->in_kpts, kpts0, sizeof(float) * 2 * n0, cudaMemcpyHostToDevice, s));
  RFE_CUDA_CHECK(cudaMemcpyAsync(c->in_kpts + static_cast<size_t>(n0p) * 2, kpts1, sizeof(float) * 2 * n1, cudaMemcpyHostToDevice, s));
  RFE_CUDA_CHECK(cudaMemcpyAsync(c->in_desc, desc0, sizeof(float) * 256 * n0, cudaMemcpyHostToDevice, s));
  RFE_CUDA_CHECK(cudaMemcpyAsync(c->in_desc + static_cast<size_t>(n0p) * 256, desc1, sizeof(float) * 256 * n1, cudaMemcpyHostToDevice, s));
  PairDesc pd{c->in_kpts, c->in_kpts + static_cast<size_t>(n0p) * 2, nullptr, nullptr, c->in_desc,
              c->in_desc + static_cast<size_t>(n0p) * 256, n0, n1, 0};
  if ((r = lg_run(c, &pd, 1, norm_h, norm_w, thresh))) return r;
  r = rfe_lg_read_result(c, 0, matches, mscores, k, n0);
  cudaEventRecord(c->ev1, s);
  cudaStreamSynchronize(s);
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, c->ev0, c->ev1);
  c->timer_match_ms += ms;
  return r;
}

int rfe_lg_match(rfe_ctx* c, const float* kpts0, int n0, const float* kpts1, int n1, const float* desc0,
                 const float* desc1, int norm_h, int norm_w, float thresh, int32_t* matches, float* mscores, int* k) {
  if (norm_h <= 0 || norm_w <= 0) {
    set_error("rfe_lg_match: norm_h / norm_w must be positive");
    return RFE_ERR_INVALID;
  }
  return lg_match_host(c, "rfe_lg_match", kpts0, n0, kpts1, n1, desc0, desc1, norm_h, norm_w, thresh, matches, mscores, k);
}

int rfe_lg_match_normalized(rfe_ctx* c, const float* kn0, int n0, const float* kn1, int n1, const float* desc0,
                            const float* desc1, float thresh, int32_t* matches, float* mscores, int* k) {
  return lg_match_host(c, "rfe_lg_match_normalized", kn0, n0, kn1, n1, desc0, desc1, 0, 0, thresh, matches, mscores, k);
}

void* rfe_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    set_error("cudaMallocHost(%zu bytes) failed", bytes);
    return nullptr;
  }
  return p;
}
void rfe_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}

int rfe_lg_copy_results_device(rfe_ctx* c, int n_pairs, int32_t* d_matches, float* d_mscores, int32_t* d_counts) {
  int r = check_ctx(c);
  if (r) return r;
  if (n_pairs <= 0 || n_pairs > c->max_batch || !d_matches || !d_counts) {
    set_error("rfe_lg_copy_results_device: invalid argument");
    return RFE_ERR_INVALID;
  }
  cudaStream_t s = c->stream;
  RFE_CUDA_CHECK(cudaMemcpyAsync(d_matches, c->res_matches, sizeof(int) * 2 * c->cap * n_pairs, cudaMemcpyDeviceToDevice, s));
  if (d_mscores) RFE_CUDA_CHECK(cudaMemcpyAsync(d_mscores, c->res_scores, sizeof(float) * c->cap * n_pairs, cudaMemcpyDeviceToDevice, s));
  RFE_CUDA_CHECK(cudaMemcpyAsync(d_counts, c->res_count, sizeof(int) * n_pairs, cudaMemcpyDeviceToDevice, s));
  return RFE_OK;
}

double rfe_get_timer_ms(rfe_ctx* c, const char* name) {
  if (!c || !name) return 0.0;
  if (!strcmp(name, "extractor")) return c->timer_extract_ms;
  if (!strcmp(name, "matcher")) return c->timer_match_ms;
  return 0.0;
}

long long rfe_kernel_launches(rfe_ctx* c) { return c ? c->launches : 0; }

int rfe_transfer_bytes(rfe_ctx* c, unsigned long long* h2d, unsigned long long* d2h) {
  if (!c) return RFE_ERR_INVALID;
  if (h2d) *h2d = c->bytes_h2d;
  if (d2h) *d2h = c->bytes_d2h;
  return RFE_OK;
}

int rfe_profile(rfe_ctx* c, int enable) {
  int r = check_ctx(c);
  if (r) return r;
  c->profiling = enable != 0;
  return RFE_OK;
}

int rfe_profile_select(rfe_ctx* c, const char* prefix) {
  int r = check_ctx(c);
  if (r) return r;
  c->prof_prefix = prefix ? prefix : "";
  return RFE_OK;
}

int rfe_profile_read(rfe_ctx* c, const char* prefix, double* total_ms, long long* launches, int reset) {
  int r = check_ctx(c);
  if (r) return r;
  RFE_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  double ms = 0.0;
  long long n = 0;
  const size_t pl = prefix ? strlen(prefix) : 0;
  for (auto& rec : c->prof) {
    if (pl && rec.tag.compare(0, pl, prefix) != 0) continue;
    float t = 0.0f;
    if (cudaEventElapsedTime(&t, rec.a, rec.b) == cudaSuccess) {
      ms += t;
      ++n;
    }
  }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = n;
  if (reset) {
    for (auto& rec : c->prof) c->prof_pool.push_back({rec.a, rec.b});
    c->prof.clear();
  }
  return RFE_OK;
}

int rfe_debug_read(rfe_ctx* c, const char* name, void* dst, size_t capacity, size_t* bytes) {
  int r = check_ctx(c);
  if (r) return r;
  if (!name || !bytes) {
    set_error("rfe_debug_read: null argument");
    return RFE_ERR_INVALID;
  }
  const size_t B = c->last_batch, H = c->last_h, W = c->last_w;
  const SplitBuf* sb = nullptr;
  const float* fb = nullptr;
  size_t n = 0;
  const std::string s(name);
  if (s == "sp.a1a") {
    if (!c->a1a.hi) {
      set_error("rfe_debug_read: sp.a1a is never materialised when conv1a is fused into conv1b (run with RFE_FUSE_CONV1A=0)");
      return RFE_ERR_INVALID;
    }
    sb = &c->a1a; n = B * H * W * 64;
  }
  else if (s == "sp.pool1") { sb = &c->a1; n = B * H * W / 4 * 64; }
  else if (s == "sp.a2a") { sb = &c->a2a; n = B * H * W / 4 * 64; }
  else if (s == "sp.pool2") { sb = &c->a2; n = B * H * W / 16 * 64; }
  else if (s == "sp.a3a") { sb = &c->a3a; n = B * H * W / 16 * 128; }
  else if (s == "sp.pool3") { sb = &c->a3; n = B * H * W / 64 * 128; }
  else if (s == "sp.a4a") { sb = &c->a4a; n = B * H * W / 64 * 128; }
  else if (s == "sp.feat") { sb = &c->feat; n = B * H * W / 64 * 128; }
  else if (s == "sp.pa") { sb = &c->pa; n = B * H * W / 64 * 256; }
  else if (s == "sp.da") { sb = &c->da; n = B * H * W / 64 * 256; }
  else if (s == "sp.heat") { fb = c->heat; n = B * H * W; }
  else if (s == "sp.nms") { fb = c->nmsmap; n = B * H * W; }
  else if (s == "sp.dense") {
    fb = c->dense; n = B * H * W / 64 * 256;
    if (c->dense_deferred && dst && n) {     // normalise into the debug scratch: the product normalises inside the sampler
      if (c->dbg_bytes < n * sizeof(float)) {
        if ((r = dev_alloc(c, &c->dbg, n))) return r;
        c->dbg_bytes = n * sizeof(float);
      }
      launch_dense_normalize(c->stream, c->dense, c->dense_ss, n / 256, c->dbg);
      fb = c->dbg;
    }
  }
  else if (s == "lg.x") { fb = c->x; n = static_cast<size_t>(c->dbg_off1 + c->dbg_n1) * 256; }
  else if (s == "lg.sim") { fb = c->sim + static_cast<size_t>(c->dbg_pair) * c->cap * c->lg_ld; n = static_cast<size_t>(c->dbg_n0) * round_up(c->dbg_n1, 8); }
  else if (s == "lg.S") {
    if (!c->S_dbg) {   // first request arms the capture; the NEXT match fills it
      if ((r = dev_alloc(c, &c->S_dbg, static_cast<size_t>(c->cap) * c->cap))) return r;
      *bytes = 0;
      return RFE_OK;
    }
    fb = c->S_dbg;
    n = static_cast<size_t>(c->dbg_n0) * c->dbg_n1;
  } else if (s == "lg.attn_prof") {   // 16 x u64 cycle counters of the last attention launch (first request arms it)
    if (!c->attn_prof) {
      // [0,32): role counters; [32, 32 + 3*4096): per-CTA {start ns, end ns, SM id} of the last attention launch
      if ((r = dev_alloc(c, &c->attn_prof, 32 + 3 * 4096))) return r;
      RFE_CUDA_CHECK(cudaMemsetAsync(c->attn_prof, 0, (32 + 3 * 4096) * sizeof(unsigned long long), c->stream));
      *bytes = 0;
      return RFE_OK;
    }
    *bytes = (32 + 3 * 4096) * sizeof(unsigned long long);
    if (dst) {
      RFE_CUDA_CHECK(cudaMemcpyAsync(dst, c->attn_prof, *bytes < capacity ? *bytes : capacity, cudaMemcpyDeviceToHost, c->stream));
      RFE_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    return RFE_OK;
  } else {
    set_error("rfe_debug_read: unknown tensor '%s'", name);
    return RFE_ERR_INVALID;
  }
  *bytes = n * sizeof(float);
  if (n == 0 || !dst) return RFE_OK;
  if (sb) {
    if (c->dbg_bytes < n * sizeof(float)) {
      if ((r = dev_alloc(c, &c->dbg, n))) return r;
      c->dbg_bytes = n * sizeof(float);
    }
    combine_split_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c->stream>>>(sb->hi, sb->lo, c->dbg, n);
    fb = c->dbg;
  }
  const size_t cp = *bytes < capacity ? *bytes : capacity;
  RFE_CUDA_CHECK(cudaMemcpyAsync(dst, fb, cp, cudaMemcpyDeviceToHost, c->stream));
  RFE_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  return RFE_OK;
}

int rfe_debug_gemm(rfe_ctx* c, const float* a, const float* b, const float* bias, float* d, int m, int n, int kdim) {
  int r = check_ctx(c);
  if (r) return r;
  if (!a || !b || !d || m <= 0 || n <= 0 || kdim <= 0 || kdim % 8) {
    set_error("rfe_debug_gemm: invalid argument (K must be a multiple of 8)");
    return RFE_ERR_INVALID;
  }
  std::vector<__half> ah(static_cast<size_t>(m) * kdim), al(ah.size()), bh(static_cast<size_t>(n) * kdim), bl(bh.size());
  for (size_t i = 0; i < ah.size(); ++i) split_f32(a[i], ah[i], al[i]);
  for (size_t i = 0; i < bh.size(); ++i) split_f32(b[i], bh[i], bl[i]);
  __half *dah, *dal, *dbh, *dbl;
  float *dd, *dbias = nullptr;
  RFE_CUDA_CHECK(cudaMalloc(&dah, ah.size() * 2));
  RFE_CUDA_CHECK(cudaMalloc(&dal, ah.size() * 2));
  RFE_CUDA_CHECK(cudaMalloc(&dbh, bh.size() * 2));
  RFE_CUDA_CHECK(cudaMalloc(&dbl, bh.size() * 2));
  const int ldd = round_up(n, 4);      // the TMA-store epilogue needs a 16-byte row pitch
  RFE_CUDA_CHECK(cudaMalloc(&dd, static_cast<size_t>(m) * ldd * 4));
  RFE_CUDA_CHECK(cudaMemcpy(dah, ah.data(), ah.size() * 2, cudaMemcpyHostToDevice));
  RFE_CUDA_CHECK(cudaMemcpy(dal, al.data(), al.size() * 2, cudaMemcpyHostToDevice));
  RFE_CUDA_CHECK(cudaMemcpy(dbh, bh.data(), bh.size() * 2, cudaMemcpyHostToDevice));
  RFE_CUDA_CHECK(cudaMemcpy(dbl, bl.data(), bl.size() * 2, cudaMemcpyHostToDevice));
  if (bias) {
    RFE_CUDA_CHECK(cudaMalloc(&dbias, n * 4));
    RFE_CUDA_CHECK(cudaMemcpy(dbias, bias, n * 4, cudaMemcpyHostToDevice));
  }
  Operand A{dah, dal, m, kdim, kdim, 0, 1};
  Operand Bo{dbh, dbl, n, kdim, kdim, 0, 1};
  UmmaParams p = default_params();
  p.bias = dbias;
  p.out_f32 = dd;
  p.ld_f32 = ldd;
  r = gemm_linear(c, "debug.gemm", A, Bo, p, n <= 64 ? 64 : 128);
  if (!r) {
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
      set_error("debug gemm failed: %s", cudaGetErrorString(e));
      r = RFE_ERR_CUDA;
    } else {
      cudaMemcpy2D(d, static_cast<size_t>(n) * 4, dd, static_cast<size_t>(ldd) * 4, static_cast<size_t>(n) * 4, m, cudaMemcpyDeviceToHost);
    }
  }
  cudaFree(dah); cudaFree(dal); cudaFree(dbh); cudaFree(dbl); cudaFree(dd);
  if (dbias) cudaFree(dbias);
  return r;
}


#ifdef RFE_ENABLE_PROBES
// Hardware probe 0 (see probe_kernels.cu): a [136][64], b [64][64] fp32 in (rounded to fp16), out [9][3][128][64] fp32.
int rfe_debug_probe(rfe_ctx* c, int which, const float* a, const float* b, float* out) {
  int r = check_ctx(c);
  if (r) return r;
  if ((which == 1 || which == 3) && out) {      // MMA issue-rate probe / softmax-role probe: out[16] (see probe_kernels.cu)
    float* dout1;
    RFE_CUDA_CHECK(cudaMalloc(&dout1, 16 * 4));
    RFE_CUDA_CHECK(cudaMemset(dout1, 0, 16 * 4));
    if (which == 1 ? rfe::launch_probe_mma_rate(c->stream, dout1, 512) : rfe::launch_probe_softmax_role(c->stream, dout1, 512)) {
      set_error("probe launch failed");
      return RFE_ERR_CUDA;
    }
    cudaError_t e1 = cudaStreamSynchronize(c->stream);
    if (e1 != cudaSuccess) {
      set_error("probe 1 failed: %s", cudaGetErrorString(e1));
      return RFE_ERR_CUDA;
    }
    RFE_CUDA_CHECK(cudaMemcpy(out, dout1, 16 * 4, cudaMemcpyDeviceToHost));
    cudaFree(dout1);
    return RFE_OK;
  }
  if (which == 2 && a && b && out) {   // TS-mode probe: a [128][64], b [64][64] fp32 in, out [8192 + 8] fp32
    std::vector<__half> ha2(128 * 64), hb2(64 * 64);
    for (size_t i = 0; i < ha2.size(); ++i) ha2[i] = __float2half_rn(a[i]);
    for (size_t i = 0; i < hb2.size(); ++i) hb2[i] = __float2half_rn(b[i]);
    __half *da2, *db2;
    float* do2;
    RFE_CUDA_CHECK(cudaMalloc(&da2, ha2.size() * 2));
    RFE_CUDA_CHECK(cudaMalloc(&db2, hb2.size() * 2));
    RFE_CUDA_CHECK(cudaMalloc(&do2, (8192 + 8) * 4));
    RFE_CUDA_CHECK(cudaMemset(do2, 0, (8192 + 8) * 4));
    RFE_CUDA_CHECK(cudaMemcpy(da2, ha2.data(), ha2.size() * 2, cudaMemcpyHostToDevice));
    RFE_CUDA_CHECK(cudaMemcpy(db2, hb2.data(), hb2.size() * 2, cudaMemcpyHostToDevice));
    if (rfe::launch_probe_ts(c->stream, da2, db2, do2, 512)) {
      set_error("probe launch failed");
      return RFE_ERR_CUDA;
    }
    cudaError_t e2 = cudaStreamSynchronize(c->stream);
    if (e2 != cudaSuccess) {
      set_error("probe 2 failed: %s", cudaGetErrorString(e2));
      return RFE_ERR_CUDA;
    }
    RFE_CUDA_CHECK(cudaMemcpy(out, do2, (8192 + 8) * 4, cudaMemcpyDeviceToHost));
    cudaFree(da2); cudaFree(db2); cudaFree(do2);
    return RFE_OK;
  }
  if (which != 0 || !a || !b || !out) {
    set_error("rfe_debug_probe: invalid argument");
    return RFE_ERR_INVALID;
  }
  std::vector<__half> ha(136 * 64), hb(64 * 64);
  for (size_t i = 0; i < ha.size(); ++i) ha[i] = __float2half_rn(a[i]);
  for (size_t i = 0; i < hb.size(); ++i) hb[i] = __float2half_rn(b[i]);
  __half *da, *db;
  float* dout;
  const size_t no = 9 * 3 * 128 * 64;
  RFE_CUDA_CHECK(cudaMalloc(&da, ha.size() * 2));
  RFE_CUDA_CHECK(cudaMalloc(&db, hb.size() * 2));
  RFE_CUDA_CHECK(cudaMalloc(&dout, no * 4));
  RFE_CUDA_CHECK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
  RFE_CUDA_CHECK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  if (rfe::launch_probe_shift(c->stream, da, db, dout)) {
    set_error("probe launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return RFE_ERR_CUDA;
  }
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    set_error("probe failed: %s", cudaGetErrorString(e));
    return RFE_ERR_CUDA;
  }
  RFE_CUDA_CHECK(cudaMemcpy(out, dout, no * 4, cudaMemcpyDeviceToHost));
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return RFE_OK;
}
#endif  // RFE_ENABLE_PROBES

}  // extern "C"
