// Fused multi-head attention for LightGlue's self and cross blocks, split-fp16 on tcgen05:
//     O[q] = softmax_k( Q[q] . K[k] ) V[k]          4 heads x 64 dims, Q and K pre-scaled by 64^-1/4
// (lightglue_sim.onnx /inner_attn: Mul, Mul_1, MatMul, Softmax, MatMul_1 -- e.g. layer 0 self nodes 50-54,
//  cross nodes 141-151; run by the reference at src/Matchers/lightglue_onnx.cpp:210-214).
// The N x N score matrix never leaves the SM: one CTA owns a 128-query tile of one (problem, head) and walks the
// 64-key tiles twice.
//   pass 1: row maximum from the hi*hi product alone (one MMA instead of three; softmax is invariant to the constant
//           that is subtracted, it only has to be within ~1e-3 of the true maximum to keep exp() in range).  Pass 1
//           walks 128-key tiles (N = 128 MMAs, the size at which an SS-mode MMA is math- rather than shared-memory-
//           bound) through a 7-deep ring of K_hi stages that borrows the P and V buffers pass 2 uses later;
//   pass 2: fp32-equivalent scores, P = exp(S - max) through the SFU, row sums, P handed back to the tensor core
//           through shared memory (K-major, 128-byte swizzle, double buffered) for O += P V.
// The epilogue divides by the row sum.
//
// Tensor-core work per 64-key tile, using a [hi ; lo] concatenated B operand (the hi and lo planes of K and of V^T sit
// next to each other in the stage, so one N=128 MMA computes a_hi*b_hi and a_hi*b_lo at once):
//     S:  [S_hh | S_hl] = Q_hi [K_hi;K_lo]^T (N=128) ;  S_hl += Q_lo K_hi^T (N=64)        -> S = S_hh + 2^-11 S_hl
//     O:  [O_hh | O_hl] += P_hi [V_hi;V_lo]^T (N=128);  O_hl += P_lo V_hi^T (N=64)        -> O = (O_hh + O_hl) / (256 sum E)
// P and V^T carry their two planes at ONE scale each (common.cuh, RFE_ATTN_V_SCALE): P_hi = rn16(E), P_lo = rn16(E - P_hi)
// with E = 2^11 exp(s - max) straight from the SFU (the 2^11 is added to the exponent argument), so the low part costs
// one mixed-precision FHFMA and one convert per element: no unpack, subtract, multiply chain.
//
// CTA = 640 threads: warps 0..15 softmax/epilogue, warp 16 K producer, warp 17 V producer, warp 18 issues the score
// MMAs, warp 19 (+ TMEM alloc) issues the P V MMAs.
//   * The epilogue uses all sixteen softmax warps on the one O tile (the four warps with the same w%4 share a TMEM lane
//     quarter and split the columns).  In both passes they work as two groups of eight on alternate key tiles, so one
//     group's arithmetic (max chains; SFU) overlaps the other group's TMEM-load / shared-store / fence / barrier phase
//     (measured: pass 2 1361 -> 1297 cycles per tile, pass 1 427 -> 379 cycles per 128 keys, launch 257 -> 240 us).
//   * Two MMA-issuing threads feed the one tensor pipe: while one polls an mbarrier the other keeps the queue full.
//     Ordering between the two instruction streams is carried by the mbarriers (S -> softmax -> P -> PV) alone.
//   * The single-thread roles sit in the HIGHEST warp ids on purpose: the warp scheduler arbitrates highest-warp-id-
//     first, so the threads that feed the tensor pipe are never starved by the softmax warps sharing their schedulers.
// TMEM: 3 score buffers x 128 columns + O 128 columns = 512.  The score MMAs run up to three tiles ahead of the P V MMAs
// so that the TMEM -> registers -> exp -> shared memory -> fence -> mbarrier latency of the softmax stage is hidden.
// Measured on B200 (tools/gpu_probe.py, profiles/r01_role_breakdown.txt): 16 warps drain TMEM at 412 B/clk (not a
// limit); ex2 8.2, cvt.f16x2.f32 ~5 cycles per warp-instruction per scheduler; kind::f16 rejects A = bf16 with B = fp16
// (illegal instruction), so P_lo cannot be a free bf16 truncation.
#pragma once

#include <cuda/std/type_traits>

#include "common.cuh"

namespace rfe {

constexpr int kAttnMaxProblems = 32;   // 16 pairs x 2 directions per launch
constexpr int kAttnPartSlots = 256;    // key-range parts of attn2_kernel's tail items (at most one per SM)

struct AttnParams {
  int nq[kAttnMaxProblems], nk[kAttnMaxProblems];          // per problem (blockIdx.z): query rows, key rows
  int q_row0[kAttnMaxProblems], k_row0[kAttnMaxProblems];  // row offsets inside the head-major Q / K tensors and the V^T columns
  int item_prefix[kAttnMaxProblems + 1];   // attn2_kernel: work items (4 heads x 128-query tiles) before problem z; [nprob] = total
  int nprob;
  // attn2_kernel, tail balancing: the last `items % CTAs` work items would leave most SMs idle for a whole item, so each of
  // them is cut into split_s parts along the keys (128-key granularity).  Virtual item v >= split_first is part
  // (v - split_first) % split_s of real item split_first + (v - split_first) / split_s; a part writes its un-normalised O, its
  // row maximum and row sum to part_o / part_ml (slot v - split_first); the part that finishes last merges them.
  int n_items;               // virtual work items (== item_prefix[nprob] when nothing is split)
  int split_first, split_s;
  float* part_o;             // [slots][128][64]
  float* part_ml;            // [slots][2][128]: row maximum (score units), row sum of E
  unsigned* part_cnt;        // [slots]: arrival tickets per split item (the last arriver resets its ticket to 0)
  __half* out_hi;            // split-fp16 [rows][256]
  __half* out_lo;
  unsigned long long* prof;  // optional cycle counters written by CTA prof_cta: where the MMA / softmax roles wait
  int prof_cta;              // linear CTA index (z, y, x) whose role counters are recorded (RFE_ATTN_PROF_CTA, default 0)
};

constexpr int kAttnSoftmaxWarps = 16;                             // 4 per TMEM lane quarter: 16 of a tile's 64 columns each
constexpr int kAttnThreads = 128 + 32 * kAttnSoftmaxWarps;      // + warps 16..19: K producer, V producer, idle, MMA
constexpr int kAttnWarpK = kAttnSoftmaxWarps, kAttnWarpV = kAttnSoftmaxWarps + 1, kAttnWarpMmaS = kAttnSoftmaxWarps + 2,
              kAttnWarpMma = kAttnSoftmaxWarps + 3;
constexpr int kAttnKStages = 4;
constexpr int kAttnVStages = 3;
constexpr int kAttnSBufs = 3;                                   // score buffers in TMEM: S runs two tiles ahead of P V
constexpr int kAttnKeyTile = 64;
constexpr int kAttnKVBytes = 2 * 8192;                          // one K stage (K_hi | K_lo) or one V stage (Vt_hi | Vt_lo)
constexpr int kAttnQBytes = 2 * 16384;                          // Q_hi, Q_lo (128 rows x 128 B)
constexpr int kAttnPBytes = 2 * 16384;                          // one P buffer: P_hi, P_lo (128 rows x 128 B)
constexpr int kAttnP1Stages = 5;                                // pass-1 K_hi ring: 128 keys x 128 B per stage: sP (4) + last V stage
constexpr int kAttnSmemBytes =
    kAttnQBytes + 2 * kAttnPBytes + (kAttnKStages + kAttnVStages) * kAttnKVBytes + 1024 + 4096;

#ifdef __CUDACC__

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// tensor maps: Q and K: 3-D (64, rows_total, 4 heads), box (64, 128) resp. (64, 64); V^T: 3-D (cols_total, 64, 4), box (64, 64)
// PROF = true compiles the clock64 role counters in (tools/gpu_probe.py); the production instantiation has none: the
// MMA thread's issue loop is on the critical path and every clock read costs it tens of cycles.
template <bool PROF>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_kernel(const __grid_constant__ CUtensorMap tmQ_hi, const __grid_constant__ CUtensorMap tmQ_lo,
            const __grid_constant__ CUtensorMap tmK_hi, const __grid_constant__ CUtensorMap tmK_lo,
            const __grid_constant__ CUtensorMap tmV_hi, const __grid_constant__ CUtensorMap tmV_lo, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by offsetting the __shared__ array itself (keeps the shared address space: STS/LDS, not generic)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                   // Q_hi | Q_lo
  uint8_t* sK = smem + kAttnQBytes;                     // K stages: K_hi | K_lo
  uint8_t* sP = sK + kAttnKStages * kAttnKVBytes;       // 2 x (P_hi | P_lo)
  uint8_t* sV = sP + 2 * kAttnPBytes;                   // V stages: Vt_hi | Vt_lo
  uint8_t* tail = sV + kAttnVStages * kAttnKVBytes;
  // pass 1: 5 stages of 16 KB (128 keys of K_hi): the two P buffers and the LAST V stage, so that the first two V
  // tiles (and all K stages) of pass 2 are prefetched while pass 1 is still running
  auto s1_stage = [&](int st) { return st < 4 ? sP + st * 16384 : sV + (kAttnVStages - 1) * kAttnKVBytes; };
  static_assert(2 * kAttnPBytes == 4 * 16384 && kAttnKVBytes == 16384 && kAttnP1Stages == 5, "pass-1 ring layout");
  uint64_t* q_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* k_full = q_full + 1;                        // [4]
  uint64_t* k_empty = k_full + kAttnKStages;
  uint64_t* v_full = k_empty + kAttnKStages;            // [3]
  uint64_t* v_empty = v_full + kAttnVStages;
  uint64_t* s_full = v_empty + kAttnVStages;            // [3]
  uint64_t* s_empty = s_full + kAttnSBufs;              // [3]
  uint64_t* p_full = s_empty + kAttnSBufs;              // [2]
  uint64_t* p_empty = p_full + 2;                       // [2]
  uint64_t* o_full = p_empty + 2;
  uint64_t* k1_full = o_full + 1;                       // [7]
  uint64_t* k1_empty = k1_full + kAttnP1Stages;         // [7]
  uint64_t* p1_done = k1_empty + kAttnP1Stages;         // all pass-1 MMAs have retired: sP | sV may be rewritten
  uint64_t* s1_full = p1_done + 1;                      // [3] pass-1 score buffers (all 16 softmax warps drain each)
  uint64_t* s1_empty = s1_full + kAttnSBufs;            // [3]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(s1_empty + kAttnSBufs);
  float* stat = reinterpret_cast<float*>(tail + 512);   // [4][128] partial row max, then partial row sum

  auto tick = [&]() -> long long { return PROF ? clock64() : 0; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int z = blockIdx.z;
  const int head = blockIdx.y;
  const int nq = p.nq[z], nk = p.nk[z];
  const int m0 = blockIdx.x * 128;
  if (m0 >= nq) return;                                 // uniform per CTA
  const int cta_lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  if (PROF && p.prof && threadIdx.x == 0 && cta_lin < 4096) {      // per-CTA timeline (tools/gpu_attn_timeline.py)
    unsigned long long t;
    unsigned smid;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    p.prof[32 + 3 * cta_lin] = t;
    p.prof[32 + 3 * cta_lin + 2] = smid;
  }
  const long long cta_c0 = PROF ? clock64() : 0;
  const int T = (nk + kAttnKeyTile - 1) / kAttnKeyTile;
  const int T1 = (nk + 127) / 128;                      // pass-1 tiles (128 keys)
  const int qrow = p.q_row0[z] + m0, krow = p.k_row0[z];

  if (warp == kAttnWarpK && lane == 0) {
    tma_prefetch_desc(&tmQ_hi); tma_prefetch_desc(&tmQ_lo); tma_prefetch_desc(&tmK_hi);
    tma_prefetch_desc(&tmK_lo); tma_prefetch_desc(&tmV_hi); tma_prefetch_desc(&tmV_lo);
    mbar_init(q_full, 1);
    for (int s = 0; s < kAttnKStages; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < kAttnVStages; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    // pass 2: the softmax warps work as two groups of eight on alternate key tiles (group = tile & 1 = P buffer)
    for (int s = 0; s < kAttnSBufs; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], kAttnSoftmaxWarps / 2); }
    for (int s = 0; s < kAttnSBufs; ++s) { mbar_init(&s1_full[s], 1); mbar_init(&s1_empty[s], kAttnSoftmaxWarps / 2); }
    for (int s = 0; s < 2; ++s) { mbar_init(&p_full[s], kAttnSoftmaxWarps / 2); mbar_init(&p_empty[s], 1); }
    mbar_init(o_full, 1);
    for (int s = 0; s < kAttnP1Stages; ++s) { mbar_init(&k1_full[s], 1); mbar_init(&k1_empty[s], 1); }
    mbar_init(p1_done, 1);
    fence_barrier_init();
  }
  if (warp == kAttnWarpMma) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // TMEM columns: score buffer b: [hh | hl] at b*128 ; O: [hh | hl] at 384

  if (warp == kAttnWarpK) {
    // ===== K producer: Q once; pass 1 needs K_hi only, pass 2 both planes =====================================
    if (elect_one()) {
      mbar_expect_tx(q_full, kAttnQBytes);
      tma_load_3d(sQ, &tmQ_hi, q_full, 0, qrow, head);
      tma_load_3d(sQ + 16384, &tmQ_lo, q_full, 0, qrow, head);
      for (int g = 0; g < T1; ++g) {           // pass 1: 128 keys of K_hi per stage
        const int st = g % kAttnP1Stages;
        mbar_wait(&k1_empty[st], ((g / kAttnP1Stages) & 1) ^ 1);
        uint8_t* sb = s1_stage(st);
        mbar_expect_tx(&k1_full[st], 16384);
        tma_load_3d(sb, &tmK_hi, &k1_full[st], 0, krow + g * 128, head);
        tma_load_3d(sb + 8192, &tmK_hi, &k1_full[st], 0, krow + g * 128 + 64, head);
      }
      for (int t = 0; t < T; ++t) {            // pass 2: the sK ring is untouched by pass 1, so these loads run ahead
        const int st = t % kAttnKStages;
        mbar_wait(&k_empty[st], ((t / kAttnKStages) & 1) ^ 1);
        uint8_t* sb = sK + st * kAttnKVBytes;
        mbar_expect_tx(&k_full[st], kAttnKVBytes);
        tma_load_3d(sb, &tmK_hi, &k_full[st], 0, krow + t * kAttnKeyTile, head);
        tma_load_3d(sb + 8192, &tmK_lo, &k_full[st], 0, krow + t * kAttnKeyTile, head);
      }
    }
  } else if (warp == kAttnWarpV) {
    // ===== V producer =================================================================================================
    if (elect_one()) {
      for (int t = 0; t < T; ++t) {
        const int st = t % kAttnVStages;
        if (t == kAttnVStages - 1) mbar_wait(p1_done, 0);   // the last V stage doubles as a pass-1 K_hi stage
        mbar_wait(&v_empty[st], ((t / kAttnVStages) & 1) ^ 1);
        uint8_t* sb = sV + st * kAttnKVBytes;
        mbar_expect_tx(&v_full[st], kAttnKVBytes);
        tma_load_3d(sb, &tmV_hi, &v_full[st], krow + t * kAttnKeyTile, 0, head);
        tma_load_3d(sb + 8192, &tmV_lo, &v_full[st], krow + t * kAttnKeyTile, 0, head);
      }
    }
  } else if (warp == kAttnWarpMmaS) {
    // ===== MMA issuer 1: the score products (pass 1 and pass 2) ==========================================
    // Two issuing threads feed the one tensor pipe: while this one polls a barrier (K tile landed? score buffer
    // drained?) the other keeps the pipe's queue full, and vice versa.  With a single issuer every barrier round trip
    // (~70-150 cycles, four per key tile) was a bubble in the pipe.  Ordering between the two instruction streams is
    // carried by the mbarriers alone (S -> softmax -> P -> PV), never by issue order.
    if (elect_one()) {
      constexpr uint32_t idesc64 = make_idesc_f16(128, 64);
      constexpr uint32_t idesc128 = make_idesc_f16(128, 128);
      const uint32_t q_hi = smem_u32(sQ), q_lo = q_hi + 16384;
      const bool prof = PROF && p.prof && cta_lin == p.prof_cta;
      long long w_k = 0, w_se = 0, w_k1 = 0, w_se1 = 0;
      const long long t_begin = tick();
      mbar_wait(q_full, 0);
      const long long t_q = tick();
      for (int g = 0; g < T1; ++g) {           // pass 1: S_hh of 128 keys, one N=128 MMA per k-step
        const int st = g % kAttnP1Stages, b = g % kAttnSBufs;
        long long c0 = tick();
        mbar_wait(&k1_full[st], (g / kAttnP1Stages) & 1);
        long long c1 = tick();
        mbar_wait(&s1_empty[b], ((g / kAttnSBufs) & 1) ^ 1);
        w_k1 += c1 - c0;
        w_se1 += tick() - c1;
        tc_fence_after();
        const uint32_t k_hi = smem_u32(s1_stage(st));
        const uint32_t s_base = tmem_base + b * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(s_base, make_sw128_kmajor_desc(q_hi + k * 32), make_sw128_kmajor_desc(k_hi + k * 32), idesc128, k > 0);
        umma_commit(&s1_full[b]);
        umma_commit(&k1_empty[st]);
      }
      umma_commit(p1_done);
      const long long t_p1 = tick();
      for (int t = 0; t < T; ++t) {            // pass 2: fp32-equivalent scores of key tile t (the buffer ring continues)
        const int g = T1 + t;
        const int st = t % kAttnKStages, b = g % kAttnSBufs;
        long long c0 = tick();
        mbar_wait(&k_full[st], (t / kAttnKStages) & 1);
        long long c1 = tick();
        // the first three pass-2 tiles wait for the pass-1 drain of their buffer
        if (t < kAttnSBufs && g >= kAttnSBufs) mbar_wait(&s1_empty[b], ((g - kAttnSBufs) / kAttnSBufs) & 1);
        mbar_wait(&s_empty[b], ((t / kAttnSBufs) & 1) ^ 1);
        w_k += c1 - c0;
        w_se += tick() - c1;
        tc_fence_after();
        const uint32_t k_hi = smem_u32(sK + st * kAttnKVBytes);   // K_lo follows at +8192: [K_hi;K_lo] is one N=128 operand
        const uint32_t s_base = tmem_base + b * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t dk = make_sw128_kmajor_desc(k_hi + k * 32);
          umma_f16(s_base, make_sw128_kmajor_desc(q_hi + k * 32), dk, idesc128, k > 0);        // [S_hh | S_hl]
          umma_f16(s_base + 64, make_sw128_kmajor_desc(q_lo + k * 32), dk, idesc64, 1u);       // S_hl += Q_lo K_hi^T
        }
        umma_commit(&s_full[b]);
        umma_commit(&k_empty[st]);          // the K stage is free as soon as these MMAs retire
      }
      if (prof) {
        const long long t_end = tick();
        p.prof[0] = t_q - t_begin;      // wait for Q
        p.prof[1] = t_p1 - t_q;         // pass 1 issue time
        p.prof[2] = t_end - t_p1;       // pass 2 issue time (score issuer)
        p.prof[3] = w_k;                // waiting for K tiles
        p.prof[4] = w_se;               // waiting for a free score buffer
        p.prof[8] = w_k1;               // pass 1: waiting for K tiles
        p.prof[9] = w_se1;              // pass 1: waiting for a free score buffer
      }
    }
  } else if (warp == kAttnWarpMma) {
    // ===== MMA issuer 2: O += P V ===========================================================================
    if (elect_one()) {
      constexpr uint32_t idesc64 = make_idesc_f16(128, 64);
      constexpr uint32_t idesc128 = make_idesc_f16(128, 128);
      const uint32_t o_base = tmem_base + 384;
      const bool prof = PROF && p.prof && cta_lin == p.prof_cta;
      long long w_v = 0, w_p = 0;
      for (int t = 0; t < T; ++t) {            // consumes P buffer t&1 and V stage t%3
        const int st = t % kAttnVStages, pb = t & 1;
        long long c0 = tick();
        mbar_wait(&v_full[st], (t / kAttnVStages) & 1);
        long long c1 = tick();
        mbar_wait(&p_full[pb], (t >> 1) & 1);
        w_v += c1 - c0;
        w_p += tick() - c1;
        tc_fence_after();
        const uint32_t p_hi = smem_u32(sP + pb * kAttnPBytes), p_lo = p_hi + 16384;
        const uint32_t v_hi = smem_u32(sV + st * kAttnKVBytes);       // V_lo follows at +8192
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t dv = make_sw128_kmajor_desc(v_hi + k * 32);
          umma_f16(o_base, make_sw128_kmajor_desc(p_hi + k * 32), dv, idesc128, (t > 0 || k > 0) ? 1u : 0u);
          umma_f16(o_base + 64, make_sw128_kmajor_desc(p_lo + k * 32), dv, idesc64, 1u);
        }
        umma_commit(&v_empty[st]);
        umma_commit(&p_empty[pb]);
      }
      umma_commit(o_full);
      if (prof) {
        p.prof[5] = w_v;                // waiting for V tiles
        p.prof[6] = w_p;                // waiting for P (softmax)
        p.prof[7] = T;
      }
    }
  } else if (warp < kAttnSoftmaxWarps) {
    // ===== softmax / epilogue warps =========================================================================
    const int sw = warp;                     // 0..15
    const int q = warp & 3;                  // TMEM lane quarter
    const int cq = sw >> 2;                  // pass 1 / epilogue: which quarter of the columns
    const int grp = sw >> 3;                 // pass 2: the group that owns key tiles t with (t & 1) == grp
    const int ch2 = (sw >> 2) & 1;           // pass 2: which 32-column half of the group's 64-key tile
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float NEG = -INFINITY;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr int kSmThreads = 32 * kAttnSoftmaxWarps;

    // ---- pass 1: row maximum of the hi*hi scores ----
    float mx = NEG;
    const bool sprof = PROF && p.prof && cta_lin == p.prof_cta && warp == 0 && lane == 0;
    long long sw_s1 = 0, sw_s = 0, sw_ld = 0, sw_p = 0;
    const long long st_begin = tick();
    // two groups of eight warps take alternate 128-key tiles (as in pass 2); a warp owns 32 rows x 64 key columns of its tile,
    // so one group's TMEM loads overlap the other group's max reduction and barrier round trip
    for (int g = grp; g < T1; g += 2) {
      const int b = g % kAttnSBufs;
      const long long w0 = tick();
      mbar_wait(&s1_full[b], (g / kAttnSBufs) & 1);
      sw_s1 += tick() - w0;
      tc_fence_after();
      uint32_t a0[32], a1[32];
      tmem_ld32(tlane + b * 128 + ch2 * 64, a0);
      tmem_ld32(tlane + b * 128 + ch2 * 64 + 32, a1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s1_empty[b]);
      const int c0 = g * 128 + ch2 * 64;
      float m0a = NEG, m1a = NEG, m2a = NEG, m3a = NEG;          // four independent chains
      if (c0 + 64 <= nk) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          m0a = fmaxf(m0a, __uint_as_float(a0[j]));
          m1a = fmaxf(m1a, __uint_as_float(a0[j + 1]));
          m2a = fmaxf(m2a, __uint_as_float(a1[j]));
          m3a = fmaxf(m3a, __uint_as_float(a1[j + 1]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (c0 + j < nk) m0a = fmaxf(m0a, __uint_as_float(a0[j]));
          if (c0 + 32 + j < nk) m2a = fmaxf(m2a, __uint_as_float(a1[j]));
        }
      }
      mx = fmaxf(mx, fmaxf(fmaxf(m0a, m1a), fmaxf(m2a, m3a)));
    }
    stat[cq * 128 + row] = mx;
    named_bar_sync(1, kSmThreads);
    mx = fmaxf(fmaxf(stat[row], stat[128 + row]), fmaxf(stat[256 + row], stat[384 + row]));
    named_bar_sync(1, kSmThreads);
    const float mx_l2 = mx * kLog2e;
    const long long st_p1 = tick();

    // ---- pass 2: P = exp(S - max) -> smem (K-major, 128-byte swizzle), row sum ----
    // Two groups of eight warps take alternate key tiles, so one group's SFU phase overlaps the other group's
    // TMEM-load / shared-store / fence / barrier phase (with all sixteen warps in lock step on one tile those phases
    // serialised and the tile period was the sum of both).  A warp owns 32 rows x 32 key columns of its tile.
    // Packed fp32x2 arithmetic (FFMA2 / FADD2); only the last key tile can be ragged, so the column mask lives in a
    // separate instantiation of the tile body.
    f32x2 lsum = pk2(0.0f, 0.0f);
    const f32x2 kL2 = pk2(kLog2e, kLog2e), kL2s = pk2(kLog2e * RFE_SPLIT_INV, kLog2e * RFE_SPLIT_INV);
    const f32x2 nmx = pk2(11.0f - mx_l2, 11.0f - mx_l2);        // P is produced as E = 2^11 P
    auto tile_body = [&](int t, auto masked_tag) {
      constexpr bool kMasked = decltype(masked_tag)::value;
      const int g = T1 + t, b = g % kAttnSBufs, pb = t & 1;
      const long long w0 = tick();
      mbar_wait(&s_full[b], (t / kAttnSBufs) & 1);
      const long long w1 = tick();
      tc_fence_after();
      uint32_t a0[2][16], x0[2][16];
      const uint32_t base = tlane + b * 128 + ch2 * 32;
      tmem_ld16(base, a0[0]);
      tmem_ld16(base + 64, x0[0]);
      tmem_ld16(base + 16, a0[1]);
      tmem_ld16(base + 80, x0[1]);
      tmem_ld_wait();
      sw_s += w1 - w0;
      sw_ld += tick() - w1;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[b]);
      uint32_t ph[2][8], pl[2][8];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int c0 = t * kAttnKeyTile + ch2 * 32 + hf * 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // E = 2^11 exp(s - max) through the SFU: ex2.approx(s*log2e - max*log2e + 11); relative error ~2^-22 +
          // |s-max|*2^-23, the same order as the split operand that carries P into the tensor core.  The 2^11 lives in
          // the exponent, so E - rn16(E) IS the scaled low part: no unpack / subtract / multiply chain on the FMA
          // pipe, one mixed-precision FHFMA per element instead.
          f32x2 x = fma2(pk2u(a0[hf][2 * j], a0[hf][2 * j + 1]), kL2, nmx);
          x = fma2(pk2u(x0[hf][2 * j], x0[hf][2 * j + 1]), kL2s, x);
          float x_0, x_1;
          upk2(x, x_0, x_1);
          float e0 = fast_exp2(x_0), e1 = fast_exp2(x_1);
          if (kMasked) {
            if (c0 + 2 * j >= nk) e0 = 0.0f;
            if (c0 + 2 * j + 1 >= nk) e1 = 0.0f;
          }
          lsum = add2(lsum, pk2(e0, e1));
          uint32_t hE;
          asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hE) : "f"(e1), "f"(e0));       // low half = e0
          ph[hf][j] = hE;                                                          // P_hi = rn16(E)
          float d0, d1;
          asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\t"
              "fma.rn.f32.f16 %0, l, %5, %3;\n\tfma.rn.f32.f16 %1, h, %5, %4;\n\t}"
              : "=f"(d0), "=f"(d1)
              : "r"(hE), "f"(e0), "f"(e1), "h"(static_cast<unsigned short>(0xBC00)));   // E - rn16(E), exact
          asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pl[hf][j]) : "f"(d1), "f"(d0));
        }
      }
      const long long w2 = tick();
      mbar_wait(&p_empty[pb], ((t >> 1) & 1) ^ 1);       // PV(t-2) has consumed this P buffer
      sw_p += tick() - w2;
      uint8_t* prow_hi = sP + pb * kAttnPBytes + row * 128;
      uint8_t* prow_lo = prow_hi + 16384;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const int sc = ((ch2 * 4 + ch) ^ (row & 7)) << 4;
        const int hf = ch >> 1, o = 4 * (ch & 1);
        *reinterpret_cast<uint4*>(prow_hi + sc) = make_uint4(ph[hf][o], ph[hf][o + 1], ph[hf][o + 2], ph[hf][o + 3]);
        *reinterpret_cast<uint4*>(prow_lo + sc) = make_uint4(pl[hf][o], pl[hf][o + 1], pl[hf][o + 2], pl[hf][o + 3]);
      }
      fence_proxy_async();                    // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[pb]);
    };
    const int t_full = (nk % kAttnKeyTile) ? T - 1 : T;
#pragma unroll 1
    for (int t = grp; t < t_full; t += 2) tile_body(t, cuda::std::false_type{});
    if (t_full < T && ((T - 1) & 1) == grp) tile_body(T - 1, cuda::std::true_type{});
    if (sprof) {
      p.prof[10] = tick() - st_p1;    // softmax warp 0: pass-2 loop
      p.prof[11] = sw_s;              //   waiting for scores
      p.prof[12] = sw_p;              //   waiting for a free P buffer
      p.prof[13] = st_p1 - st_begin;  // pass-1 loop (incl. the max exchange)
      p.prof[14] = sw_s1;             //   waiting for scores
      p.prof[15] = sw_ld;             // pass 2: tcgen05.ld + wait::ld
    }
    float l;
    {
      float l0, l1;
      upk2(lsum, l0, l1);
      l = l0 + l1;
    }
    stat[cq * 128 + row] = l;                // cq = 2 * grp + ch2: four partial sums per row
    named_bar_sync(1, kSmThreads);
    l = (stat[row] + stat[128 + row]) + (stat[256 + row] + stat[384 + row]);

    // ---- epilogue: O / l -> split-fp16 [rows][256] ----
    mbar_wait(o_full, 0);
    tc_fence_after();
    {
      uint32_t a0[16], x0[16];
      const uint32_t base = tlane + 384 + cq * 16;
      tmem_ld16(base, a0);
      tmem_ld16(base + 64, x0);
      tmem_ld_wait();
      __align__(16) __half oh[16];
      __align__(16) __half ol[16];
      const float inv_l = 1.0f / (RFE_ATTN_V_SCALE * l);   // l = sum E ; both operand scales cancel here
#pragma unroll
      for (int j = 0; j < 16; ++j)
        split_f32((__uint_as_float(a0[j]) + __uint_as_float(x0[j])) * inv_l, oh[j], ol[j]);
      if (m0 + row < nq) {
        const size_t o = static_cast<size_t>(qrow + row) * 256 + head * 64 + cq * 16;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          reinterpret_cast<uint4*>(p.out_hi + o)[ch] = reinterpret_cast<const uint4*>(oh)[ch];
          reinterpret_cast<uint4*>(p.out_lo + o)[ch] = reinterpret_cast<const uint4*>(ol)[ch];
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == kAttnWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (PROF && p.prof && threadIdx.x == 0 && cta_lin < 4096) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.prof[32 + 3 * cta_lin + 1] = t;
    p.prof[32 + 3 * cta_lin + 2] |= static_cast<unsigned long long>(clock64() - cta_c0) << 16;   // SM cycles of this CTA
  }
}

#endif  // __CUDACC__

}  // namespace rfe
