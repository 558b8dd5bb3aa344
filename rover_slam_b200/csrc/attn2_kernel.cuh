// Persistent form of the fused LightGlue attention kernel (attn_kernel.cuh holds the arithmetic's description:
// split-fp16 scores, pass 1 = row maximum from the hi*hi product, pass 2 = exp / P V, same-scale P and V^T planes).
//
// What changes: ONE CTA per SM walks a list of work items (problem, head, 128-query tile) and every ring -- K stages, V
// stages, score buffers in TMEM, P buffers -- keeps running ACROSS items with monotonically counted phases (no barrier is
// ever re-initialised).  The per-item head and tail of the one-item-per-CTA kernel (barrier init, TMEM allocation, the
// first K tiles' L2 latency, the wait for Q, the O epilogue with an idle tensor pipe, CTA teardown and relaunch: 8.5 K of
// 49.5 K cycles per item, profiles/r01_attn_timeline.txt) overlap with the neighbouring items instead:
//   * pass 1 (128 keys of K_hi) and pass 2 (64 keys of K_hi | K_lo) tiles both occupy one 16 KB stage of the SAME K
//     ring, so the K producer streams  [item i pass 1][item i pass 2][item i+1 pass 1] ...  without ever aliasing the P / V
//     buffers: the first K tiles of the next item land while the current item is still in its last key tiles;
//   * the score issuer starts the next item's pass 1 as soon as its Q tile has landed (Q is reloaded by the V producer the
//     moment the last score MMA of the current item retires, i.e. under the last P V products and the epilogue);
//   * TMEM is allocated once per CTA; O is handed back to the P V issuer through an o_empty barrier.
// Ring counters: n = number of score tiles (pass 1 + pass 2) issued so far by this CTA: K stage n % NK, score buffer n % NS;
// v = number of pass-2 tiles so far: V stage v % NV, P buffer v % NP.  Every role walks the same item list and derives the
// same counters.  Default instantiation (rover_fe.cu, RFE_ATTN_CFG): two softmax groups of eight warps, NK 4, NV 3, NP 2 and
// NS 2 score buffers -- the tensor pipe is ONE queue, and a score issuer that can run three tiles ahead delays the P V
// products that free the P buffers (three buffers measured 1.5 % slower).
//   * tail balancing: the items of the last, partly filled round are cut into key-range parts (AttnParams::split_*); every
//     part writes un-normalised O, row maximum and row sum, and the part that arrives last merges them in its epilogue.
// Replaces (reference): the attention MatMul / Softmax / MatMul nodes of lightglue_sim.onnx (layer 0 self: nodes 50-54,
// cross: 141-151) executed by ONNXRuntime at src/Matchers/lightglue_onnx.cpp:210-214.
#pragma once

#include "attn_kernel.cuh"

namespace rfe {

// Q 32 KB + K 4 x 16 KB + P 2 x 32 KB + V 3 x 16 KB + 1 KB alignment slack + tail (barriers 512 B, row statistics 2 x 2 KB)
constexpr int kAttn2TailBytes = 512 + 2 * 4 * 128 * 4;
constexpr int attn2_smem_bytes(int nk, int nv, int np) {
  return kAttnQBytes + np * kAttnPBytes + (nk + nv) * kAttnKVBytes + 1024 + kAttn2TailBytes;
}
constexpr int kAttn2SmemBytes = attn2_smem_bytes(kAttnKStages, kAttnVStages, 2);

#ifdef __CUDACC__

struct AttnItem {
  int z, head, m0, nq, nk, qrow, krow, T, T1;
  int part, slot;            // part >= 0: this item covers only a sub-range of the keys (AttnParams::split_*)
};
__device__ __forceinline__ AttnItem attn_decode(const AttnParams& p, int vitem) {
  int item = vitem, part = -1, slot = 0;
  if (vitem >= p.split_first) {
    slot = vitem - p.split_first;
    item = p.split_first + slot / p.split_s;
    part = slot - (item - p.split_first) * p.split_s;
  }
  int z = 0;
  while (z + 1 < p.nprob && item >= p.item_prefix[z + 1]) ++z;
  AttnItem a;
  a.z = z;
  a.nq = p.nq[z];
  a.nk = p.nk[z];
  const int local = item - p.item_prefix[z];
  const int qtiles = (a.nq + 127) >> 7;
  a.head = local / qtiles;
  a.m0 = (local - a.head * qtiles) * 128;
  a.qrow = p.q_row0[z] + a.m0;
  a.krow = p.k_row0[z];
  a.part = part;
  a.slot = slot;
  if (part >= 0) {           // keys [part * per, min(nk, (part + 1) * per)), per a multiple of 128
    const int per = ((((a.nk + 127) >> 7) + p.split_s - 1) / p.split_s) * 128;
    const int kb = part * per;
    a.krow += kb;
    a.nk = (a.nk - kb < per) ? a.nk - kb : per;
  }
  a.T = (a.nk + kAttnKeyTile - 1) / kAttnKeyTile;
  a.T1 = (a.nk + 127) >> 7;
  return a;
}

// Merges the key-range parts of a split work item: O = sum_c w_c O_c / (256 sum_c w_c l_c), w_c = exp(max_c - max).  Called by
// the softmax warps of the CTA whose part arrived LAST (thread = (row, 16 output columns) like the epilogue); the parts are
// always summed in index order from the scratch buffers, so the result does not depend on which part that was.
__device__ __forceinline__ void attn2_merge_parts(const AttnParams& p, const AttnItem& a, int row, int cq) {
  const int slot0 = a.slot - a.part;
  float M = -INFINITY;
  for (int c = 0; c < p.split_s; ++c) M = fmaxf(M, __ldcg(p.part_ml + (static_cast<size_t>(slot0 + c) * 2) * 128 + row));
  float L = 0.0f, acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
  for (int c = 0; c < p.split_s; ++c) {
    const size_t s = slot0 + c;
    const float w = expf(__ldcg(p.part_ml + (s * 2) * 128 + row) - M);
    L = fmaf(w, __ldcg(p.part_ml + (s * 2 + 1) * 128 + row), L);
    const float4* po = reinterpret_cast<const float4*>(p.part_o + (s * 128 + row) * 64 + cq * 16);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 v = __ldcg(po + j);
      acc[4 * j] = fmaf(w, v.x, acc[4 * j]);
      acc[4 * j + 1] = fmaf(w, v.y, acc[4 * j + 1]);
      acc[4 * j + 2] = fmaf(w, v.z, acc[4 * j + 2]);
      acc[4 * j + 3] = fmaf(w, v.w, acc[4 * j + 3]);
    }
  }
  __align__(16) __half oh[16];
  __align__(16) __half ol[16];
  const float inv_l = 1.0f / (RFE_ATTN_V_SCALE * L);
  const bool live = a.m0 + row < a.nq;
#pragma unroll
  for (int j = 0; j < 16; ++j) split_f32(live ? acc[j] * inv_l : 0.0f, oh[j], ol[j]);
  if (a.m0 + row < ((a.nq + 7) & ~7)) {
    const size_t o = static_cast<size_t>(a.qrow + row) * 256 + a.head * 64 + cq * 16;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      reinterpret_cast<uint4*>(p.out_hi + o)[ch] = reinterpret_cast<const uint4*>(oh)[ch];
      reinterpret_cast<uint4*>(p.out_lo + o)[ch] = reinterpret_cast<const uint4*>(ol)[ch];
    }
  }
}

// G softmax groups (2 x 8 warps or 4 x 4 warps) on key tiles t = grp (mod G); NK / NV / NP / NS: K stages, V stages, P buffers,
// score buffers in TMEM.
template <bool PROF, int G = 2, int NK = kAttnKStages, int NV = kAttnVStages, int NP = 2, int NS = kAttnSBufs>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn2_kernel(const __grid_constant__ CUtensorMap tmQ_hi, const __grid_constant__ CUtensorMap tmQ_lo,
             const __grid_constant__ CUtensorMap tmK_hi, const __grid_constant__ CUtensorMap tmK_lo,
             const __grid_constant__ CUtensorMap tmV_hi, const __grid_constant__ CUtensorMap tmV_lo, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                   // Q_hi | Q_lo
  uint8_t* sK = smem + kAttnQBytes;                     // K stages: pass 1 K_hi(128 keys), pass 2 K_hi | K_lo (64 keys)
  uint8_t* sP = sK + NK * kAttnKVBytes;                 // NP x (P_hi | P_lo)
  uint8_t* sV = sP + NP * kAttnPBytes;                  // V stages: Vt_hi | Vt_lo
  uint8_t* tail = sV + NV * kAttnKVBytes;
  constexpr int WG = kAttnSoftmaxWarps / G;             // warps per softmax group
  uint64_t* q_full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* q_empty = q_full + 1;
  uint64_t* k_full = q_empty + 1;                       // [NK]
  uint64_t* k_empty = k_full + NK;
  uint64_t* v_full = k_empty + NK;                      // [NV]
  uint64_t* v_empty = v_full + NV;
  uint64_t* s_full = v_empty + NV;                      // [3]
  uint64_t* s_empty = s_full + NS;              // [3]
  uint64_t* p_full = s_empty + NS;              // [NP]
  uint64_t* p_empty = p_full + NP;                      // [NP]
  uint64_t* o_full = p_empty + NP;
  uint64_t* o_empty = o_full + 1;
  // G = 4 only: score-ready barriers per (group, parity of the group's own tile count).  With four groups a group may be more
  // than one phase away from a barrier shared by all groups, and a parity wait cannot tell phase k from phase k - 2; a group's
  // own two barriers alternate over its own consecutive tiles, so it is never more than one phase behind them.
  uint64_t* sfg = o_empty + 1;                          // [4][2]
  // ... and P-buffer permissions the same way: when the P V issuer has issued tile u it frees buffer u % NP for tile u + NP
  // and commits to the barrier of THAT tile's group (ordinal = pass-2 tiles of that group so far).
  uint64_t* peg = sfg + 8;                              // [4][2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(peg + 8);
  uint32_t* last_part = tmem_ptr_smem + 1;               // 1: this CTA's part of a split item arrived last and merges
  float* stat = reinterpret_cast<float*>(tail + 512);   // [2 (item parity)][4][128] partial row max, then partial row sum
  static_assert((4 + 16 + 2 * (NK + NV + NS + NP)) * 8 + 8 <= 512 && attn2_smem_bytes(NK, NV, NP) <= 227 * 1024,
                "tail region / shared-memory budget");
  static_assert(G == 2 || G == 4, "softmax groups");

  auto tick = [&]() -> long long { return PROF ? clock64() : 0; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool prof_cta = PROF && p.prof && static_cast<int>(blockIdx.x) == p.prof_cta;
  const int n_items = p.n_items;
  const long long cta_c0 = PROF ? clock64() : 0;
  if (PROF && p.prof && threadIdx.x == 0 && blockIdx.x < 4096) {
    unsigned long long t;
    unsigned smid;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    p.prof[32 + 3 * blockIdx.x] = t;
    p.prof[32 + 3 * blockIdx.x + 2] = smid;
  }

  if (warp == kAttnWarpK && lane == 0) {
    tma_prefetch_desc(&tmQ_hi); tma_prefetch_desc(&tmQ_lo); tma_prefetch_desc(&tmK_hi);
    tma_prefetch_desc(&tmK_lo); tma_prefetch_desc(&tmV_hi); tma_prefetch_desc(&tmV_lo);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < NK; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < NV; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    // both passes: the softmax warps work as G groups of WG warps on key tiles t = grp (mod G)
    for (int s = 0; s < NS; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], WG); }
    for (int s = 0; s < NP; ++s) { mbar_init(&p_full[s], WG); mbar_init(&p_empty[s], 1); }
    for (int s = 0; s < 8; ++s) { mbar_init(&sfg[s], 1); mbar_init(&peg[s], 1); }
    if (G == 4)                                         // the NP buffers are free at the start: ordinal 0 of groups 0 .. NP-1
      for (int s = 0; s < NP; ++s) mbar_arrive(&peg[s * 2]);
    mbar_init(o_full, 1);
    mbar_init(o_empty, kAttnSoftmaxWarps);
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
    // ===== K producer: one 16 KB stage per score tile, pass 1 and pass 2 alike, running across items ====================
    if (elect_one()) {
      uint32_t n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const AttnItem a = attn_decode(p, item);
        for (int g = 0; g < a.T1; ++g, ++n) {        // pass 1: 128 keys of K_hi
          const int st = n % NK;
          mbar_wait(&k_empty[st], ((n / NK) & 1) ^ 1);
          uint8_t* sb = sK + st * kAttnKVBytes;
          mbar_expect_tx(&k_full[st], kAttnKVBytes);
          tma_load_3d(sb, &tmK_hi, &k_full[st], 0, a.krow + g * 128, a.head);
          tma_load_3d(sb + 8192, &tmK_hi, &k_full[st], 0, a.krow + g * 128 + 64, a.head);
        }
        for (int t = 0; t < a.T; ++t, ++n) {         // pass 2: 64 keys, both planes
          const int st = n % NK;
          mbar_wait(&k_empty[st], ((n / NK) & 1) ^ 1);
          uint8_t* sb = sK + st * kAttnKVBytes;
          mbar_expect_tx(&k_full[st], kAttnKVBytes);
          tma_load_3d(sb, &tmK_hi, &k_full[st], 0, a.krow + t * kAttnKeyTile, a.head);
          tma_load_3d(sb + 8192, &tmK_lo, &k_full[st], 0, a.krow + t * kAttnKeyTile, a.head);
        }
      }
    }
  } else if (warp == kAttnWarpV) {
    // ===== Q + V producer: Q of item i+1 is fetched the moment the last score MMA of item i has retired =================
    if (elect_one()) {
      uint32_t v = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const AttnItem a = attn_decode(p, item);
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_expect_tx(q_full, kAttnQBytes);
        tma_load_3d(sQ, &tmQ_hi, q_full, 0, a.qrow, a.head);
        tma_load_3d(sQ + 16384, &tmQ_lo, q_full, 0, a.qrow, a.head);
        for (int t = 0; t < a.T; ++t, ++v) {
          const int st = v % NV;
          mbar_wait(&v_empty[st], ((v / NV) & 1) ^ 1);
          uint8_t* sb = sV + st * kAttnKVBytes;
          mbar_expect_tx(&v_full[st], kAttnKVBytes);
          tma_load_3d(sb, &tmV_hi, &v_full[st], a.krow + t * kAttnKeyTile, 0, a.head);
          tma_load_3d(sb + 8192, &tmV_lo, &v_full[st], a.krow + t * kAttnKeyTile, 0, a.head);
        }
      }
    }
  } else if (warp == kAttnWarpMmaS) {
    // ===== MMA issuer 1: the score products of both passes ===============================================================
    if (elect_one()) {
      constexpr uint32_t idesc64 = make_idesc_f16(128, 64);
      constexpr uint32_t idesc128 = make_idesc_f16(128, 128);
      const uint32_t q_hi = smem_u32(sQ), q_lo = q_hi + 16384;
      long long w_q = 0, w_k = 0, w_se = 0, w_k1 = 0, w_se1 = 0, t_p1 = 0, t_p2 = 0;
      uint32_t n = 0, it = 0;
      uint32_t gc = 0;                               // G = 4: tiles handed to each group so far, four 8-bit counters
      auto score_ready = [&](int g, int b) -> uint64_t* {
        if constexpr (G == 4) {
          const uint32_t sh = 8u * static_cast<uint32_t>(g), cg = (gc >> sh) & 0xffu;
          gc = (gc & ~(0xffu << sh)) | (((cg + 1u) & 0xffu) << sh);
          return &sfg[g * 2 + (cg & 1u)];
        } else {
          return &s_full[b];
        }
      };
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const AttnItem a = attn_decode(p, item);
        const long long t0 = tick();
        mbar_wait(q_full, it & 1);
        const long long t1 = tick();
        w_q += t1 - t0;
        for (int g = 0; g < a.T1; ++g, ++n) {        // pass 1: S_hh of 128 keys, one N=128 MMA per k-step
          const int st = n % NK, b = n % NS;
          const long long c0 = tick();
          mbar_wait(&k_full[st], (n / NK) & 1);
          const long long c1 = tick();
          mbar_wait(&s_empty[b], ((n / NS) & 1) ^ 1);
          w_k1 += c1 - c0;
          w_se1 += tick() - c1;
          tc_fence_after();
          const uint32_t k_hi = smem_u32(sK + st * kAttnKVBytes);
          const uint32_t s_base = tmem_base + b * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(s_base, make_sw128_kmajor_desc(q_hi + k * 32), make_sw128_kmajor_desc(k_hi + k * 32), idesc128, k > 0);
          umma_commit(score_ready(g & 3, b));
          umma_commit(&k_empty[st]);
        }
        const long long t2 = tick();
        t_p1 += t2 - t1;
        for (int t = 0; t < a.T; ++t, ++n) {         // pass 2: fp32-equivalent scores of a 64-key tile
          const int st = n % NK, b = n % NS;
          const long long c0 = tick();
          mbar_wait(&k_full[st], (n / NK) & 1);
          const long long c1 = tick();
          mbar_wait(&s_empty[b], ((n / NS) & 1) ^ 1);
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
          umma_commit(score_ready(t & 3, b));
          umma_commit(&k_empty[st]);
        }
        umma_commit(q_empty);                        // the Q tile may be replaced once these MMAs have retired
        t_p2 += tick() - t2;
      }
      if (prof_cta) {
        p.prof[0] = w_q;        // waiting for Q (all items)
        p.prof[1] = t_p1;       // pass 1 issue loops
        p.prof[2] = t_p2;       // pass 2 issue loops
        p.prof[3] = w_k;        // pass 2: waiting for K tiles
        p.prof[4] = w_se;       // pass 2: waiting for a free score buffer
        p.prof[8] = w_k1;       // pass 1: waiting for K tiles
        p.prof[9] = w_se1;      // pass 1: waiting for a free score buffer
      }
    }
  } else if (warp == kAttnWarpMma) {
    // ===== MMA issuer 2: O += P V =========================================================================================
    if (elect_one()) {
      constexpr uint32_t idesc64 = make_idesc_f16(128, 64);
      constexpr uint32_t idesc128 = make_idesc_f16(128, 128);
      const uint32_t o_base = tmem_base + 384;
      long long w_v = 0, w_p = 0, w_o = 0;
      uint32_t v = 0, it = 0, tiles = 0;
      uint32_t pc = 0;                               // G = 4: P-buffer permissions handed to each group, four 8-bit counters
      for (int s = 0; s < NP; ++s) pc |= 1u << (8 * s);
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const AttnItem a = attn_decode(p, item);
        for (int t = 0; t < a.T; ++t, ++v) {
          const int st = v % NV, pb = v % NP;
          const long long c0 = tick();
          mbar_wait(&v_full[st], (v / NV) & 1);
          const long long c1 = tick();
          mbar_wait(&p_full[pb], (v / NP) & 1);
          const long long c2 = tick();
          if (t == 0) mbar_wait(o_empty, (it & 1) ^ 1);     // the previous item's epilogue has read O out of TMEM
          w_v += c1 - c0;
          w_p += c2 - c1;
          w_o += tick() - c2;
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
          if constexpr (G == 4) {
            const int tn = t + NP;                   // the next user of this buffer: same item, or tile tn - T of the next one
            const uint32_t g2 = static_cast<uint32_t>(tn < a.T ? tn : tn - a.T) & 3u;
            const uint32_t sh = 8u * g2, ord = (pc >> sh) & 0xffu;
            pc = (pc & ~(0xffu << sh)) | (((ord + 1u) & 0xffu) << sh);
            umma_commit(&peg[g2 * 2 + (ord & 1u)]);
          } else {
            umma_commit(&p_empty[pb]);
          }
        }
        umma_commit(o_full);
        tiles += a.T;
      }
      if (prof_cta) {
        p.prof[5] = w_v;        // waiting for V tiles
        p.prof[6] = w_p;        // waiting for P (softmax)
        p.prof[7] = tiles;      // pass-2 key tiles of this CTA
        p.prof[16] = w_o;       // waiting for the O hand-back
        p.prof[17] = it;        // items of this CTA
      }
    }
  } else if (warp < kAttnSoftmaxWarps) {
    // ===== softmax / epilogue warps =======================================================================================
    const int sw = warp;                     // 0..15
    const int q = warp & 3;                  // TMEM lane quarter
    const int cq = sw >> 2;                  // pass 1 / epilogue: which quarter of the columns
    const int grp = sw / WG;                 // the group that owns key tiles t with t % G == grp
    // G = 2: a warp owns 32 of a tile's 64 columns (half ch2 = bit 2 of the warp index); G = 4: all 64, as two sequential halves
    constexpr int NCH = (G == 2) ? 1 : 2;
    const int ch_first = (G == 2) ? ((sw >> 2) & 1) : 0;
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float NEG = -INFINITY;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr int kSmThreads = 32 * kAttnSoftmaxWarps;
    const bool sprof = prof_cta && warp == 0 && lane == 0;
    long long sw_s1 = 0, sw_s = 0, sw_p = 0, sw_o = 0, t_sm1 = 0, t_sm2 = 0, t_epi = 0;
    long long ph_ld = 0, ph_math = 0, ph_sts = 0, ph_fence = 0;     // pass-2 phases of one warp (PROF build)
    uint32_t n_base = 0, v_base = 0, it = 0;
    uint32_t mycnt = 0, myp = 0;                     // G = 4: tiles / pass-2 tiles this group has taken so far
    auto wait_scores = [&](int b, uint32_t n) {
      if constexpr (G == 4) {
        mbar_wait(&sfg[grp * 2 + (mycnt & 1u)], (mycnt >> 1) & 1u);
        ++mycnt;
      } else {
        mbar_wait(&s_full[b], (n / NS) & 1);
      }
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const AttnItem a = attn_decode(p, item);
      const int nk = a.nk, T = a.T, T1 = a.T1;
      float* st_buf = stat + (it & 1) * 512;
      const long long st_begin = tick();

      // ---- pass 1: row maximum of the hi*hi scores (two groups on alternate 128-key tiles, 64 columns per warp) ----
      float mx = NEG;
      for (int g = grp; g < T1; g += G) {
        const uint32_t n = n_base + g;
        const int b = n % NS;
        const long long w0 = tick();
        wait_scores(b, n);
        sw_s1 += tick() - w0;
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < NCH; ++c) {              // 64 of the tile's 128 columns per step
          const int ch2 = ch_first + c;
          uint32_t a0[32], a1[32];
          tmem_ld32(tlane + b * 128 + ch2 * 64, a0);
          tmem_ld32(tlane + b * 128 + ch2 * 64 + 32, a1);
          tmem_ld_wait();
          if (c == NCH - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[b]);
          }
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
      }
      st_buf[cq * 128 + row] = mx;
      named_bar_sync(1, kSmThreads);
      mx = fmaxf(fmaxf(st_buf[row], st_buf[128 + row]), fmaxf(st_buf[256 + row], st_buf[384 + row]));
      named_bar_sync(1, kSmThreads);
      const float mx_l2 = mx * kLog2e;
      const long long st_p1 = tick();
      t_sm1 += st_p1 - st_begin;

      // ---- pass 2: P = exp(S - max) -> smem (K-major, 128-byte swizzle), row sum ----
      f32x2 lsum = pk2(0.0f, 0.0f);
      const f32x2 kL2 = pk2(kLog2e, kLog2e), kL2s = pk2(kLog2e * RFE_SPLIT_INV, kLog2e * RFE_SPLIT_INV);
      const f32x2 nmx = pk2(11.0f - mx_l2, 11.0f - mx_l2);        // P is produced as E = 2^11 P
      auto tile_body = [&](int t, auto masked_tag) {
        constexpr bool kMasked = decltype(masked_tag)::value;
        const uint32_t n = n_base + T1 + t, v = v_base + t;
        const int b = n % NS, pb = v % NP;
        const long long w0 = tick();
        wait_scores(b, n);
        const long long w0b = tick();
        sw_s += w0b - w0;
        tc_fence_after();
        uint8_t* prow_hi = sP + pb * kAttnPBytes + row * 128;
        uint8_t* prow_lo = prow_hi + 16384;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int ch2 = ch_first + c;
          const long long w1a = tick();
          uint32_t a0[2][16], x0[2][16];
          const uint32_t base = tlane + b * 128 + ch2 * 32;
          tmem_ld16(base, a0[0]);
          tmem_ld16(base + 64, x0[0]);
          tmem_ld16(base + 16, a0[1]);
          tmem_ld16(base + 80, x0[1]);
          tmem_ld_wait();
          if (c == NCH - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[b]);
          }
          const long long w1b = tick();
          ph_ld += w1b - w1a;
          uint32_t ph[2][8], pl[2][8];
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int c0 = t * kAttnKeyTile + ch2 * 32 + hf * 16;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              // E = 2^11 exp(s - max) through the SFU (see attn_kernel.cuh); E - rn16(E) IS the scaled low part
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
          ph_math += w2 - w1b;
          if (c == 0) {                            // the P V product NP tiles back has consumed this P buffer
            if constexpr (G == 4) {
              mbar_wait(&peg[grp * 2 + (myp & 1u)], (myp >> 1) & 1u);
              ++myp;
            } else {
              mbar_wait(&p_empty[pb], ((v / NP) & 1) ^ 1);
            }
          }
          const long long w2b = tick();
          sw_p += w2b - w2;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const int sc = ((ch2 * 4 + ch) ^ (row & 7)) << 4;
            const int hf = ch >> 1, o = 4 * (ch & 1);
            *reinterpret_cast<uint4*>(prow_hi + sc) = make_uint4(ph[hf][o], ph[hf][o + 1], ph[hf][o + 2], ph[hf][o + 3]);
            *reinterpret_cast<uint4*>(prow_lo + sc) = make_uint4(pl[hf][o], pl[hf][o + 1], pl[hf][o + 2], pl[hf][o + 3]);
          }
          ph_sts += tick() - w2b;
        }
        const long long w3b = tick();
        fence_proxy_async();                    // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[pb]);
        ph_fence += tick() - w3b;
      };
      const int t_full = (nk % kAttnKeyTile) ? T - 1 : T;
#pragma unroll 1
      for (int t = grp; t < t_full; t += G) tile_body(t, cuda::std::false_type{});
      if (t_full < T && ((T - 1) % G) == grp) tile_body(T - 1, cuda::std::true_type{});
      const long long st_p2 = tick();
      t_sm2 += st_p2 - st_p1;
      float l;
      {
        float l0, l1;
        upk2(lsum, l0, l1);
        l = l0 + l1;
      }
      st_buf[cq * 128 + row] = l;                // four partial sums per row (G = 2: cq = 2 * grp + ch2; G = 4: cq = grp)
      named_bar_sync(1, kSmThreads);
      l = (st_buf[row] + st_buf[128 + row]) + (st_buf[256 + row] + st_buf[384 + row]);

      // ---- epilogue: O / l -> split-fp16 [rows][256] ----
      const long long w3 = tick();
      mbar_wait(o_full, it & 1);
      sw_o += tick() - w3;
      tc_fence_after();
      {
        uint32_t a0[16], x0[16];
        const uint32_t base = tlane + 384 + cq * 16;
        tmem_ld16(base, a0);
        tmem_ld16(base + 64, x0);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);     // O may be overwritten by the next item's first P V product
        if (a.part >= 0) {                       // a key-range part: raw O, row maximum and row sum; the last part to arrive merges
          float4* po = reinterpret_cast<float4*>(p.part_o + (static_cast<size_t>(a.slot) * 128 + row) * 64 + cq * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            po[j] = make_float4(__uint_as_float(a0[4 * j]) + __uint_as_float(x0[4 * j]),
                                __uint_as_float(a0[4 * j + 1]) + __uint_as_float(x0[4 * j + 1]),
                                __uint_as_float(a0[4 * j + 2]) + __uint_as_float(x0[4 * j + 2]),
                                __uint_as_float(a0[4 * j + 3]) + __uint_as_float(x0[4 * j + 3]));
          if (cq == 0) {
            p.part_ml[(static_cast<size_t>(a.slot) * 2) * 128 + row] = mx;
            p.part_ml[(static_cast<size_t>(a.slot) * 2 + 1) * 128 + row] = l;
          }
          // the part that arrives last merges all of them (no extra launch): writes, fence, CTA barrier, one atomic ticket
          __threadfence();
          named_bar_sync(1, kSmThreads);
          if (threadIdx.x == 0) {
            unsigned* cnt = p.part_cnt + (a.slot - a.part) / p.split_s;
            const bool last = atomicAdd(cnt, 1u) == static_cast<unsigned>(p.split_s - 1);
            if (last) *cnt = 0u;                   // nobody else touches the ticket of this item before the next launch
            *last_part = last;
          }
          named_bar_sync(1, kSmThreads);
          if (*last_part) {
            __threadfence();
            AttnItem whole = a;                    // a.qrow / m0 / nq / head are those of the whole item already
            attn2_merge_parts(p, whole, row, cq);
          }
          t_epi += tick() - st_p2;
          n_base += T1 + T;
          v_base += T;
          continue;
        }
        __align__(16) __half oh[16];
        __align__(16) __half ol[16];
        const float inv_l = 1.0f / (RFE_ATTN_V_SCALE * l);   // l = sum E ; both operand scales cancel here
        const bool live = a.m0 + row < a.nq;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          // rows of the 8-row padding behind the image are written as zeros (they feed the next GEMM as finite values)
          const float o = live ? (__uint_as_float(a0[j]) + __uint_as_float(x0[j])) * inv_l : 0.0f;
          split_f32(o, oh[j], ol[j]);
        }
        if (a.m0 + row < ((a.nq + 7) & ~7)) {
          const size_t o = static_cast<size_t>(a.qrow + row) * 256 + a.head * 64 + cq * 16;
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            reinterpret_cast<uint4*>(p.out_hi + o)[ch] = reinterpret_cast<const uint4*>(oh)[ch];
            reinterpret_cast<uint4*>(p.out_lo + o)[ch] = reinterpret_cast<const uint4*>(ol)[ch];
          }
        }
      }
      t_epi += tick() - st_p2;
      n_base += T1 + T;
      v_base += T;
    }
    if (sprof) {
      p.prof[10] = t_sm2;     // softmax warp 0: pass-2 loops
      p.prof[11] = sw_s;      //   waiting for scores
      p.prof[12] = sw_p;      //   waiting for a free P buffer
      p.prof[13] = t_sm1;     // pass-1 loops (incl. the max exchange)
      p.prof[14] = sw_s1;     //   waiting for scores
      p.prof[15] = t_epi;     // l exchange + wait for O + epilogue
      p.prof[18] = sw_o;      //   of which waiting for o_full
      p.prof[19] = ph_ld;     // pass 2: tcgen05.ld + wait + score-buffer release
      p.prof[20] = ph_math;   // pass 2: exp / split arithmetic
      p.prof[21] = ph_sts;    // pass 2: shared-memory stores of P
      p.prof[22] = ph_fence;  // pass 2: fence.proxy.async + arrive
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == kAttnWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (PROF && p.prof && threadIdx.x == 0 && blockIdx.x < 4096) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.prof[32 + 3 * blockIdx.x + 1] = t;
    p.prof[32 + 3 * blockIdx.x + 2] |= static_cast<unsigned long long>(clock64() - cta_c0) << 16;   // SM cycles of this CTA
  }
}

#endif  // __CUDACC__

}  // namespace rfe
