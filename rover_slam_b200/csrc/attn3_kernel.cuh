// Single-pass (online-softmax) persistent attention for LightGlue's self and cross blocks, split-fp16 on tcgen05.
//     O[q] = softmax_k( Q[q] . K[k] ) V[k]          4 heads x 64 dims, Q and K pre-scaled by 64^-1/4
// Same arithmetic as attn_kernel.cuh (three-product split-fp16 scores, same-scale P and V^T planes, E = 2^c exp(s - m)
// straight from the SFU), but the separate row-maximum pass is gone: the reference maximum m is tracked ON THE FLY and only
// raised when a key tile exceeds it by more than 2^7 ("lazy rescaling"), in which case the partial output is corrected in
// tensor memory.  Softmax is invariant to m, so the result is the exact softmax whatever sequence of m the tiles see; m only
// has to keep E = 2^8 2^(s log2e - m) inside fp16: E < 2^(8 + 7).
//
// Why it is cheap here: the sixteen softmax warps already work as two groups on alternate 64-key tiles.  Each GROUP gets
// its own output accumulator O_g in tensor memory (the score ring shrinks from three buffers to one per group: 2 x 128 +
// 2 x 128 = 512 columns) and its own running maximum, so a group is an independent online-softmax stream over its tiles:
//   * P V for tile t (group g = t & 1) accumulates into O_g; P buffer g and score buffer g belong to group g;
//   * when group g must raise m it multiplies O_g by 2^(m_old - m_new) with tcgen05.ld / tcgen05.st between "the P V
//     product of my previous tile has retired" (p_empty[g], which it waits for anyway before rewriting its P buffer) and
//     "my next P tile is ready" (p_full[g]): in that window no MMA touches O_g, so no extra protocol is needed;
//   * the two warps that share a row of a tile (32 columns each) agree on the tile maximum through a 64-thread named
//     barrier; the two groups never synchronise until the epilogue, which merges (m_0, l_0, O_0) and (m_1, l_1, O_1).
// Against the two-pass kernel this removes 1/7 of the MMAs, the second read of K_hi, the pass-1 drain of the score buffers
// and the pass-1 / pass-2 hand-over; the kernel is persistent (one CTA per SM walks (problem, head, 128-query tile) items,
// every ring runs across items with counted phases) and Q is double buffered, so the next item's scores start while the
// softmax warps are still in the epilogue of the current one.
// Replaces (reference): the attention MatMul / Softmax / MatMul nodes of lightglue_sim.onnx (layer 0 self: nodes 50-54,
// cross: 141-151) executed by ONNXRuntime at src/Matchers/lightglue_onnx.cpp:210-214.
#pragma once

#include "attn2_kernel.cuh"

namespace rfe {

constexpr int kA3KStages = 3;
constexpr int kA3VStages = 2;
constexpr int kA3ExpShift = 8;                 // E = 2^8 exp(s - m)
constexpr float kA3Tau = 7.0f;                 // raise m when a tile exceeds it by more than 2^7: E < 2^15 always
// tail: barriers (512 B) | pair maxima [2 parity][2 groups][2 halves][128] | epilogue statistics [2 item parity][6][128]
constexpr int kA3TailBytes = 512 + 2 * 2 * 2 * 128 * 4 + 2 * 6 * 128 * 4;
constexpr int kA3SmemBytes = 2 * kAttnQBytes + 2 * kAttnPBytes + (kA3KStages + kA3VStages) * kAttnKVBytes + 1024 + kA3TailBytes;

#ifdef __CUDACC__

// O_g[my lanes, my 32 columns of hh and of hl] *= fix  (whole warp; kept out of line: it runs once or twice per item and
// its 16 scratch registers would otherwise count against the 96-register budget of the hot loop)
__device__ __noinline__ void attn3_rescale_o(uint32_t t_o, int ch2, float fix) {
  tc_fence_after();
  const f32x2 f2 = pk2(fix, fix);
#pragma unroll 1
  for (int part = 0; part < 4; ++part) {       // 16 columns at a time
    const uint32_t ta = t_o + (part >> 1) * 64 + ch2 * 32 + (part & 1) * 16;
    uint32_t r[16];
    tmem_ld16(ta, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const f32x2 v = mul2(pk2u(r[j], r[j + 1]), f2);
      float v0, v1;
      upk2(v, v0, v1);
      r[j] = __float_as_uint(v0);
      r[j + 1] = __float_as_uint(v1);
    }
    tmem_st16(ta, r);
  }
  tmem_st_wait();
  tc_fence_before();
}

template <bool PROF>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn3_kernel(const __grid_constant__ CUtensorMap tmQ_hi, const __grid_constant__ CUtensorMap tmQ_lo,
             const __grid_constant__ CUtensorMap tmK_hi, const __grid_constant__ CUtensorMap tmK_lo,
             const __grid_constant__ CUtensorMap tmV_hi, const __grid_constant__ CUtensorMap tmV_lo, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                   // 2 x (Q_hi | Q_lo)
  uint8_t* sK = sQ + 2 * kAttnQBytes;                   // K stages: K_hi | K_lo (64 keys)
  uint8_t* sP = sK + kA3KStages * kAttnKVBytes;         // P buffer of group 0, of group 1: P_hi | P_lo
  uint8_t* sV = sP + 2 * kAttnPBytes;                   // V stages: Vt_hi | Vt_lo
  uint8_t* tail = sV + kA3VStages * kAttnKVBytes;
  uint64_t* q_full = reinterpret_cast<uint64_t*>(tail);   // [2]
  uint64_t* q_empty = q_full + 2;                         // [2]
  uint64_t* k_full = q_empty + 2;                         // [3]
  uint64_t* k_empty = k_full + kA3KStages;
  uint64_t* v_full = k_empty + kA3KStages;                // [2]
  uint64_t* v_empty = v_full + kA3VStages;
  uint64_t* s_full = v_empty + kA3VStages;                // [2] one score buffer per group
  uint64_t* s_empty = s_full + 2;
  uint64_t* p_full = s_empty + 2;                         // [2] one P buffer per group
  uint64_t* p_empty = p_full + 2;
  uint64_t* o_full = p_empty + 2;
  uint64_t* o_empty = o_full + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_empty + 1);
  float* pair_max = reinterpret_cast<float*>(tail + 512);            // [2][2][2][128]
  float* stat = pair_max + 2 * 2 * 2 * 128;                          // [2][6][128]: m of group 0 / 1, four partial row sums
  static_assert(24 * 8 + 4 <= 512 && kA3SmemBytes <= 227 * 1024, "tail region / shared-memory budget");

  auto tick = [&]() -> long long { return PROF ? clock64() : 0; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool prof_cta = PROF && p.prof && static_cast<int>(blockIdx.x) == p.prof_cta;
  const int n_items = p.item_prefix[p.nprob];
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
    for (int s = 0; s < 2; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
    for (int s = 0; s < kA3KStages; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < kA3VStages; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], kAttnSoftmaxWarps / 2); }
    for (int s = 0; s < 2; ++s) { mbar_init(&p_full[s], kAttnSoftmaxWarps / 2); mbar_init(&p_empty[s], 1); }
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
  // TMEM columns: scores of group g: [hh | hl] at g*128 ; O_g: [hh | hl] at 256 + g*128

  if (warp == kAttnWarpK) {
    // ===== K producer ====================================================================================================
    if (elect_one()) {
      uint32_t n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const AttnItem a = attn_decode(p, item);
        for (int t = 0; t < a.T; ++t, ++n) {
          const int st = n % kA3KStages;
          mbar_wait(&k_empty[st], ((n / kA3KStages) & 1) ^ 1);
          uint8_t* sb = sK + st * kAttnKVBytes;
          mbar_expect_tx(&k_full[st], kAttnKVBytes);
          tma_load_3d(sb, &tmK_hi, &k_full[st], 0, a.krow + t * kAttnKeyTile, a.head);
          tma_load_3d(sb + 8192, &tmK_lo, &k_full[st], 0, a.krow + t * kAttnKeyTile, a.head);
        }
      }
    }
  } else if (warp == kAttnWarpV) {
    // ===== Q + V producer: Q of item i+1 goes into the other Q buffer two key tiles into item i ======================
    if (elect_one()) {
      uint32_t n = 0, it = 0;
      auto load_q = [&](int item, uint32_t iq) {
        const AttnItem a = attn_decode(p, item);
        const int b = iq & 1;
        mbar_wait(&q_empty[b], ((iq >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[b], kAttnQBytes);
        tma_load_3d(sQ + b * kAttnQBytes, &tmQ_hi, &q_full[b], 0, a.qrow, a.head);
        tma_load_3d(sQ + b * kAttnQBytes + 16384, &tmQ_lo, &q_full[b], 0, a.qrow, a.head);
      };
      if (static_cast<int>(blockIdx.x) < n_items) load_q(blockIdx.x, 0);
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const AttnItem a = attn_decode(p, item);
        const int t_pref = a.T > 2 ? 2 : a.T - 1;
        for (int t = 0; t < a.T; ++t, ++n) {
          if (t == t_pref && item + static_cast<int>(gridDim.x) < n_items) load_q(item + gridDim.x, it + 1);
          const int st = n % kA3VStages;
          mbar_wait(&v_empty[st], ((n / kA3VStages) & 1) ^ 1);
          uint8_t* sb = sV + st * kAttnKVBytes;
          mbar_expect_tx(&v_full[st], kAttnKVBytes);
          tma_load_3d(sb, &tmV_hi, &v_full[st], a.krow + t * kAttnKeyTile, 0, a.head);
          tma_load_3d(sb + 8192, &tmV_lo, &v_full[st], a.krow + t * kAttnKeyTile, 0, a.head);
        }
      }
    }
  } else if (warp == kAttnWarpMmaS) {
    // ===== MMA issuer 1: scores ==========================================================================================
    if (elect_one()) {
      constexpr uint32_t idesc64 = make_idesc_f16(128, 64);
      constexpr uint32_t idesc128 = make_idesc_f16(128, 128);
      long long w_q = 0, w_k = 0, w_se = 0, t_loop = 0;
      uint32_t n = 0, it = 0, cg[2] = {0, 0};
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const AttnItem a = attn_decode(p, item);
        const int qb = it & 1;
        const uint32_t q_hi = smem_u32(sQ + qb * kAttnQBytes), q_lo = q_hi + 16384;
        const long long t0 = tick();
        mbar_wait(&q_full[qb], (it >> 1) & 1);
        const long long t1 = tick();
        w_q += t1 - t0;
        for (int t = 0; t < a.T; ++t, ++n) {
          const int st = n % kA3KStages, g = t & 1;
          const uint32_t c = cg[g]++;
          const long long c0 = tick();
          mbar_wait(&k_full[st], (n / kA3KStages) & 1);
          const long long c1 = tick();
          mbar_wait(&s_empty[g], (c & 1) ^ 1);
          w_k += c1 - c0;
          w_se += tick() - c1;
          tc_fence_after();
          const uint32_t k_hi = smem_u32(sK + st * kAttnKVBytes);   // K_lo follows at +8192: [K_hi;K_lo] is one N=128 operand
          const uint32_t s_base = tmem_base + g * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t dk = make_sw128_kmajor_desc(k_hi + k * 32);
            umma_f16(s_base, make_sw128_kmajor_desc(q_hi + k * 32), dk, idesc128, k > 0);        // [S_hh | S_hl]
            umma_f16(s_base + 64, make_sw128_kmajor_desc(q_lo + k * 32), dk, idesc64, 1u);       // S_hl += Q_lo K_hi^T
          }
          umma_commit(&s_full[g]);
          umma_commit(&k_empty[st]);
        }
        umma_commit(&q_empty[qb]);
        t_loop += tick() - t1;
      }
      if (prof_cta) {
        p.prof[0] = w_q;        // waiting for Q
        p.prof[2] = t_loop;     // score issue loops
        p.prof[3] = w_k;        // waiting for K tiles
        p.prof[4] = w_se;       // waiting for the group's score buffer
      }
    }
  } else if (warp == kAttnWarpMma) {
    // ===== MMA issuer 2: O_g += P V ======================================================================================
    if (elect_one()) {
      constexpr uint32_t idesc64 = make_idesc_f16(128, 64);
      constexpr uint32_t idesc128 = make_idesc_f16(128, 128);
      long long w_v = 0, w_p = 0, w_o = 0;
      uint32_t n = 0, it = 0, tiles = 0, cg[2] = {0, 0};
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const AttnItem a = attn_decode(p, item);
        for (int t = 0; t < a.T; ++t, ++n) {
          const int st = n % kA3VStages, g = t & 1;
          const uint32_t c = cg[g]++;
          const long long c0 = tick();
          mbar_wait(&v_full[st], (n / kA3VStages) & 1);
          const long long c1 = tick();
          mbar_wait(&p_full[g], c & 1);
          const long long c2 = tick();
          if (t == 0) mbar_wait(o_empty, (it & 1) ^ 1);     // the previous item's epilogue has read both accumulators
          w_v += c1 - c0;
          w_p += c2 - c1;
          w_o += tick() - c2;
          tc_fence_after();
          const uint32_t p_hi = smem_u32(sP + g * kAttnPBytes), p_lo = p_hi + 16384;
          const uint32_t v_hi = smem_u32(sV + st * kAttnKVBytes);       // V_lo follows at +8192
          const uint32_t o_base = tmem_base + 256 + g * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t dv = make_sw128_kmajor_desc(v_hi + k * 32);
            umma_f16(o_base, make_sw128_kmajor_desc(p_hi + k * 32), dv, idesc128, (t > 1 || k > 0) ? 1u : 0u);
            umma_f16(o_base + 64, make_sw128_kmajor_desc(p_lo + k * 32), dv, idesc64, 1u);
          }
          umma_commit(&v_empty[st]);
          umma_commit(&p_empty[g]);
        }
        umma_commit(o_full);
        tiles += a.T;
      }
      if (prof_cta) {
        p.prof[5] = w_v;        // waiting for V tiles
        p.prof[6] = w_p;        // waiting for P (softmax)
        p.prof[7] = tiles;      // key tiles of this CTA
        p.prof[16] = w_o;       // waiting for the O hand-back
        p.prof[17] = it;        // items of this CTA
      }
    }
  } else if (warp < kAttnSoftmaxWarps) {
    // ===== softmax / epilogue warps =======================================================================================
    const int sw = warp;                     // 0..15
    const int q = warp & 3;                  // TMEM lane quarter
    const int cq = sw >> 2;                  // epilogue: which quarter of the output columns; = 2 * grp + ch2
    const int grp = sw >> 3;                 // the group that owns key tiles t with (t & 1) == grp
    const int ch2 = (sw >> 2) & 1;           // which 32-column half of the group's tile
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t t_s = tlane + grp * 128, t_o = tlane + 256 + grp * 128;
    const int pair_bar = 2 + grp * 4 + q;    // named barrier of the two warps that share my rows of a tile
    const float NEG = -INFINITY;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr int kSmThreads = 32 * kAttnSoftmaxWarps;
    const bool sprof = prof_cta && warp == 0 && lane == 0;
    long long sw_s = 0, sw_p = 0, sw_o = 0, sw_x = 0, t_sm = 0, t_epi = 0;
    uint32_t it = 0, c = 0, n_fix = 0;       // c: tiles this group has processed so far (all items)
    const f32x2 kL2 = pk2(kLog2e, kLog2e), kL2s = pk2(kLog2e * RFE_SPLIT_INV, kLog2e * RFE_SPLIT_INV);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const AttnItem a = attn_decode(p, item);
      const int nk = a.nk, T = a.T;
      const long long st_begin = tick();
      float m_cur = NEG;                      // reference maximum of MY group's stream, in log2 units (s * log2 e)
      f32x2 lsum = pk2(0.0f, 0.0f);           // sum of E over my columns, relative to m_cur
      auto tile_body = [&](int t, auto masked_tag) {
        constexpr bool kMasked = decltype(masked_tag)::value;
        const long long w0 = tick();
        mbar_wait(&s_full[grp], c & 1);
        sw_s += tick() - w0;
        tc_fence_after();
        uint32_t a0[2][16], x0[2][16];
        const uint32_t base = t_s + ch2 * 32;
        tmem_ld16(base, a0[0]);
        tmem_ld16(base + 64, x0[0]);
        tmem_ld16(base + 16, a0[1]);
        tmem_ld16(base + 80, x0[1]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[grp]);
        // tile maximum of my row from the hi*hi product (within 2^-11 |s| of the true one: m only has to be approximate)
        float lm;
        {
          float m0a = NEG, m1a = NEG, m2a = NEG, m3a = NEG;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int c0 = t * kAttnKeyTile + ch2 * 32 + hf * 16;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float v0 = __uint_as_float(a0[hf][j]), v1 = __uint_as_float(a0[hf][j + 1]);
              float v2 = __uint_as_float(a0[hf][j + 2]), v3 = __uint_as_float(a0[hf][j + 3]);
              if (kMasked) {
                if (c0 + j >= nk) v0 = NEG;
                if (c0 + j + 1 >= nk) v1 = NEG;
                if (c0 + j + 2 >= nk) v2 = NEG;
                if (c0 + j + 3 >= nk) v3 = NEG;
              }
              m0a = fmaxf(m0a, v0); m1a = fmaxf(m1a, v1); m2a = fmaxf(m2a, v2); m3a = fmaxf(m3a, v3);
            }
          }
          lm = fmaxf(fmaxf(m0a, m1a), fmaxf(m2a, m3a)) * kLog2e;
        }
        const long long w1 = tick();
        float* pm = pair_max + (((c & 1) * 2 + grp) * 2) * 128;       // [parity][grp][ch2][128]
        pm[ch2 * 128 + row] = lm;
        named_bar_sync(pair_bar, 64);
        lm = fmaxf(lm, pm[(ch2 ^ 1) * 128 + row]);
        sw_x += tick() - w1;
        // lazy rescaling: raise the reference only when this tile exceeds it by more than 2^tau
        float fix = 1.0f;
        if (lm > m_cur + kA3Tau) {
          fix = fast_exp2(m_cur - lm);         // 0 when m_cur = -inf (first tile of the stream)
          m_cur = lm;
          lsum = mul2(lsum, pk2(fix, fix));
        }
        const f32x2 nmx = pk2(static_cast<float>(kA3ExpShift) - m_cur, static_cast<float>(kA3ExpShift) - m_cur);
        uint32_t ph[2][8], pl[2][8];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int c0 = t * kAttnKeyTile + ch2 * 32 + hf * 16;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            // E = 2^8 exp(s - m) through the SFU; E - rn16(E) IS the low part at the same scale (attn_kernel.cuh)
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
            ph[hf][j] = hE;
            float d0, d1;
            asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\t"
                "fma.rn.f32.f16 %0, l, %5, %3;\n\tfma.rn.f32.f16 %1, h, %5, %4;\n\t}"
                : "=f"(d0), "=f"(d1)
                : "r"(hE), "f"(e0), "f"(e1), "h"(static_cast<unsigned short>(0xBC00)));   // E - rn16(E), exact
            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pl[hf][j]) : "f"(d1), "f"(d0));
          }
        }
        const long long w2 = tick();
        mbar_wait(&p_empty[grp], (c & 1) ^ 1);       // the P V product of my group's previous tile has retired
        sw_p += tick() - w2;
        // O_g is quiescent from here until I publish P: correct it if any row of this warp raised its reference
        // (t >= 2: O_g holds earlier tiles of this item; the first tile of the stream starts the accumulator afresh)
        if (t >= 2 && __any_sync(0xffffffffu, fix != 1.0f)) {
          attn3_rescale_o(t_o, ch2, fix);
          if (PROF) ++n_fix;
        }
        uint8_t* prow_hi = sP + grp * kAttnPBytes + row * 128;
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
        if (lane == 0) mbar_arrive(&p_full[grp]);
        ++c;
      };
      const int t_full = (nk % kAttnKeyTile) ? T - 1 : T;
#pragma unroll 1
      for (int t = grp; t < t_full; t += 2) tile_body(t, cuda::std::false_type{});
      if (t_full < T && ((T - 1) & 1) == grp) tile_body(T - 1, cuda::std::true_type{});
      const long long st_sm = tick();
      t_sm += st_sm - st_begin;

      // ---- epilogue: merge the two groups' streams, O / l -> split-fp16 [rows][256] ----
      float* st_buf = stat + (it & 1) * 6 * 128;
      {
        float l0, l1;
        upk2(lsum, l0, l1);
        st_buf[(2 + cq) * 128 + row] = l0 + l1;
        if (ch2 == 0) st_buf[grp * 128 + row] = m_cur;
      }
      named_bar_sync(1, kSmThreads);
      const float m0 = st_buf[row], m1 = st_buf[128 + row];         // m1 = -inf when the item has a single key tile
      const float M = fmaxf(m0, m1);
      const float f0 = fast_exp2(m0 - M), f1 = (m1 == NEG) ? 0.0f : fast_exp2(m1 - M);
      const float l = f0 * (st_buf[2 * 128 + row] + st_buf[3 * 128 + row]) + f1 * (st_buf[4 * 128 + row] + st_buf[5 * 128 + row]);
      const long long w3 = tick();
      mbar_wait(o_full, it & 1);
      sw_o += tick() - w3;
      tc_fence_after();
      {
        uint32_t a0[16], x0[16], a1[16], x1[16];
        const uint32_t base = tlane + 256 + cq * 16;
        tmem_ld16(base, a0);
        tmem_ld16(base + 64, x0);
        tmem_ld16(base + 128, a1);
        tmem_ld16(base + 192, x1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);     // both accumulators may be overwritten by the next item's P V products
        __align__(16) __half oh[16];
        __align__(16) __half ol[16];
        const float inv_l = 1.0f / (RFE_ATTN_V_SCALE * l);   // the 2^8 of E and the 256 of V^T cancel here
        const bool live = a.m0 + row < a.nq;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float o = f0 * (__uint_as_float(a0[j]) + __uint_as_float(x0[j]));
          if (f1 != 0.0f) o = fmaf(f1, __uint_as_float(a1[j]) + __uint_as_float(x1[j]), o);     // O_1 is undefined when T == 1
          // rows of the 8-row padding behind the image are written as zeros (they feed the next GEMM as finite values)
          split_f32(live ? o * inv_l : 0.0f, oh[j], ol[j]);
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
      t_epi += tick() - st_sm;
    }
    if (sprof) {
      p.prof[10] = t_sm;      // softmax warp 0: key-tile loops
      p.prof[11] = sw_s;      //   waiting for scores
      p.prof[12] = sw_p;      //   waiting for a free P buffer
      p.prof[13] = sw_x;      //   pair-maximum exchange
      p.prof[14] = n_fix;     //   O corrections executed by this warp
      p.prof[15] = t_epi;     // merge + wait for O + epilogue
      p.prof[18] = sw_o;      //   of which waiting for o_full
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
