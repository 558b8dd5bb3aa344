// 3x3 convolution 64 -> 64 channels (SuperPoint conv1b / conv2a / conv2b = 65 % of the extractor's FLOPs) as a
// "strip" implicit GEMM that reads every activation from L2 ONCE instead of once per tap.
//
// Why a second conv kernel: umma_kernel's A_CONV3 mode fetches nine shifted 16x8-pixel TMA boxes per tile, i.e. each
// input element crosses L2 -> shared memory 9 times and the kernel is bound by shared-memory bandwidth (ncu:
// tensor pipe 44 % active, l1tex 65 %).  Here an M-tile is 128 consecutive pixels of ONE image row.  The probe in
// probe_kernels.cu established that a K-major SWIZZLE_128B operand may start at any 128-byte row of a swizzled buffer
// (the swizzle is a function of the absolute shared-memory address, base_offset = 0), so all three dx taps of an
// input row are the SAME shared-memory row buffer (130 px x 128 B) read at start offsets 0 / 128 / 256 bytes, and the
// three dy taps are three different row buffers.  A CTA walks down a 128-pixel-wide column strip two output rows
// at a time with a 4-slot ring of input rows: each iteration loads only the two new rows.
//
//   per iteration (2 output rows x 128 px): A traffic 2 rows x 130 px x 256 B = 66.5 KB, weights 9 x 16 KB = 147 KB
//   (umma_kernel: 2 tiles x 9 x 48 KB = 864 KB).
//
// CTA = 384 threads: warps 0..7 = epilogue (two warps per TMEM lane quarter, 32 channels each), warp 8 = TMA producer
// for input rows, warp 9 = TMA producer for the per-tap weight tiles, warp 11 = MMA issuer (+TMEM alloc).  The
// single-thread roles have the highest warp ids: the scheduler arbitrates highest-id-first.
// Tensor-core cost model (probe 1 in probe_kernels.cu): an SS-mode MMA costs max(~60, N/2) cycles, so N = 64 wastes
// half the pipe.  Each k16 step therefore issues  A_hi x [W_hi;W_lo]^T  as ONE N=128 MMA (hi*hi and hi*lo products
// land in adjacent accumulators) plus  A_lo x W_hi^T  (N=64): 2 instructions instead of 3.
// TMEM (384 of 512 columns): output row r in {0,1} at r*192: [hh_a | hl | hh_b].  Kernel rows dy = 0,2 accumulate
// hi*hi into hh_a with the weights staged as [hi;lo] (MMA at column 0), dy = 1 into hh_b with the weights staged as
// [lo;hi] (MMA at column 64), so both share the one lo accumulator hl; two hi*hi accumulators halve the number of
// truncating accumulation steps per accumulator (see umma_kernel.cuh on accumulation truncation).
// Epilogue: bias + ReLU (+ 2x2 max-pool: vertical partner = the same thread's other row, horizontal = lane ^ 1)
// -> split-fp16 NHWC.
#pragma once

#include "common.cuh"

namespace rfe {

struct StripParams {
  int B, H, W;              // input = output spatial size (before pooling)
  int n_strips, n_segs;     // 128-px column strips per row, row segments per image
  int seg_rows;             // rows per segment (even)
  int num_items;            // B * n_strips * n_segs
  int pool;
  const float* bias;        // [64]
  __half* out_hi;           // NHWC [B][Ho][Wo][64]
  __half* out_lo;
  unsigned long long* prof;   // optional [8] cycle counters of CTA 0's MMA thread
  // FUSE1A (conv1a computed in the kernel): the u8 image and the fp32 conv1a weights replace the activation tensor maps
  const uint8_t* img;       // [B][H][img_stride]
  int img_stride;
  const float* w1a;         // [64][9]
  const float* b1a;         // [64]
};

constexpr int kStripThreads = 384;
constexpr int kStripWarpRows = 8, kStripWarpW = 9, kStripWarpMma = 11;
// FUSE1A: a fourth warpgroup (warps 12..15) computes the input rows -- conv1a + ReLU + split of the u8 image -- instead of
// the TMA row producer fetching them.  Registers are per scheduler (16 K each, four warps per scheduler at 16 warps): the
// kernel is compiled for 128 registers per thread and re-balances at run time with setmaxnreg -- the two epilogue
// warpgroups grow to 160 (their 64 accumulator values), the single-thread role warpgroup (TMA, MMA issue) shrinks to 40,
// the row producers grow to 152 (72 of them hold the conv1a weights): 160 + 160 + 40 + 152 = 512 per scheduler.
constexpr int kStripThreadsFused = 512;
constexpr int kStripRowProducerWarps = 4;
constexpr int kStripRowBytes = 130 * 128;        // one plane of one input row of the strip (with 1-px halo each side)
constexpr int kStripSlotBytes = 17 * 1024;       // slot pitch (1024-aligned for the swizzle)
constexpr int kStripRowSlots = 4;
constexpr int kStripWStages = 4;                 // per tap: [W_hi | W_lo] (dy = 0,2) or [W_lo | W_hi] (dy = 1)
constexpr int kStripWStageBytes = 2 * 8192;      // W_hi | W_lo of one tap: 64 couts x 128 B each
constexpr int kStripSmemBytes =
    2 * kStripRowSlots * kStripSlotBytes + kStripWStages * kStripWStageBytes + 1024 /*align*/ + 256 /*barriers*/;

#ifdef __CUDACC__

// FUSE1A = true is SuperPoint's conv1 group (u8 -> 1/255 -> conv1a 1->64 + ReLU -> conv1b 64->64 + ReLU ->
// 2x2 max-pool, superpoint.onnx nodes 1-5) in ONE kernel: the 480 x 640 x 64 conv1a activation (1.26 GB per 16 frames as
// split fp16) is never written to or read from HBM.  Four producer warps compute each input row of the strip on the CUDA
// cores -- K = 9 and |w| <= 197: exact fp32 FMAs in the order of conv1a_kernel, bit-identical to it -- and write it straight
// into the row slot in the 128-byte-swizzled layout the TMA box would have produced (pixel p = one 128-byte row, 16-byte
// chunk c at (c ^ (p & 7)); pixels outside the image are zeros = conv1b's padding), then fence.proxy.async + arrive.
// tmA_*: 4-D (C=64, W, H, B) box (64, 130, 1, 1);  tmW_*: 3-D (K=576, 64, 1) box (64, 64, 1)
template <bool FUSE1A, bool FAST = false>
__device__ __forceinline__ void conv64_strip_body(const CUtensorMap& tmA_hi, const CUtensorMap& tmA_lo,
                                                  const CUtensorMap& tmW_hi, const CUtensorMap& tmW_lo,
                                                  const StripParams& p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by offsetting the __shared__ array itself (keeps the shared address space: STS/LDS, not generic)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sRowHi = smem;                                              // [4] slots
  uint8_t* sRowLo = smem + kStripRowSlots * kStripSlotBytes;
  uint8_t* sW = smem + 2 * kStripRowSlots * kStripSlotBytes;           // [4] stages
  uint64_t* row_full = reinterpret_cast<uint64_t*>(sW + kStripWStages * kStripWStageBytes);
  uint64_t* row_empty = row_full + kStripRowSlots;
  uint64_t* w_full = row_empty + kStripRowSlots;
  uint64_t* w_empty = w_full + kStripWStages;
  uint64_t* acc_full = w_empty + kStripWStages;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Cycle-counter reads around the MMA issuer's waits (role counters for tools/gpu_probe.py): compiled in only with
  // -DRFE_STRIP_PACE=1 (make pace) or in the debug library.  Round 1 measured conv1b 15 % FASTER with them in
  // (profiles/r01_strip_clock_ab.txt) and shipped them; with round 2's kernel the effect is gone -- 932 us with, 910 us
  // without under ncu, tensor pipe 59.6 % vs 61.3 % (profiles/r02_strip_pace_ab.txt) -- so production no longer depends on it.
  // (A RUN-TIME switch here instead of the macro changed the kernel's register allocation, 160 -> 126, and cost 40 %.)
#if defined(RFE_STRIP_PACE) ? RFE_STRIP_PACE : defined(RFE_DEBUG_WAIT)
  auto tick = [&]() -> long long { return clock64(); };
#else
  auto tick = [&]() -> long long { return 0; };
#endif

  if (warp == kStripWarpRows && lane == 0) {
    tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmW_hi); tma_prefetch_desc(&tmW_lo);
    for (int s = 0; s < kStripRowSlots; ++s) { mbar_init(&row_full[s], FUSE1A ? kStripRowProducerWarps : 1); mbar_init(&row_empty[s], 1); }
    for (int s = 0; s < kStripWStages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 8);
    fence_barrier_init();
  }
  if (warp == kStripWarpMma) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // item -> (image, strip, segment); every role walks the same item sequence
  auto item_coords = [&](int item, int& b, int& x0, int& y_begin, int& iters) {
    const int sx = item % p.n_strips;
    const int r = item / p.n_strips;
    const int sg = r % p.n_segs;
    b = r / p.n_segs;
    x0 = sx * 128;
    y_begin = sg * p.seg_rows;
    const int rows = (p.H - y_begin) < p.seg_rows ? (p.H - y_begin) : p.seg_rows;
    iters = rows >> 1;
  };

  const int prod_rank = warp >= 12 ? warp - 12 : -1;
  if (FUSE1A && prod_rank >= 0) {
    // ===== input-row producers (FUSE1A): conv1a + ReLU + split of the u8 image, written as the TMA box would land ======
    // thread = (run of 8 pixels, group of 8 output channels): 16 runs x 8 groups = the 128 producer threads cover pixels
    // 0..127 of the 130-pixel row in one pass; pixels 128 / 129 are a one-pixel pass of 16 threads
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int tid = prod_rank * 32 + lane;
    const int cg = tid & 7;
    f32x2 wr[4][9], br[4];
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      const int c = cg * 8 + 2 * jp;
      br[jp] = pk2(__ldg(p.b1a + c), __ldg(p.b1a + c + 1));
#pragma unroll
      for (int t = 0; t < 9; ++t) wr[jp][t] = pk2(__ldg(p.w1a + c * 9 + t), __ldg(p.w1a + (c + 1) * 9 + t));
    }
    // one output pixel (8 channels) from its 3 x 3 window, as conv1a_kernel computes it (same FMA order: bit-identical)
    auto pixel = [&](const float (&w0)[3], const float (&w1)[3], const float (&w2)[3], uint4& hi4, uint4& lo4) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        f32x2 acc = pk2(0.0f, 0.0f);
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) acc = fma2(pk2(w0[dx], w0[dx]), wr[jp][dx], acc);
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) acc = fma2(pk2(w1[dx], w1[dx]), wr[jp][3 + dx], acc);
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) acc = fma2(pk2(w2[dx], w2[dx]), wr[jp][6 + dx], acc);
        acc = add2(acc, br[jp]);
        float a0, a1;
        upk2(acc, a0, a1);
        split2(pk2(fmaxf(a0, 0.0f), fmaxf(a1, 0.0f)), hi[jp], lo[jp]);
      }
      hi4 = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      lo4 = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    };
    uint32_t n = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int b, x0, y_begin, iters;
      item_coords(item, b, x0, y_begin, iters);
      const uint8_t* im = p.img + static_cast<size_t>(b) * p.H * p.img_stride;
      const int nrows = 2 * iters + 2;
      for (int k = 0; k < nrows; ++k, ++n) {
        const int slot = n & 3;
        mbar_wait(&row_empty[slot], ((n >> 2) & 1) ^ 1);
        uint8_t* dh = sRowHi + slot * kStripSlotBytes;
        uint8_t* dl = sRowLo + slot * kStripSlotBytes;
        const int y = y_begin - 1 + k;                          // row of the conv1a output map = image row
        const bool row_in = y >= 0 && y < p.H;
        // pixel pp of the slot is image column x0 - 1 + pp
        auto load_px = [&](int yy, int xx) -> float {           // u8 * (1/255) with conv1a's zero padding (transform.cpp:8)
          return (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W)
                     ? static_cast<float>(__ldg(im + static_cast<size_t>(yy) * p.img_stride + xx)) * 0.003921568859368563f
                     : 0.0f;
        };
        {
          const int run = tid >> 3;
          const int pp0 = run * 8, xs = x0 - 1 + pp0;           // pixels pp0 .. pp0 + 7
          float in[3][10];
          if (row_in) {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
              for (int i = 0; i < 10; ++i) in[dy][i] = load_px(y + dy - 1, xs - 1 + i);
          }
#pragma unroll
          for (int px = 0; px < 8; ++px) {
            const int pp = pp0 + px, x = xs + px;
            uint4 hi4 = make_uint4(0u, 0u, 0u, 0u), lo4 = hi4;
            if (row_in && x >= 0 && x < p.W) {
              const float w0[3] = {in[0][px], in[0][px + 1], in[0][px + 2]};
              const float w1[3] = {in[1][px], in[1][px + 1], in[1][px + 2]};
              const float w2[3] = {in[2][px], in[2][px + 1], in[2][px + 2]};
              pixel(w0, w1, w2, hi4, lo4);
            }
            const int o = pp * 128 + ((cg ^ (pp & 7)) << 4);
            *reinterpret_cast<uint4*>(dh + o) = hi4;
            *reinterpret_cast<uint4*>(dl + o) = lo4;
          }
        }
        if (tid < 16) {                                         // pixels 128, 129
          const int pp = 128 + (tid >> 3), x = x0 - 1 + pp;
          uint4 hi4 = make_uint4(0u, 0u, 0u, 0u), lo4 = hi4;
          if (row_in && x >= 0 && x < p.W) {
            float w[3][3];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) w[dy][dx] = load_px(y + dy - 1, x + dx - 1);
            pixel(w[0], w[1], w[2], hi4, lo4);
          }
          const int o = pp * 128 + ((cg ^ (pp & 7)) << 4);
          *reinterpret_cast<uint4*>(dh + o) = hi4;
          *reinterpret_cast<uint4*>(dl + o) = lo4;
        }
        fence_proxy_async();                                    // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(&row_full[slot]);
      }
    }
  } else if (!FUSE1A && warp == kStripWarpRows) {
    // ===== input-row producer: row sequence number n -> slot n & 3 =====================================================
    if (elect_one()) {
      uint32_t n = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        int b, x0, y_begin, iters;
        item_coords(item, b, x0, y_begin, iters);
        const int nrows = 2 * iters + 2;                       // input rows y_begin-1 .. y_begin+2*iters
        for (int k = 0; k < nrows; ++k, ++n) {
          const int slot = n & 3;
          mbar_wait(&row_empty[slot], ((n >> 2) & 1) ^ 1);
          mbar_expect_tx(&row_full[slot], 2 * kStripRowBytes);
          tma_load_4d(sRowHi + slot * kStripSlotBytes, &tmA_hi, &row_full[slot], 0, x0 - 1, y_begin - 1 + k, b);
          tma_load_4d(sRowLo + slot * kStripSlotBytes, &tmA_lo, &row_full[slot], 0, x0 - 1, y_begin - 1 + k, b);
        }
      }
    }
  } else if (warp >= 8 && warp < 12) {
    // the single-thread roles (one warpgroup: the register hand-over below is a warpgroup-wide instruction)
    if (FUSE1A) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == kStripWarpW) {
    // ===== weight producer: tap sequence number m -> stage m & 3 ======================================================
    if (elect_one()) {
      uint32_t m = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        int b, x0, y_begin, iters;
        item_coords(item, b, x0, y_begin, iters);
        for (int it = 0; it < iters; ++it)
          for (int tap = 0; tap < 9; ++tap, ++m) {
            const int st = m & 3;
            mbar_wait(&w_empty[st], ((m >> 2) & 1) ^ 1);
            mbar_expect_tx(&w_full[st], kStripWStageBytes);
            const bool swapped = (tap / 3) == 1;        // kernel row 1: [lo;hi]
            tma_load_3d(sW + st * kStripWStageBytes + (swapped ? 8192 : 0), &tmW_hi, &w_full[st], tap * 64, 0, 0);
            tma_load_3d(sW + st * kStripWStageBytes + (swapped ? 0 : 8192), &tmW_lo, &w_full[st], tap * 64, 0, 0);
          }
      }
    }
    } else if (warp == kStripWarpMma) {
    // ===== MMA issuer ==================================================================================================
    if (elect_one()) {
      constexpr uint32_t idesc64 = make_idesc_f16(128, 64);
      constexpr uint32_t idesc128 = make_idesc_f16(128, 128);
      uint32_t n_base = 0;     // row sequence number of the current item's first input row
      uint32_t m = 0;          // tap sequence number
      uint32_t gi = 0;         // iteration counter (accumulator hand-over phase)
      long long w_acc = 0, w_w = 0, w_row = 0;
      const long long t_begin = tick();
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        int b, x0, y_begin, iters;
        item_coords(item, b, x0, y_begin, iters);
        for (int it = 0; it < iters; ++it, ++gi) {
          long long c0 = tick();
          mbar_wait(acc_empty, (gi & 1) ^ 1);                 // the epilogue has drained the accumulators
          w_acc += tick() - c0;
          tc_fence_after();
          for (int dy = 0; dy < 3; ++dy) {
            for (int dx = 0; dx < 3; ++dx, ++m) {
              const int st = m & 3;
              long long c1 = tick();
              mbar_wait(&w_full[st], (m >> 2) & 1);
              w_w += tick() - c1;
              const uint32_t w_cat = smem_u32(sW + st * kStripWStageBytes);          // [hi;lo] (dy 0,2) or [lo;hi] (dy 1)
              const uint32_t w_hi = w_cat + (dy == 1 ? 8192 : 0);
#pragma unroll
              for (int r = 0; r < 2; ++r) {
                const uint32_t seq = n_base + 2 * it + r + dy;           // input row feeding output row r through kernel row dy
                const int slot = seq & 3;
                long long c2 = tick();
                mbar_wait(&row_full[slot], (seq >> 2) & 1);
                w_row += tick() - c2;
                tc_fence_after();
                const uint32_t a_hi = smem_u32(sRowHi + slot * kStripSlotBytes) + dx * 128;
                const uint32_t a_lo = smem_u32(sRowLo + slot * kStripSlotBytes) + dx * 128;
                const uint32_t row_base = tmem_base + r * 192;
                const uint32_t d_cat = row_base + (dy == 1 ? 64 : 0);    // [hh_a | hl] or [hl | hh_b]
                const uint32_t d_hl = row_base + 64;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t da_hi = make_sw128_kmajor_desc(a_hi + k * 32);
                  const uint64_t da_lo = make_sw128_kmajor_desc(a_lo + k * 32);
                  const uint64_t db_cat = make_sw128_kmajor_desc(w_cat + k * 32);
                  const uint64_t db_hi = make_sw128_kmajor_desc(w_hi + k * 32);
                  if constexpr (FAST) {          // hh_a (dy 0, 2) / hh_b (dy 1) += A_hi W_hi, nothing else
                    umma_f16(row_base + (dy == 1 ? 128 : 0), da_hi, db_hi, idesc64, (dy < 2 && dx == 0 && k == 0) ? 0u : 1u);
                    continue;
                  }
                  if (dy == 1 && dx == 0 && k == 0) {
                    // first touch of hh_b while hl already holds dy = 0: two N=64 MMAs with different accumulate flags
                    umma_f16(row_base + 128, da_hi, db_hi, idesc64, 0u);                                  // hh_b  = A_hi W_hi
                    umma_f16(d_hl, da_hi, make_sw128_kmajor_desc(w_cat + k * 32), idesc64, 1u);          // hl   += A_hi W_lo
                  } else {
                    umma_f16(d_cat, da_hi, db_cat, idesc128, (dy == 0 && dx == 0 && k == 0) ? 0u : 1u);
                  }
                  umma_f16(d_hl, da_lo, db_hi, idesc64, 1u);                                               // hl   += A_lo W_hi
                }
              }
              umma_commit(&w_empty[st]);
            }
            // input row (2*it + dy) is not read again by this item: rows 2it, 2it+1 are dead after dy = 0, 1;
            // on the item's last iteration the two look-ahead rows die after dy = 2.
            if (dy < 2) {
              umma_commit(&row_empty[(n_base + 2 * it + dy) & 3]);
            } else if (it == iters - 1) {
              umma_commit(&row_empty[(n_base + 2 * it + 2) & 3]);
              umma_commit(&row_empty[(n_base + 2 * it + 3) & 3]);
            }
          }
          umma_commit(acc_full);
        }
        n_base += 2 * iters + 2;
      }
      if (p.prof && blockIdx.x == 0) {
        p.prof[0] = tick() - t_begin;   // whole MMA-thread loop
        p.prof[1] = w_acc;                 // waiting for the epilogue to drain TMEM
        p.prof[2] = w_w;                   // waiting for weight tiles
        p.prof[3] = w_row;                 // waiting for input rows
        p.prof[4] = gi;                    // iterations
      }
    }
    }
  } else if (warp < 8) {
    // ===== epilogue warps ==============================================================================================
    if (FUSE1A) asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
    const int q = warp & 3;
    const int half = warp >> 2;                     // channels [32*half, 32*half + 32)
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t gi = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int b, x0, y_begin, iters;
      item_coords(item, b, x0, y_begin, iters);
      const int x = x0 + q * 32 + lane;
      const int Ho = p.pool ? p.H >> 1 : p.H, Wo = p.pool ? p.W >> 1 : p.W;
      for (int it = 0; it < iters; ++it, ++gi) {
        const int y = y_begin + 2 * it;
        mbar_wait(acc_full, gi & 1);
        tc_fence_after();
        // all of this warp's accumulator values first (2 rows x 32 channels), so the accumulators go back to the MMA
        // warp before the bias / pool / split / store work starts (the single accumulator set is not double-buffered)
        float v[2][32];
#pragma unroll
        for (int c = 0; c < 32; c += 16) {
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            uint32_t a0[16], a1[16], xl[16];
            const uint32_t base = tlane + r * 192 + half * 32 + c;
            tmem_ld16(base, a0);            // hh_a
            tmem_ld16(base + 128, a1);      // hh_b
            tmem_ld16(base + 64, xl);       // hl
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
              v[r][c + j] = (__uint_as_float(a0[j]) + __uint_as_float(a1[j])) + (FAST ? 0.0f : __uint_as_float(xl[j]) * RFE_SPLIT_INV);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
#pragma unroll
        for (int c = 0; c < 32; c += 16) {
          const int ch = half * 32 + c;
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int j = 0; j < 16; ++j) v[r][c + j] = fmaxf(v[r][c + j] + __ldg(p.bias + ch + j), 0.0f);
          if (p.pool) {
            __align__(16) __half hi[16];
            __align__(16) __half lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float t = fmaxf(v[0][c + j], v[1][c + j]);
              t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, 1));
              split_f32(t, hi[j], lo[j]);
            }
            if ((lane & 1) == 0 && x < p.W) {
              const size_t o = ((static_cast<size_t>(b) * Ho + (y >> 1)) * Wo + (x >> 1)) * 64 + ch;
              reinterpret_cast<uint4*>(p.out_hi + o)[0] = reinterpret_cast<const uint4*>(hi)[0];
              reinterpret_cast<uint4*>(p.out_hi + o)[1] = reinterpret_cast<const uint4*>(hi)[1];
              reinterpret_cast<uint4*>(p.out_lo + o)[0] = reinterpret_cast<const uint4*>(lo)[0];
              reinterpret_cast<uint4*>(p.out_lo + o)[1] = reinterpret_cast<const uint4*>(lo)[1];
            }
          } else {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              __align__(16) __half hi[16];
              __align__(16) __half lo[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) split_f32(v[r][c + j], hi[j], lo[j]);
              if (x < p.W) {
                const size_t o = ((static_cast<size_t>(b) * Ho + (y + r)) * Wo + x) * 64 + ch;
                reinterpret_cast<uint4*>(p.out_hi + o)[0] = reinterpret_cast<const uint4*>(hi)[0];
                reinterpret_cast<uint4*>(p.out_hi + o)[1] = reinterpret_cast<const uint4*>(hi)[1];
                reinterpret_cast<uint4*>(p.out_lo + o)[0] = reinterpret_cast<const uint4*>(lo)[0];
                reinterpret_cast<uint4*>(p.out_lo + o)[1] = reinterpret_cast<const uint4*>(lo)[1];
              }
            }
          }
        }
      }
    }
  }

  __syncthreads();
  if (warp == kStripWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

__global__ void __launch_bounds__(kStripThreads, 1)
conv64_strip_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                    const __grid_constant__ StripParams p) {
  conv64_strip_body<false>(tmA_hi, tmA_lo, tmW_hi, tmW_lo, p);
}
// labelled fast mode (rfe_set_fast_mode): hi*hi products only -- a separate instantiation, the exact kernel's code is untouched
__global__ void __launch_bounds__(kStripThreads, 1)
conv64_strip_fast_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                         const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                         const __grid_constant__ StripParams p) {
  conv64_strip_body<false, true>(tmA_hi, tmA_lo, tmW_hi, tmW_lo, p);
}
__global__ void __launch_bounds__(kStripThreadsFused, 1)
conv64_strip_fused_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                          const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                          const __grid_constant__ StripParams p) {
  conv64_strip_body<true>(tmA_hi, tmA_lo, tmW_hi, tmW_lo, p);
}

#endif  // __CUDACC__

}  // namespace rfe
