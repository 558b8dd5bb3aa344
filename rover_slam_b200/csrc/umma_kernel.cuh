// The one tensor-core kernel of the library: a TMA-fed, tcgen05 (UMMA) split-fp16 contraction
//     D[M x N] = A[M x K] * B[N x K]^T        (fp32-equivalent, see common.cuh)
// with the A tile fetched either as a plain GEMM tile (rows x K) or as the implicit-GEMM tile of a
// 3x3 convolution over an NHWC activation tensor (9 shifted TMA boxes, zero-filled out of bounds =
// the conv's zero padding), and a fused epilogue per use (bias/ReLU/2x2 max-pool/split-fp16 store for
// the SuperPoint conv stack; softmax-65 + 8x8 depth-to-space for the detector head; channel L2-norm for
// the descriptor head; bias/scale/residual/transpose variants for LightGlue's linears and attention).
//
// CTA: warps 0..E-1 = epilogue (E = 4 or 8; warp w owns TMEM lanes 32*(w%4)..+31), warp E = TMA producer (one elected
// lane), warp E+1 = TMEM allocator + MMA issuer (one elected lane).  The single-thread roles have the highest warp
// ids because the warp scheduler arbitrates highest-id-first: the MMA issuer must never wait behind epilogue warps.
// Pipeline: NUM_STAGES smem stages {A_hi, A_lo, B_hi, B_lo}, full/empty mbarriers, one tmem_full barrier.
// Replaces (reference): every Conv/MatMul node that ONNXRuntime executes for superpoint.onnx /
// lightglue_sim.onnx (src/Extractors/superpoint_onnx.cc:133-136, src/Matchers/lightglue_onnx.cpp:210-214).
#pragma once

#include "common.cuh"

namespace rfe {

enum AMode { A_GEMM = 0, A_CONV3 = 1 };
enum Epi {
  EPI_LINEAR = 0,    // (acc + bias) * scale (+ residual) -> fp32 and/or split-fp16 (row-major, head-major or transposed)
  EPI_CONV = 1,      // bias + ReLU (+ 2x2 max-pool) -> split-fp16 NHWC
  EPI_DET = 2,       // bias + softmax over 65 + drop dustbin + 8x8 depth-to-space -> fp32 heat-map
  EPI_DESC = 3,      // bias + L2 normalise over N=256 -> fp32 NHWC
  EPI_QKV = 4,       // LightGlue Wqkv: bias + rotary(q,k) * 64^-1/4 -> head-major split q,k ; v -> transposed split
};

struct UmmaParams {
  // problem
  int num_k_steps;     // K / 64 (K tail is zero-filled by TMA)
  int cin_chunks;      // conv: Cin / 64
  int M;               // GEMM: valid rows per batch;  conv: unused
  int N;               // valid output columns
  int H, W;            // conv: spatial size of the conv output (before pooling); DET/DESC: coarse h, w
  int tiles_x, tiles_y;  // conv: 16x8 pixel tiles per image
  int tiles_m, tiles_n, num_tiles;   // persistent tile scheduler: tile = (z * tiles_n + n) * tiles_m + m
  int a_batched, b_batched;  // GEMM: use blockIdx.z as the A / B tensor-map batch coordinate
  // epilogue
  const float* bias;       // [N] or null
  const float* residual;   // fp32 [M][ld_res] or null
  float* out_f32;          // fp32 [M][ld_f32] or null
  __half* out_hi;          // split-fp16 [M][ld_h] (or [N][ld_h] when transpose_h) or null
  __half* out_lo;
  long long bstride_f32, bstride_h, bstride_res;   // per blockIdx.z element strides
  int ld_f32, ld_h, ld_res;
  int transpose_h;
  int head_major;          // split store as [n/64][M][64] (attention heads), head_stride elements apart
  long long head_stride;
  int pool;
  float scale;
  // EPI_QKV only
  const float* cs;         // [M][32] cos of the positional encoding
  const float* sn;         // [M][32] sin
  __half* k_hi;            // q goes to out_hi/out_lo, k here (both head-major [4][M][64], head_stride apart)
  __half* k_lo;
  __half* vt_hi;           // V^T [256][ldv]
  __half* vt_lo;
  int ldv;
  unsigned long long* prof;   // optional [8] cycle counters written by CTA 0 (MMA thread: 0-3, epilogue warp 2: 4-7)
  // EPI_LINEAR, optional: per-row partial sums of squares of the stored values, [M][2 * tiles_n] (slot = 2 * n-tile + the
  // warp's column phase): SuperPoint's descriptor head defers its per-pixel L2 normalisation to the sampler
  float* rowss;
};

// TMA maps used by the LINEAR / QKV epilogues (unused members are never dereferenced).  All are 3-D:
//   f32, res : fp32 (N, M, Z), box {32, 32, 1}, SWIZZLE_128B (32 rows x 128 B staged per warp)
//   h_*, k_* : fp16 planes, box {32, 32, 1}, SWIZZLE_64B; row-major (N, M, Z) or head-major (64, M, heads)
//   vt_*     : transposed fp16 planes [cols][ld] seen as (M rows contiguous, cols, 1), box {32, 32, 1}, SWIZZLE_64B
struct EpiMaps {
  CUtensorMap f32, res, h_hi, h_lo, k_hi, k_lo, vt_hi, vt_lo;
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                  // fp16 elements = one 128-byte swizzle row
constexpr int kStageABytes = kBlockM * 128;  // one of A_hi / A_lo
// epilogue warps: two per TMEM lane quarter (each takes every other 64-column group) for the wide row-wise epilogues
__host__ __device__ constexpr int umma_epi_warps(int block_n, int epi) {
  // measured: 8 warps (two per quarter) were SLOWER than 4 (ffn0 42 vs 33 us): the epilogue is bound by TMEM reads and
  // memory latency, not by issue slots, and 10 warps cap the kernel at 168 registers.  Kept parameterised.
  // LINEAR / QKV: two warps per TMEM lane quarter, each taking every other 32-column chunk (32 accumulator values per
  // thread keep the kernel far below the 204-register cap of a 320-thread CTA, so the two warps of a scheduler overlap
  // each other's TMEM / shared-memory / TMA latencies).  The row-wise CONV / DET / DESC epilogues keep one warp per quarter.
  (void)block_n;
  return (epi == 0 || epi == 4) ? 8 : 4;
}
__host__ __device__ constexpr int umma_threads(int block_n, int epi) { return 64 + 32 * umma_epi_warps(block_n, epi); }

__host__ __device__ constexpr int umma_stage_bytes(int block_n) { return 2 * kStageABytes + 2 * block_n * 128; }
__host__ __device__ constexpr int umma_num_stages(int block_n) {
  return (192 * 1024 / umma_stage_bytes(block_n)) > 6 ? 6 : (192 * 1024 / umma_stage_bytes(block_n));
}
constexpr int kEpiStageWarpBytes = 4096;     // per epilogue warp: 32 rows x 128 B (32 fp32 or 64 fp16 per row)
__host__ __device__ constexpr int umma_num_stages(int block_n);
// B-resident mode (BRES, K <= 256): the CTA keeps ONE n-tile of the weights (all k-steps, hi+lo: up to 4 x 2 x block_n x
// 128 B) in shared memory for its whole life and streams only A, plane by plane, through kBresSlots 16 KB slots.
constexpr int kBresSlots = 4;
constexpr int kBresMaxKSteps = 4;
__host__ __device__ constexpr int umma_smem_bytes(int block_n, bool bres = false) {
  return (bres ? kBresSlots * kStageABytes + kBresMaxKSteps * 2 * block_n * 128
               : umma_num_stages(block_n) * umma_stage_bytes(block_n)) +
         8 * kEpiStageWarpBytes + 1024 /*align slack*/ + 256 /*barriers*/;
}
// Number of hi*hi accumulators.  The tensor core adds each K=16 partial sum into the fp32 accumulator with
// truncation, so the error grows linearly with the number of accumulation steps (measured: 1.1e-4 abs at K=512 on
// sums of magnitude 80).  The 3x3 convs (36-72 K-steps) therefore rotate over one accumulator per kernel row and
// the epilogue adds the three in round-to-nearest fp32.
__host__ __device__ constexpr int umma_num_acc0(int amode) { return amode == 1 ? 3 : 1; }
__host__ __device__ constexpr int umma_tmem_cols_n(int cols) {
  return cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
}
__host__ __device__ constexpr int umma_tmem_cols(int block_n, int amode) {
  return umma_tmem_cols_n((umma_num_acc0(amode) + 1) * block_n);
}

#ifdef __CUDACC__

template <int BLOCK_N, int AMODE, int EPI, bool BRES = false, bool FAST = false>
__global__ void __launch_bounds__(umma_threads(BLOCK_N, EPI), 1)
umma_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
            const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
            const __grid_constant__ EpiMaps em, const UmmaParams p) {
  // BRES: the "stages" are kBresSlots single-plane A slots (A_hi and A_lo of a k-step travel separately, so the hi MMAs
  // start while the lo plane is still in flight) followed by the resident B region [k-step][B_hi | B_lo].
  constexpr int STAGES = BRES ? kBresSlots : umma_num_stages(BLOCK_N);
  constexpr int STAGE_BYTES = BRES ? kStageABytes : umma_stage_bytes(BLOCK_N);
  constexpr int STAGE_B = BLOCK_N * 128;
  constexpr int BRES_BYTES = BRES ? kBresMaxKSteps * 2 * STAGE_B : 0;
  static_assert(!BRES || (AMODE == A_GEMM && 2 * BLOCK_N <= 256), "B-resident mode: plain GEMM with the concatenated-B MMA");
  constexpr int NACC0 = umma_num_acc0(AMODE);
  constexpr int TILE_COLS = umma_tmem_cols(BLOCK_N, AMODE);  // TMEM columns of one output tile's accumulators
  constexpr int NBUF = TILE_COLS <= 256 ? 2 : 1;             // accumulator sets: the epilogue of tile i overlaps the MMAs of tile i+1
  constexpr int TMEM_COLS = NBUF * TILE_COLS;
  constexpr int ACC_STRIDE = TILE_COLS / (NACC0 + 1);        // column pitch between accumulators
  constexpr int ACC1_COL = NACC0 * ACC_STRIDE;
  static_assert(ACC_STRIDE >= BLOCK_N && TMEM_COLS <= 512, "TMEM budget");
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "invalid UMMA N");
  static_assert(STAGES >= 2, "need at least two stages");

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by offsetting the __shared__ array itself (keeps the shared address space: STS/LDS, not generic)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int EPI_WARPS = umma_epi_warps(BLOCK_N, EPI);
  constexpr int NHALF = EPI_WARPS / 4;                   // warps sharing a TMEM lane quarter
  uint8_t* bres = smem + STAGES * STAGE_BYTES;           // BRES: resident weights
  uint8_t* epi_stage = bres + BRES_BYTES;                // 8 x kEpiStageWarpBytes
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + 8 * kEpiStageWarpBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;       // [NBUF]
  uint64_t* tmem_empty_bar = tmem_full_bar + NBUF;    // [NBUF]
  uint64_t* res_bar = tmem_empty_bar + NBUF;          // [8] one per epilogue warp: residual tile landed
  uint64_t* bres_bar = res_bar + 8;                   // BRES: the resident weights have landed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bres_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  auto tick = [&]() -> long long { return p.prof ? clock64() : 0; };   // role counters only when armed
  constexpr int WARP_TMA = EPI_WARPS, WARP_MMA = EPI_WARPS + 1;

  // ---- persistent tile scheduler: static round-robin over (m, n, z) tiles; consecutive tiles share the B tile ------
  struct Tile { int m0, n0, img, x0, y0, z; };
  auto decode = [&](int tile) {
    Tile t;
    const int bx = tile % p.tiles_m;
    const int rest = tile / p.tiles_m;
    t.n0 = (rest % p.tiles_n) * BLOCK_N;
    t.z = rest / p.tiles_n;
    t.m0 = 0; t.img = 0; t.x0 = 0; t.y0 = 0;
    if constexpr (AMODE == A_CONV3) {
      const int tx = bx % p.tiles_x;
      const int r2 = bx / p.tiles_x;
      t.x0 = tx * 16;
      t.y0 = (r2 % p.tiles_y) * 8;
      t.img = r2 / p.tiles_y;
    } else {
      t.m0 = bx * kBlockM;
    }
    return t;
  };

  // BRES: a CTA is bound to one n-tile (blockIdx.x % tiles_n) and walks m-tiles with stride gridDim.x / tiles_n; the host
  // launches a multiple of tiles_n CTAs.  Otherwise: round-robin over all tiles.
  const int bres_n = BRES ? static_cast<int>(blockIdx.x) % p.tiles_n : 0;
  const int tile_first = BRES ? bres_n * p.tiles_m + static_cast<int>(blockIdx.x) / p.tiles_n : static_cast<int>(blockIdx.x);
  const int tile_step = BRES ? static_cast<int>(gridDim.x) / p.tiles_n : static_cast<int>(gridDim.x);
  const int tile_end = BRES ? (bres_n + 1) * p.tiles_m : p.num_tiles;

  // ---- one-time setup ------------------------------------------------------------------------------
  if (warp == WARP_TMA && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmA_lo);
    tma_prefetch_desc(&tmB_hi);
    tma_prefetch_desc(&tmB_lo);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], EPI_WARPS);      // one arrival per epilogue warp
    }
    for (int w = 0; w < 8; ++w) mbar_init(&res_bar[w], 1);
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == WARP_MMA) {
    tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == WARP_TMA) {
    // ===== TMA producer =====================================================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      if constexpr (BRES) {
        if (tile_first < tile_end) {
          mbar_expect_tx(bres_bar, static_cast<uint32_t>(p.num_k_steps) * 2 * STAGE_B);
          for (int ks = 0; ks < p.num_k_steps; ++ks) {
            tma_load_3d(bres + ks * 2 * STAGE_B, &tmB_hi, bres_bar, ks * 64, bres_n * BLOCK_N, 0);
            tma_load_3d(bres + ks * 2 * STAGE_B + STAGE_B, &tmB_lo, bres_bar, ks * 64, bres_n * BLOCK_N, 0);
          }
        }
      }
      for (int tile = tile_first; tile < tile_end; tile += tile_step) {
      const Tile tc = decode(tile);
      const int m0 = tc.m0, n0 = tc.n0, img = tc.img, x0 = tc.x0, y0 = tc.y0;
      (void)m0; (void)img; (void)x0; (void)y0;
      if constexpr (BRES) {
        for (int ks = 0; ks < p.num_k_steps; ++ks)
          for (int pl = 0; pl < 2; ++pl) {       // A_hi then A_lo of this k-step, one slot each
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], kStageABytes);
            tma_load_3d(smem + stage * STAGE_BYTES, pl ? &tmA_lo : &tmA_hi, &full_bar[stage], ks * 64, m0, 0);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
      } else {
      for (int ks = 0; ks < p.num_k_steps; ++ks) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + stage * STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
        if constexpr (AMODE == A_CONV3) {
          const int tap = ks / p.cin_chunks;
          const int cc = ks - tap * p.cin_chunks;
          const int dy = tap / 3, dx = tap - dy * 3;
          tma_load_4d(st, &tmA_hi, &full_bar[stage], cc * 64, x0 + dx - 1, y0 + dy - 1, img);
          tma_load_4d(st + kStageABytes, &tmA_lo, &full_bar[stage], cc * 64, x0 + dx - 1, y0 + dy - 1, img);
        } else {
          const int zb = p.a_batched ? tc.z : 0;
          tma_load_3d(st, &tmA_hi, &full_bar[stage], ks * 64, m0, zb);
          tma_load_3d(st + kStageABytes, &tmA_lo, &full_bar[stage], ks * 64, m0, zb);
        }
        const int zb = p.b_batched ? tc.z : 0;
        tma_load_3d(st + 2 * kStageABytes, &tmB_hi, &full_bar[stage], ks * 64, n0, zb);
        tma_load_3d(st + 2 * kStageABytes + STAGE_B, &tmB_lo, &full_bar[stage], ks * 64, n0, zb);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      }
      }
    }
  } else if (warp == WARP_MMA) {
    // ===== MMA issuer ===========================================================================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(kBlockM, BLOCK_N);
      constexpr bool kConcatB = (NACC0 == 1) && (ACC_STRIDE == BLOCK_N) && (2 * BLOCK_N <= 256);
      constexpr uint32_t idesc2 = make_idesc_f16(kBlockM, kConcatB ? 2 * BLOCK_N : BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      long long w_te = 0, w_full = 0;
      const long long t_begin = tick();
      for (int tile = tile_first; tile < tile_end; tile += tile_step, ++it) {
      const int buf = it % NBUF;
      const uint32_t tile_base = tmem_base + buf * TILE_COLS;
      const uint32_t acc1 = tile_base + ACC1_COL;
      long long c0 = tick();
      mbar_wait(&tmem_empty_bar[buf], (((it / NBUF) & 1) ^ 1));     // the epilogue has drained this accumulator set
      w_te += tick() - c0;
      tc_fence_after();
      if constexpr (BRES) {
        if (it == 0) {
          c0 = tick();
          mbar_wait(bres_bar, 0);
          w_full += tick() - c0;
        }
        for (int ks = 0; ks < p.num_k_steps; ++ks) {
          const uint32_t b_hi = smem_u32(bres + ks * 2 * STAGE_B);     // B_lo follows B_hi: one N = 2*BLOCK_N operand
          // slot with A_hi: [acc0 | acc1] (+)= A_hi [B_hi;B_lo]^T
          c0 = tick();
          mbar_wait(&full_bar[stage], phase);
          w_full += tick() - c0;
          tc_fence_after();
          {
            const uint32_t a = smem_u32(smem + stage * STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_f16(tile_base, make_sw128_kmajor_desc(a + k * 32), make_sw128_kmajor_desc(b_hi + k * 32), idesc2,
                       (ks > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
          // slot with A_lo: acc1 += A_lo B_hi^T
          c0 = tick();
          mbar_wait(&full_bar[stage], phase);
          w_full += tick() - c0;
          tc_fence_after();
          {
            const uint32_t a = smem_u32(smem + stage * STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_f16(acc1, make_sw128_kmajor_desc(a + k * 32), make_sw128_kmajor_desc(b_hi + k * 32), idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      } else {
      for (int ks = 0; ks < p.num_k_steps; ++ks) {
        c0 = tick();
        mbar_wait(&full_bar[stage], phase);
        w_full += tick() - c0;
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t a_lo = a_hi + kStageABytes;
        const uint32_t b_hi = a_hi + 2 * kStageABytes;
        const uint32_t b_lo = b_hi + STAGE_B;
        uint32_t acc0 = tile_base;
        bool first0 = (ks == 0);
        if constexpr (AMODE == A_CONV3) {       // one hi*hi accumulator per kernel row dy
          const int tap = ks / p.cin_chunks;
          const int dy = tap / 3;
          acc0 = tile_base + dy * ACC_STRIDE;
          first0 = (ks == dy * 3 * p.cin_chunks);
        }
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          const uint64_t da_hi = make_sw128_kmajor_desc(a_hi + k * 32);
          const uint64_t da_lo = make_sw128_kmajor_desc(a_lo + k * 32);
          const uint64_t db_hi = make_sw128_kmajor_desc(b_hi + k * 32);
          const uint64_t db_lo = make_sw128_kmajor_desc(b_lo + k * 32);
          const uint32_t acc = (ks > 0 || k > 0) ? 1u : 0u;
          if constexpr (FAST) {                    // labelled fast mode: fp16 operands, fp32 accumulate, one product of three
            umma_f16(acc0, da_hi, db_hi, idesc, (first0 && k == 0) ? 0u : 1u);
          } else if constexpr (kConcatB) {
            // [acc0 | acc1] (+)= A_hi [B_hi;B_lo]^T in ONE N = 2*BLOCK_N MMA (B_lo follows B_hi in the stage and acc1
            // follows acc0 in TMEM), then acc1 += A_lo B_hi^T: A_hi is read from shared memory once instead of twice
            umma_f16(acc0, da_hi, db_hi, idesc2, acc);
            umma_f16(acc1, da_lo, db_hi, idesc, 1u);
            (void)db_lo;
          } else {
            umma_f16(acc0, da_hi, db_hi, idesc, (first0 && k == 0) ? 0u : 1u);
            umma_f16(acc1, da_hi, db_lo, idesc, acc);
            umma_f16(acc1, da_lo, db_hi, idesc, 1u);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      }
      umma_commit(&tmem_full_bar[buf]);
      }
      if (p.prof && blockIdx.x == 0) {
        p.prof[0] = tick() - t_begin;   // MMA-thread loop
        p.prof[1] = w_te;                  // waiting for the epilogue (TMEM drain)
        p.prof[2] = w_full;                // waiting for TMA (operands)
        p.prof[3] = it;                    // tiles
      }
    }
  } else {
    // ===== epilogue warps ===========================================================================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;                // row of the 128-row tile
    int it = 0;
    uint32_t res_phase = 0;
    long long w_tf = 0;
    const long long e_begin = tick();
#ifdef RFE_EPI_PROF
    long long ep_other = 0, ep_ld = 0, ep_math = 0, ep_res = 0, ep_f32 = 0, ep_split = 0, ep_wait = 0, ep_store = 0;
#endif
    for (int tile = tile_first; tile < tile_end; tile += tile_step, ++it) {
    const Tile tc = decode(tile);
    const int m0 = tc.m0, n0 = tc.n0, img = tc.img, x0 = tc.x0, y0 = tc.y0;
    (void)m0; (void)img; (void)x0; (void)y0;
    const int buf = it % NBUF;
    const long long e0 = tick();
    mbar_wait(&tmem_full_bar[buf], (it / NBUF) & 1);
    w_tf += tick() - e0;
    tc_fence_after();
    const uint32_t t0 = tmem_base + buf * TILE_COLS + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t t1 = t0 + ACC1_COL;

    auto load16 = [&](int col, float (&v)[16]) {
      uint32_t r0[16], r1[16];
      tmem_ld16(t0 + col, r0);
      tmem_ld16(t1 + col, r1);
      if constexpr (NACC0 == 3) {
        uint32_t ra[16], rb[16];
        tmem_ld16(t0 + ACC_STRIDE + col, ra);
        tmem_ld16(t0 + 2 * ACC_STRIDE + col, rb);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          v[j] = ((__uint_as_float(r0[j]) + __uint_as_float(ra[j])) + __uint_as_float(rb[j])) +
                 (FAST ? 0.0f : __uint_as_float(r1[j]) * RFE_SPLIT_INV);        // fast mode: acc1 was never written
      } else {
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r0[j]) + (FAST ? 0.0f : __uint_as_float(r1[j]) * RFE_SPLIT_INV);
      }
    };

    // ---- coalescing helpers: a thread owns one row of the tile, but a warp-wide store of "my row" touches 32 rows.
    // Values are therefore transposed through a per-warp staging buffer of 32 rows x 128 B (XOR-swizzled 16-byte
    // chunks, conflict free) and written back row-contiguously, 4 rows x 128 B per instruction.
    const int ew = warp;                         // epilogue warp index
    const int half = ew >> 2;                    // which of the NHALF column-group phases this warp takes
    uint8_t* wst = epi_stage + ew * kEpiStageWarpBytes;
    auto load64 = [&](int col, float (&v)[64]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float t[16];
        load16(col + 16 * i, t);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[16 * i + j] = t[j];
      }
    };
    auto st_row = [&](int ch, const uint4& x) {          // chunk ch (0..7) of MY row
      *reinterpret_cast<uint4*>(wst + lane * 128 + ((ch ^ (lane & 7)) << 4)) = x;
    };
    auto ld_row = [&](int ch) {                            // chunk ch of MY row
      return *reinterpret_cast<const uint4*>(wst + lane * 128 + ((ch ^ (lane & 7)) << 4));
    };
    auto ld_staged = [&](int i, int& r, int& ch) {         // instruction i (0..7): row r = 4i + lane/8, chunk ch = lane%8
      r = 4 * i + (lane >> 3);
      ch = lane & 7;
      return *reinterpret_cast<const uint4*>(wst + r * 128 + ((ch ^ (r & 7)) << 4));
    };
    auto st_staged = [&](int r, int ch, const uint4& x) {
      *reinterpret_cast<uint4*>(wst + r * 128 + ((ch ^ (r & 7)) << 4)) = x;
    };
    // stage 32 fp32 (v[off .. off+32)) of my row
    auto stage_f32x32 = [&](const float (&v)[64], int off) {
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const float4 t = make_float4(v[off + 4 * ch], v[off + 4 * ch + 1], v[off + 4 * ch + 2], v[off + 4 * ch + 3]);
        st_row(ch, *reinterpret_cast<const uint4*>(&t));
      }
    };
    // coalesced store of a split 64-column group: off(r, ch) returns the element offset of chunk ch (8 halves) of staged
    // row r, or -1 when that row is not to be written.  One plane at a time through the 4 KB staging buffer.
    auto store_split = [&](const float (&v)[64], auto&& off, __half* dh, __half* dl) {
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          __align__(16) __half x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __half h, l;
            split_f32(v[8 * ch + j], h, l);
            x[j] = pl ? l : h;
          }
          st_row(ch, *reinterpret_cast<const uint4*>(x));
        }
        __syncwarp();
        __half* d = pl ? dl : dh;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int r, ch;
          const uint4 x = ld_staged(i, r, ch);
          const long long o = off(r, ch);
          if (o >= 0) *reinterpret_cast<uint4*>(d + o) = x;
        }
        __syncwarp();
      }
    };

    if constexpr (EPI == EPI_CONV) {
      // lane <-> pixel (ry = 2q + lane/16, rx = lane%16) of the 16x8 tile; channels in groups of 64
      const int Ho = p.pool ? p.H >> 1 : p.H, Wo = p.pool ? p.W >> 1 : p.W;
#pragma unroll 1
      for (int c = half * 64; c < BLOCK_N; c += 64 * NHALF) {
        float v[64];
        load64(c, v);
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + j));
          v[j] = fmaxf(v[j] + b4.x, 0.0f);
          v[j + 1] = fmaxf(v[j + 1] + b4.y, 0.0f);
          v[j + 2] = fmaxf(v[j + 2] + b4.z, 0.0f);
          v[j + 3] = fmaxf(v[j + 3] + b4.w, 0.0f);
        }
        if (p.pool) {
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            float t = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
            v[j] = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, 16));
          }
        }
        // staged row r is the tile pixel (ry = 2q + r/16, rx = r%16); with pooling only the 2x2 blocks' top-left
        // pixels (r = 0, 2, .., 14) are written
        store_split(v, [&](int r, int ch) -> long long {
          const int ry = 2 * q + (r >> 4), rx = r & 15;
          const int y = y0 + ry, x = x0 + rx;
          if (y >= p.H || x >= p.W) return -1;
          if (p.pool && (((r >> 4) | rx) & 1)) return -1;
          const int yo = p.pool ? y >> 1 : y, xo = p.pool ? x >> 1 : x;
          return static_cast<long long>(((static_cast<size_t>(img) * Ho + yo) * Wo + xo) * p.N + n0 + c + ch * 8);
        }, p.out_hi, p.out_lo);
      }
    } else if constexpr (EPI == EPI_DET) {
      // row = coarse pixel; 65 logits -> softmax -> first 64 -> 8x8 block of the heat-map
      static_assert(EPI != EPI_DET || BLOCK_N == 80, "detector head uses N = 80 (65 padded)");
      const int m = m0 + row;
      const bool valid = m < p.M;
      float e[80];
#pragma unroll
      for (int c = 0; c < 80; c += 16) {
        float v[16];
        load16(c, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) e[c + j] = v[j];
      }
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 65; ++j) {
        e[j] += __ldg(p.bias + j);
        mx = fmaxf(mx, e[j]);
      }
      float sum = 0.0f;
#pragma unroll
      for (int j = 0; j < 65; ++j) {
        e[j] = expf(e[j] - mx);
        sum += e[j];
      }
      if (valid) {
        const int hw = p.H * p.W;
        const int b = m / hw;
        const int r = m - b * hw;
        const int cy = r / p.W, cx = r - cy * p.W;
        float* base = p.out_f32 + (static_cast<size_t>(b) * p.H * 8 + cy * 8) * (p.W * 8) + cx * 8;
#pragma unroll
        for (int dy = 0; dy < 8; ++dy) {
          float4 a = make_float4(e[dy * 8 + 0] / sum, e[dy * 8 + 1] / sum, e[dy * 8 + 2] / sum, e[dy * 8 + 3] / sum);
          float4 c4 = make_float4(e[dy * 8 + 4] / sum, e[dy * 8 + 5] / sum, e[dy * 8 + 6] / sum, e[dy * 8 + 7] / sum);
          float4* d = reinterpret_cast<float4*>(base + static_cast<size_t>(dy) * p.W * 8);
          d[0] = a;
          d[1] = c4;
        }
      }
    } else if constexpr (EPI == EPI_DESC) {
      float ss = 0.0f;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 16) {
        float v[16];
        load16(c, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float t = v[j] + __ldg(p.bias + n0 + c + j);
          ss += t * t;
        }
      }
      const float nrm = fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 64) {
        float v[64];
        load64(c, v);
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = (v[j] + __ldg(p.bias + n0 + c + j)) / nrm;
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          stage_f32x32(v, 32 * hb);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            int r, ch;
            const uint4 x = ld_staged(i, r, ch);
            const int mr = m0 + q * 32 + r;
            if (mr < p.M)
              *reinterpret_cast<uint4*>(p.out_f32 + static_cast<size_t>(mr) * p.ld_f32 + n0 + c + 32 * hb + ch * 4) = x;
          }
          __syncwarp();
        }
      }
    } else {
      // ===== EPI_LINEAR / EPI_QKV: 32-column chunks, TMA-stored through this warp's 4 KB staging buffer ===========
      // A thread owns one row of the tile (its TMEM lane).  Per chunk: accumulators -> registers -> bias / scale /
      // rotary / residual -> staged row-wise in the swizzle of the output tensor map -> ONE bulk tensor store per
      // plane issued by lane 0 (no per-thread global addressing, rows beyond M and columns beyond N are clipped by the
      // TMA unit).  The residual tile arrives the same way (bulk tensor load into the staging buffer, prefetched
      // before the TMEM reads).  The accumulator set is handed back to the MMA warp right after the last TMEM read.
      const int row0 = m0 + q * 32;
      const int m = row0 + lane;
      const bool valid = m < p.M;
      const int z = tc.z;
      const bool has_res = (EPI == EPI_LINEAR) && p.residual != nullptr;
      uint64_t* rbar = &res_bar[ew];
      float ss_acc = 0.0f;                       // p.rowss: sum of squares of this thread's values of the tile
#ifdef RFE_EPI_PROF
#define EPI_T(var) var += clock64() - ep_t; ep_t = clock64();
      long long ep_t = clock64();
#else
#define EPI_T(var)
#endif
      int c_last = -1;
      for (int c = half * 32; c < BLOCK_N && n0 + c < p.N; c += 64) c_last = c;
      if (c_last < 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
      }
#pragma unroll 1
      for (int c = half * 32; c <= c_last; c += 64) {
        const int nb = n0 + c;
        if (has_res && lane == 0) {
          bulk_wait_read<0>();                         // earlier stores have finished reading the staging buffer
          mbar_expect_tx(rbar, 4096);
          tma_load_3d(wst, &em.res, rbar, nb, row0, z);
        }
        EPI_T(ep_other)
        float v[32];
        {
          uint32_t r0[32], r1[32];
          tmem_ld32(t0 + c, r0);
          tmem_ld32(t1 + c, r1);
          tmem_ld_wait();
          if (c == c_last) {                           // last TMEM read of this tile by this warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]) * RFE_SPLIT_INV;
        }
        EPI_T(ep_ld)
        if (p.bias) {
          if (nb + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nb + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j < p.N) v[j] += __ldg(p.bias + nb + j);
          }
        }
        if constexpr (EPI == EPI_QKV) {
          // columns: [q (256) | k (256) | v (256)], each head-major h*64+d (weights were permuted at load time);
          // a 32-column chunk is half a head of one of q / k / v: frequencies (nb & 63)/2 .. +16
          if ((n0 >> 8) < 2 && valid) {
            const float4* c4 = reinterpret_cast<const float4*>(p.cs + static_cast<size_t>(m) * 32 + ((nb & 63) >> 1));
            const float4* s4 = reinterpret_cast<const float4*>(p.sn + static_cast<size_t>(m) * 32 + ((nb & 63) >> 1));
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 cc = __ldg(c4 + g), sn = __ldg(s4 + g);
              const float cv[4] = {cc.x, cc.y, cc.z, cc.w}, sv[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int f = 4 * g + i;
                const float a = v[2 * f], b = v[2 * f + 1];
                v[2 * f] = (a * cv[i] + (-b) * sv[i]) * p.scale;
                v[2 * f + 1] = (b * cv[i] + a * sv[i]) * p.scale;
              }
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= p.scale;
          if (p.rowss) {
#pragma unroll
            for (int j = 0; j < 32; ++j) ss_acc = fmaf(v[j], v[j], ss_acc);     // columns beyond N are zero (zero-filled weights)
          }
        }
        EPI_T(ep_math)
        if (has_res) {
          mbar_wait(rbar, res_phase);
          res_phase ^= 1;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const float4 t = *reinterpret_cast<const float4*>(wst + lane * 128 + ((ch ^ (lane & 7)) << 4));
            v[4 * ch] += t.x; v[4 * ch + 1] += t.y; v[4 * ch + 2] += t.z; v[4 * ch + 3] += t.w;
          }
        }
        EPI_T(ep_res)
        if (EPI == EPI_LINEAR && p.out_f32) {
          if (!has_res && lane == 0) bulk_wait_read<0>();
          __syncwarp();                                // everyone has read the residual / the old tile has been fetched
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            *reinterpret_cast<float4*>(wst + lane * 128 + ((ch ^ (lane & 7)) << 4)) =
                make_float4(v[4 * ch], v[4 * ch + 1], v[4 * ch + 2], v[4 * ch + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&em.f32, wst, nb, row0, z);
            bulk_commit();
          }
        }
        EPI_T(ep_f32)
        const bool is_v = (EPI == EPI_QKV) && (n0 >> 8) == 2;
        if (EPI == EPI_QKV || p.out_hi) {
          uint32_t ph[16], pl[16];
          const bool vt_store = is_v || (EPI == EPI_LINEAR && p.transpose_h);
          if (vt_store) {
            // V^T feeds only the attention kernel, whose P operand carries both planes at one scale (attn_kernel.cuh):
            // V is split the same way, hi = rn16(kAttnVScale v), lo = rn16(kAttnVScale v - hi) with NO 2^11 on lo
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float s0 = v[2 * j] * RFE_ATTN_V_SCALE, s1 = v[2 * j + 1] * RFE_ATTN_V_SCALE;
              const __half2 h2 = __floats2half2_rn(s0, s1);
              const float2 hf = __half22float2(h2);
              const __half2 l2 = __floats2half2_rn(s0 - hf.x, s1 - hf.y);
              ph[j] = *reinterpret_cast<const uint32_t*>(&h2);
              pl[j] = *reinterpret_cast<const uint32_t*>(&l2);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const __half2 h2 = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
              const float2 hf = __half22float2(h2);
              const __half2 l2 = __floats2half2_rn((v[2 * j] - hf.x) * RFE_SPLIT_SCALE, (v[2 * j + 1] - hf.y) * RFE_SPLIT_SCALE);
              ph[j] = *reinterpret_cast<const uint32_t*>(&h2);
              pl[j] = *reinterpret_cast<const uint32_t*>(&l2);
            }
          }
          if (vt_store) {
            // transposed planes [col][ld]: the 32 x 32 chunk is transposed THROUGH the staging buffer -- staged row = output
            // column j, 64 bytes = my warp's 32 rows, in the SWIZZLE_64B pattern of the vt_* maps -- and leaves as two
            // bulk tensor stores (rows beyond M are clipped by the TMA unit).  One 2-byte shared store per value instead
            // of one 2-byte global store: to_v used to be 40 % slower than to_qk for the same GEMM.
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
            const int cb = (lane >> 3), ci = (lane & 7) * 2;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const __half2 h2 = *reinterpret_cast<const __half2*>(&ph[j]);
              const __half2 l2 = *reinterpret_cast<const __half2*>(&pl[j]);
              const int r0 = 2 * j, r1 = 2 * j + 1;                 // staged rows = columns 2j, 2j+1 of the chunk
              const int o0 = r0 * 64 + ((cb ^ ((r0 >> 1) & 3)) << 4) + ci;
              const int o1 = r1 * 64 + ((cb ^ ((r1 >> 1) & 3)) << 4) + ci;
              *reinterpret_cast<__half*>(wst + o0) = __low2half(h2);
              *reinterpret_cast<__half*>(wst + o1) = __high2half(h2);
              *reinterpret_cast<__half*>(wst + 2048 + o0) = __low2half(l2);
              *reinterpret_cast<__half*>(wst + 2048 + o1) = __high2half(l2);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              const int col0 = is_v ? (nb & 255) : nb;
              tma_store_3d(&em.vt_hi, wst, row0, col0, z);
              tma_store_3d(&em.vt_lo, wst + 2048, row0, col0, z);
              bulk_commit();
            }
          } else {
            EPI_T(ep_split)
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
            EPI_T(ep_wait)
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              const int so = lane * 64 + ((ch ^ ((lane >> 1) & 3)) << 4);
              *reinterpret_cast<uint4*>(wst + so) = make_uint4(ph[4 * ch], ph[4 * ch + 1], ph[4 * ch + 2], ph[4 * ch + 3]);
              *reinterpret_cast<uint4*>(wst + 2048 + so) = make_uint4(pl[4 * ch], pl[4 * ch + 1], pl[4 * ch + 2], pl[4 * ch + 3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              const bool to_k = (EPI == EPI_QKV) && (n0 >> 8) == 1;
              const bool hm = (EPI == EPI_QKV) || p.head_major;
              const int c0 = hm ? (nb & 63) : nb;
              const int c2 = hm ? ((nb & 255) >> 6) : z;
              tma_store_3d(to_k ? &em.k_hi : &em.h_hi, wst, c0, row0, c2);
              tma_store_3d(to_k ? &em.k_lo : &em.h_lo, wst + 2048, c0, row0, c2);
              bulk_commit();
            }
            EPI_T(ep_store)
          }
        }
      }
      if (EPI == EPI_LINEAR && p.rowss && valid)
        p.rowss[static_cast<size_t>(m) * (2 * p.tiles_n) + 2 * (n0 / BLOCK_N) + half] = ss_acc;
    }
    if constexpr (EPI != EPI_LINEAR && EPI != EPI_QKV) {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);     // hand the accumulator set back to the MMA warp
    }
    }
    if constexpr (EPI == EPI_LINEAR || EPI == EPI_QKV) {
      if (lane == 0) bulk_wait<0>();       // the staging buffers must outlive the last bulk stores
    }
    if (p.prof && blockIdx.x == 0 && warp == 0 && lane == 0) {
      p.prof[4] = tick() - e_begin;     // epilogue-warp loop
#ifdef RFE_EPI_PROF
      p.prof[8] = ep_other; p.prof[9] = ep_ld; p.prof[10] = ep_math; p.prof[11] = ep_res; p.prof[12] = ep_f32;
      p.prof[13] = ep_split; p.prof[14] = ep_wait; p.prof[15] = ep_store;
#endif
      p.prof[5] = w_tf;                    // waiting for accumulators
    }
  }

  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

#endif  // __CUDACC__

}  // namespace rfe
