// Hardware probes (debug only, reached through rfe_debug_probe): measure tcgen05 behaviours the design depends on.
//  probe 0: "row-shifted" K-major SWIZZLE_128B operand.  An A buffer of 136 rows x 128 B is written with the absolute-
//           address swizzle (16-byte chunk index XOR address bits [7,10)); the MMA descriptor then starts at row s
//           (base + s*128 B, NOT 1024-aligned) with base_offset = 0 / (s & 7) / ((8 - s) & 7).  Reports, per (s, mode),
//           the max-abs error against the exact product A[s : s+128] B^T.  This decides whether a 3x3 conv can read
//           its 9 taps from ONE halo tile in shared memory.
#include "common.cuh"

namespace rfe {

__device__ __forceinline__ uint64_t make_desc_sw128_bo(uint32_t smem_addr, uint32_t base_offset) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_offset & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// a: [136][64] fp16, b: [64][64] fp16, out: [9 shifts][3 modes][128][64] fp32
__global__ void __launch_bounds__(128, 1) probe_shift_kernel(const __half* __restrict__ a, const __half* __restrict__ b,
                                                             float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by offsetting the __shared__ array itself (keeps the shared address space: STS/LDS, not generic)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                 // 136 rows * 128 B = 17408 -> pad to 18432
  uint8_t* sB = smem + 18432;         // 64 rows * 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 8192);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 136 * 8; i += 128) {     // 16-byte chunks
    const int r = i >> 3, ch = i & 7;
    const uint32_t addr = smem_u32(sA) + r * 128;
    const int sw = ch ^ ((addr >> 7) & 7);
    *reinterpret_cast<uint4*>(sA + r * 128 + sw * 16) = *reinterpret_cast<const uint4*>(a + r * 64 + ch * 8);
  }
  for (int i = threadIdx.x; i < 64 * 8; i += 128) {
    const int r = i >> 3, ch = i & 7;
    const uint32_t addr = smem_u32(sB) + r * 128;
    const int sw = ch ^ ((addr >> 7) & 7);
    *reinterpret_cast<uint4*>(sB + r * 128 + sw * 16) = *reinterpret_cast<const uint4*>(b + r * 64 + ch * 8);
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tptr, 64);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  constexpr uint32_t idesc = make_idesc_f16(128, 64);
  int phase = 0;
  for (int s = 0; s <= 8; ++s) {
    for (int mode = 0; mode < 3; ++mode) {
      const uint32_t bo = mode == 0 ? 0u : mode == 1 ? static_cast<uint32_t>(s & 7) : static_cast<uint32_t>((8 - s) & 7);
      if (threadIdx.x == 0) {
        tc_fence_after();
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = make_desc_sw128_bo(smem_u32(sA) + s * 128 + k * 32, bo);
          const uint64_t db = make_desc_sw128_bo(smem_u32(sB) + k * 32, 0);
          umma_f16(tmem, da, db, idesc, k > 0);
        }
        umma_commit(bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      float* o = out + ((static_cast<size_t>(s) * 3 + mode) * 128 + warp * 32 + lane) * 64;
      for (int c = 0; c < 64; c += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, r);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) o[c + j] = __uint_as_float(r[j]);
      }
      tc_fence_before();
      __syncthreads();
    }
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

int launch_probe_shift(cudaStream_t s, const __half* a, const __half* b, float* out) {
  const int smem = 18432 + 8192 + 64 + 1024;
  if (cudaFuncSetAttribute(probe_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
  probe_shift_kernel<<<1, 128, smem, s>>>(a, b, out);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}


//  probe 1: tensor-pipe issue cost.  `reps` back-to-back tcgen05.mma (M=128, K=16, kind::f16, both operands in shared
//           memory, 128-byte swizzle) per N in {64, 128, 256}, timed with clock64 from issue of the first to the
//           mbarrier completion of the last.  out[i] = cycles per MMA.  Variant b uses a different A tile for every
//           other MMA.  One CTA, so no bandwidth contention: this is the per-SM floor for SS-mode MMAs.
__global__ void __launch_bounds__(128, 1) probe_mma_rate_kernel(float* __restrict__ out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by offsetting the __shared__ array itself (keeps the shared address space: STS/LDS, not generic)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                 // 2 x 16 KB
  uint8_t* sB = smem + 32768;         // 32 KB (256 rows)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(tptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  int phase = 0;
  int slot = 0;
  for (int variant = 0; variant < 2; ++variant) {
    for (int ni = 0; ni < 3; ++ni) {
      const int n = 64 << ni;
      const uint32_t idesc = make_idesc_f16(128, n);
      long long t0 = 0, t1 = 0;
      if (threadIdx.x == 0) {
        t0 = clock64();
        for (int i = 0; i < reps; ++i) {
          const uint32_t a = smem_u32(sA) + ((variant && (i & 1)) ? 16384 : 0) + (i & 3) * 32;
          umma_f16(tmem, make_sw128_kmajor_desc(a), make_sw128_kmajor_desc(smem_u32(sB) + (i & 3) * 32), idesc, 1u);
        }
        umma_commit(bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1;
      if (threadIdx.x == 0) {
        t1 = clock64();
        out[slot] = static_cast<float>(t1 - t0) / reps;
      }
      ++slot;
      __syncthreads();
    }
  }
  // row-shifted A start (+128 B, +256 B: the dx taps of the strip convolution), N = 128
  for (int sh = 1; sh <= 2; ++sh) {
    const uint32_t idesc = make_idesc_f16(128, 128);
    long long t0 = 0;
    if (threadIdx.x == 0) {
      t0 = clock64();
      for (int i = 0; i < reps; ++i)
        umma_f16(tmem, make_sw128_kmajor_desc(smem_u32(sA) + sh * 128 + (i & 3) * 32),
                 make_sw128_kmajor_desc(smem_u32(sB) + (i & 3) * 32), idesc, 1u);
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    if (threadIdx.x == 0) out[7 + sh] = static_cast<float>(clock64() - t0) / reps;
    __syncthreads();
  }
  // issue-only cost (no completion wait between): N=64, measure issue loop time
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_f16(128, 64);
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i)
      umma_f16(tmem, make_sw128_kmajor_desc(smem_u32(sA)), make_sw128_kmajor_desc(smem_u32(sB)), idesc, 1u);
    const long long t1 = clock64();
    umma_commit(bar);
    out[6] = static_cast<float>(t1 - t0) / reps;
  }
  mbar_wait(bar, phase);
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int launch_probe_mma_rate(cudaStream_t s, float* out, int reps) {
  const int smem = 65536 + 64 + 1024;
  if (cudaFuncSetAttribute(probe_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
  probe_mma_rate_kernel<<<1, 128, smem, s>>>(out, reps);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

//  probe 2: A operand from TMEM ("TS" mode).  Each thread packs its fp16 A row (64 values -> 32 columns of half2, low
//           half = even k) and writes it with tcgen05.st.32x32b; the MMA for k-step k reads A at column base + 8*k.
//           Checks D = A B^T and times back-to-back TS MMAs for N = 64 / 128 (out[8192 + i]).
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(128, 1) probe_ts_kernel(const __half* __restrict__ a, const __half* __restrict__ b,
                                                          float* __restrict__ out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;                 // 128 rows * 128 B (rows 64.. repeat b)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 16384);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 128 * 8; i += 128) {
    const int r = i >> 3, ch = i & 7;
    const uint32_t addr = smem_u32(sB) + r * 128;
    const int sw = ch ^ ((addr >> 7) & 7);
    *reinterpret_cast<uint4*>(sB + r * 128 + sw * 16) = *reinterpret_cast<const uint4*>(b + (r & 63) * 64 + ch * 8);
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tptr, 256);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  const uint32_t a_col = 128;                     // A lives in columns [128, 160)
  {
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(a + static_cast<size_t>(threadIdx.x) * 64);
    const uint32_t tl = tmem + (static_cast<uint32_t>(warp * 32) << 16) + a_col;
#pragma unroll
    for (int c = 0; c < 32; c += 8) {
      uint32_t r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = arow[c + j];
      tmem_st8(tl + c, r);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  int phase = 0;
  if (threadIdx.x == 0) {
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_f16(128, 64);
    for (int k = 0; k < 4; ++k)
      umma_f16_ts(tmem, tmem + a_col + k * 8, make_sw128_kmajor_desc(smem_u32(sB) + k * 32), idesc, k > 0);
    umma_commit(bar);
  }
  mbar_wait(bar, phase);
  phase ^= 1;
  tc_fence_after();
  {
    float* o = out + static_cast<size_t>(warp * 32 + lane) * 64;
    for (int c = 0; c < 64; c += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, r);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) o[c + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  // timing: TS MMAs, N = 64 and N = 128, then a 2:1 mix SS N=128 + TS N=64 (the attention inner loop)
  for (int v = 0; v < 3; ++v) {
    long long t0 = 0;
    if (threadIdx.x == 0) {
      tc_fence_after();
      constexpr uint32_t id64 = make_idesc_f16(128, 64), id128 = make_idesc_f16(128, 128);
      t0 = clock64();
      for (int i = 0; i < reps; ++i) {
        const uint64_t db = make_sw128_kmajor_desc(smem_u32(sB) + (i & 3) * 32);
        if (v == 0) umma_f16_ts(tmem, tmem + a_col + (i & 3) * 8, db, id64, 1u);
        else if (v == 1) umma_f16_ts(tmem, tmem + a_col + (i & 3) * 8, db, id128, 1u);
        else {
          umma_f16(tmem, make_sw128_kmajor_desc(smem_u32(sB) + (i & 3) * 32), db, id128, 1u);
          umma_f16_ts(tmem + 64, tmem + a_col + (i & 3) * 8, db, id64, 1u);
        }
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    if (threadIdx.x == 0) out[8192 + v] = static_cast<float>(clock64() - t0) / reps;
    __syncthreads();
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

int launch_probe_ts(cudaStream_t s, const __half* a, const __half* b, float* out, int reps) {
  const int smem = 16384 + 64 + 1024;
  probe_ts_kernel<<<1, 128, smem, s>>>(a, b, out, reps);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

//  probe 3: what bounds the softmax role of the attention kernel.  640 threads like attn_kernel (16 worker warps + 4).
//    out[0..2]  cycles per 16 KB drained from TMEM (128 lanes x 32 columns) with 4 / 8 / 16 warps issuing
//               tcgen05.ld.32x32b.x32 back to back, tensor pipe idle
//    out[3]     the same with 16 warps while warp 19 issues N = 128 MMAs back to back; out[4] = cycles per MMA then
//    out[5..8]  cycles per warp-instruction per SM sub-partition (16 warps = 4 per scheduler, 8-way independent chains):
//               ex2.approx.f32, cvt.rn.f16x2.f32 (F2FP), cvt.f32.f16 (HADD2.F32), fma.rn.f32x2
__global__ void __launch_bounds__(640, 1) probe_softmax_role_kernel(float* __restrict__ out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                 // 16 KB
  uint8_t* sB = smem + 16384;         // 16 KB (128 rows)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  long long* tstamp = reinterpret_cast<long long*>(smem + 32768 + 64);   // [2]
  for (int i = threadIdx.x; i < 32768 / 4; i += 640) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 19) {
    tmem_alloc(tptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  const uint32_t tlane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t sink = 0;
  int phase = 0;
  for (int cfg = 0; cfg < 4; ++cfg) {
    const int nw = cfg == 0 ? 4 : cfg == 1 ? 8 : 16;
    const bool with_mma = cfg == 3;
    __syncthreads();
    long long t0 = clock64();
    if (warp < nw) {
      for (int i = 0; i < reps; ++i) {
        uint32_t r[32];
        tmem_ld32(tlane + ((i + (warp >> 2)) & 7) * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) sink ^= r[j];
      }
      long long t1 = clock64();
      if (warp == 0 && lane == 0) tstamp[0] = t1 - t0;
    } else if (warp == 19 && with_mma && lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, 128);
      for (int i = 0; i < reps; ++i)
        umma_f16(tmem + 256, make_sw128_kmajor_desc(smem_u32(sA) + (i & 3) * 32),
                 make_sw128_kmajor_desc(smem_u32(sB) + (i & 3) * 32), idesc, 1u);
      umma_commit(bar);
      mbar_wait(bar, phase);
      tstamp[1] = clock64() - t0;
    }
    if (with_mma) phase ^= 1;
    __syncthreads();
    if (threadIdx.x == 0) {
      // each LDTM.x32 of one warp moves 32 lanes x 32 columns x 4 B = 4 KB; nw warps in parallel -> nw/4 x 16 KB per rep
      out[cfg] = static_cast<float>(tstamp[0]) / reps / (nw / 4);
      if (with_mma) out[4] = static_cast<float>(tstamp[1]) / reps;
    }
  }
  // instruction issue rates, 16 warps, 8 independent chains per thread
  for (int op = 0; op < 4; ++op) {
    __syncthreads();
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = 0.001f * (threadIdx.x + j) - 0.3f;
    long long t0 = clock64();
    if (warp < 16) {
      for (int i = 0; i < reps; ++i) {
        if (op == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
        } else if (op == 1) {
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            uint32_t h;
            asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x[j]), "f"(x[j + 1]));
            uint32_t h2;
            asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(x[j + 1]), "f"(x[j]));
            x[j] = __uint_as_float(h ^ 0x00010001u);
            x[j + 1] = __uint_as_float(h2 ^ 0x00010001u);
          }
        } else if (op == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            unsigned short hs = static_cast<unsigned short>(__float_as_uint(x[j]));
            asm volatile("cvt.f32.f16 %0, %1;" : "=f"(x[j]) : "h"(hs));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            f32x2 v = pk2(x[j], x[j + 1]);
            asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v) : "l"(v));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v) : "l"(v));
            upk2(v, x[j], x[j + 1]);
          }
        }
      }
      long long t1 = clock64();
      if (warp == 0 && lane == 0) tstamp[0] = t1 - t0;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sink ^= __float_as_uint(x[j]);
    __syncthreads();
    // 4 warps per scheduler x 8 instructions per rep (op 3: 4 x 2)
    if (threadIdx.x == 0) out[5 + op] = static_cast<float>(tstamp[0]) / reps / (4 * (op == 3 ? 4 : 8));
  }
  if (sink == 0x12345678u) out[15] = 1.0f;
  __syncthreads();
  if (warp == 19) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int launch_probe_softmax_role(cudaStream_t s, float* out, int reps) {
  const int smem = 32768 + 128 + 1024;
  if (cudaFuncSetAttribute(probe_softmax_role_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
  probe_softmax_role_kernel<<<1, 640, smem, s>>>(out, reps);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace rfe
