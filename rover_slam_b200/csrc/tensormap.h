// Host-side CUtensorMap construction (driver entry point fetched through the runtime, so the library
// does not link libcuda directly).
#pragma once

#include <cuda.h>
#include <stdint.h>

namespace rfe {

// fp16 tensor, innermost dim contiguous, 128-byte swizzle, zero fill out of bounds.
// dims[0] is the innermost extent (elements); strides_bytes[i] is the byte stride of dims[i+1].
// Returns 0 on success (error text via rfe::set_error).
int make_tmap_f16_sw128(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box);

// Generic form: dtype 0 = fp16, 1 = fp32; swizzle 0 = none, 1 = 32 B, 2 = 64 B, 3 = 128 B.  Encoded maps are cached per
// thread by (base, geometry): the launch sequences re-use the same few dozen maps every layer / every step.
int make_tmap(CUtensorMap* out, int dtype, int swizzle, const void* base, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box);

void set_error(const char* fmt, ...);
const char* get_error();

}  // namespace rfe
