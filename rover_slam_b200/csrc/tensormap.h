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

void set_error(const char* fmt, ...);
const char* get_error();

}  // namespace rfe
