#include "tensormap.h"

#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>

#include <string.h>

#include <mutex>
#include <unordered_map>

namespace rfe {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::once_flag g_once;

static void resolve() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
}

int make_tmap_f16_sw128(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tmap(out, 0, 3, base, rank, dims, strides_bytes, box);
}

namespace {
struct MapKey {
  const void* base;
  int dtype, swizzle, rank;
  uint64_t dims[5], strides[4];
  uint32_t box[5];
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return static_cast<size_t>(h);
  }
};
thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_cache;
}  // namespace

int make_tmap(CUtensorMap* out, int dtype, int swizzle, const void* base, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box) {
  std::call_once(g_once, resolve);
  if (!g_encode) {
    set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return 1;
  }
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base;
  key.dtype = dtype;
  key.swizzle = swizzle;
  key.rank = rank;
  for (int i = 0; i < rank; ++i) {
    key.dims[i] = dims[i];
    key.box[i] = box[i];
    if (i + 1 < rank) key.strides[i] = strides_bytes[i];
  }
  auto it = g_cache.find(key);
  if (it != g_cache.end()) {
    *out = it->second;
    return 0;
  }
  if (g_cache.size() > 4096) g_cache.clear();
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  static const CUtensorMapSwizzle kSw[4] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_SWIZZLE_64B,
                                            CU_TENSOR_MAP_SWIZZLE_128B};
  CUresult r = g_encode(out, dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                        static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bdim, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, kSw[swizzle & 3], CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return 1;
  }
  g_cache.emplace(key, *out);
  return 0;
}

}  // namespace rfe
