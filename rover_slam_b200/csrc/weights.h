// RFW1 weight blob reader (host).  The blob is produced by tools/pack_weights.py from the reference's
// onnxmodel/superpoint.onnx and onnxmodel/lightglue_sim.onnx initialisers.
//   header : "RFW1", u32 n_tensors, u64 total_bytes
//   table  : n x 128 B  { char name[80]; u32 ndim; u32 dims[4]; u64 offset; u64 nbytes; pad }
//   data   : fp32 little endian, 256-byte aligned
#pragma once

#include <stdint.h>

#include <map>
#include <string>
#include <vector>

namespace rfe {

struct HostTensor {
  std::vector<int> dims;
  const float* data = nullptr;
  size_t size() const {
    size_t n = 1;
    for (int d : dims) n *= static_cast<size_t>(d);
    return n;
  }
};

class WeightBlob {
 public:
  // Returns 0 on success; error text via rfe::set_error.
  int load(const char* path);
  const HostTensor* find(const std::string& name) const;

 private:
  std::vector<uint8_t> buf_;
  std::map<std::string, HostTensor> tensors_;
};

}  // namespace rfe
