#include "weights.h"

#include <stdio.h>
#include <string.h>

#include "tensormap.h"

namespace rfe {

int WeightBlob::load(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) {
    set_error("cannot open weight blob '%s'", path);
    return 1;
  }
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (sz < 16) {                       // also a failing ftell (-1): never resize((size_t)-1) across the C ABI
    fclose(f);
    set_error("'%s' is not an RFW1 weight blob (size %ld)", path, sz);
    return 1;
  }
  buf_.resize(static_cast<size_t>(sz));
  const size_t rd = fread(buf_.data(), 1, buf_.size(), f);
  fclose(f);
  if (rd != buf_.size() || sz < 16 || memcmp(buf_.data(), "RFW1", 4) != 0) {
    set_error("'%s' is not an RFW1 weight blob", path);
    return 1;
  }
  uint32_t n;
  memcpy(&n, buf_.data() + 4, 4);
  if (16 + static_cast<size_t>(n) * 128 > buf_.size()) {
    set_error("'%s': truncated table", path);
    return 1;
  }
  for (uint32_t i = 0; i < n; ++i) {
    const uint8_t* e = buf_.data() + 16 + static_cast<size_t>(i) * 128;
    char name[81];
    memcpy(name, e, 80);
    name[80] = 0;
    uint32_t nd, d[4];
    uint64_t off, nb;
    memcpy(&nd, e + 80, 4);
    memcpy(d, e + 84, 16);
    memcpy(&off, e + 100, 8);
    memcpy(&nb, e + 108, 8);
    if (nd > 4 || off > buf_.size() || nb > buf_.size() - off || (off & 3)) {    // overflow-safe bounds
      set_error("'%s': bad entry %s", path, name);
      return 1;
    }
    HostTensor t;
    bool zero_dim = false;
    for (uint32_t k = 0; k < nd; ++k) {
      t.dims.push_back(static_cast<int>(d[k]));
      zero_dim |= d[k] == 0 || d[k] > 0x7fffffffu;
    }
    if (zero_dim) {
      set_error("'%s': zero / oversized dimension in %s", path, name);
      return 1;
    }
    t.data = reinterpret_cast<const float*>(buf_.data() + off);
    if (t.size() * 4 != nb) {
      set_error("'%s': size mismatch in %s", path, name);
      return 1;
    }
    tensors_[name] = t;
  }
  return 0;
}

const HostTensor* WeightBlob::find(const std::string& name) const {
  auto it = tensors_.find(name);
  return it == tensors_.end() ? nullptr : &it->second;
}

}  // namespace rfe
