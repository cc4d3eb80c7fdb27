#include "weights.h"

#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace ssb {

const HostTensor* WeightArchive::get(const std::string& name, std::initializer_list<int> dims) const {
  auto it = tensors.find(name);
  if (it == tensors.end()) {
    set_last_error("weight archive: tensor '%s' missing", name.c_str());
    return nullptr;
  }
  std::vector<int> want(dims);
  if (it->second.dims != want) {
    set_last_error("weight archive: tensor '%s' has unexpected shape", name.c_str());
    return nullptr;
  }
  return &it->second;
}

int load_archive(const char* path, WeightArchive* out) {
  FILE* f = path ? std::fopen(path, "rb") : nullptr;
  SSB_CHECK(f != nullptr, SSB_ERR_IO, "cannot open weight archive '%s'", path ? path : "(null)");
  std::vector<uint8_t> buf;
  std::fseek(f, 0, SEEK_END);
  long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  if (sz < 12) {
    std::fclose(f);
    set_last_error("weight archive '%s' too small", path);
    return SSB_ERR_IO;
  }
  buf.resize(static_cast<size_t>(sz));
  size_t got = std::fread(buf.data(), 1, buf.size(), f);
  std::fclose(f);
  SSB_CHECK(got == buf.size(), SSB_ERR_IO, "short read on '%s'", path);
  SSB_CHECK(std::memcmp(buf.data(), "SSBW", 4) == 0, SSB_ERR_IO, "'%s' is not an SSBW archive", path);
  auto rd32 = [&](size_t off, uint32_t* v) -> bool {
    if (off + 4 > buf.size()) return false;
    std::memcpy(v, buf.data() + off, 4);
    return true;
  };
  auto rd64 = [&](size_t off, uint64_t* v) -> bool {
    if (off + 8 > buf.size()) return false;
    std::memcpy(v, buf.data() + off, 8);
    return true;
  };
  uint32_t ver = 0, n = 0;
  rd32(4, &ver);
  rd32(8, &n);
  SSB_CHECK(ver == 1, SSB_ERR_IO, "'%s': unsupported SSBW version %u", path, ver);
  size_t p = 12;
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t ln = 0, dtype = 0, ndim = 0;
    SSB_CHECK(rd32(p, &ln) && p + 4 + ln <= buf.size(), SSB_ERR_IO, "'%s': truncated header", path);
    p += 4;
    std::string name(reinterpret_cast<const char*>(buf.data() + p), ln);
    p += ln;
    SSB_CHECK(rd32(p, &dtype) && rd32(p + 4, &ndim) && ndim <= 8, SSB_ERR_IO, "'%s': bad entry", path);
    p += 8;
    HostTensor t;
    size_t numel = 1;
    for (uint32_t d = 0; d < ndim; ++d) {
      uint32_t v = 0;
      SSB_CHECK(rd32(p, &v), SSB_ERR_IO, "'%s': truncated dims", path);
      p += 4;
      t.dims.push_back(static_cast<int>(v));
      SSB_CHECK(v == 0 || numel <= buf.size() / v, SSB_ERR_IO, "'%s': tensor '%s' is larger than the file", path,
                name.c_str());   // also keeps the product from wrapping around
      numel *= v;
    }
    uint64_t off = 0, size = 0;
    SSB_CHECK(rd64(p, &off) && rd64(p + 8, &size), SSB_ERR_IO, "'%s': truncated entry", path);
    p += 16;
    SSB_CHECK(dtype == 0 && size == numel * 4 && off <= buf.size() && size <= buf.size() - off, SSB_ERR_IO,
              "'%s': tensor '%s' has bad extent", path, name.c_str());
    t.data.resize(numel);
    std::memcpy(t.data.data(), buf.data() + off, size);
    out->tensors.emplace(std::move(name), std::move(t));
  }
  return SSB_OK;
}

}  // namespace ssb
