// SSBW tensor archive reader (format: superslam_b200/weights_io.py).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace ssb {

struct HostTensor {
  std::vector<int> dims;
  std::vector<float> data;
  size_t numel() const { return data.size(); }
};

struct WeightArchive {
  std::map<std::string, HostTensor> tensors;
  // Returns nullptr (and sets last_error) when missing or when the shape differs.
  const HostTensor* get(const std::string& name, std::initializer_list<int> dims) const;
  bool has(const std::string& name) const { return tensors.count(name) != 0; }
};

// 0 on success; SSB_ERR_IO otherwise (last_error set).
int load_archive(const char* path, WeightArchive* out);

}  // namespace ssb
