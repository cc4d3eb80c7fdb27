// Test-side entry point over the library's weight-archive reader (superslam_b200/csrc/weights.cpp), which the C-ABI only
// reaches after a device has been selected: tests/test_weights_reader.py builds this file together with weights.cpp and
// runtime.cu (nvcc, no GPU needed) and feeds the reader valid, truncated and corrupted archives.  TEST INFRASTRUCTURE.
#include <cstring>

#include "../superslam_b200/csrc/common.cuh"
#include "../superslam_b200/csrc/weights.h"

extern "C" int shim_load_archive(const char* path, int* n_tensors, double* checksum, char* err, int err_bytes) {
  ssb::WeightArchive ar;
  int st = SSB_ERR_IO;
  try {
    st = ssb::load_archive(path, &ar);
  } catch (const std::bad_alloc&) {   // what SSB_API_BEGIN / END turn into a status at the boundary
    st = -1;
  }
  *n_tensors = static_cast<int>(ar.tensors.size());
  double s = 0;
  for (const auto& kv : ar.tensors)
    for (float v : kv.second.data) s += v;
  *checksum = s;
  std::strncpy(err, ssb::last_error(), err_bytes - 1);
  err[err_bytes - 1] = 0;
  return st;
}
