// C-ABI shim over the REFERENCE's only CUDA kernel, compiled by nvcc from the source where it lies:
//   /root/reference/src/DescriptorGather.cu   gather_normalize_kernel + launch_gather_descriptors (:14-82)
// Built by oracle/Makefile into oracle/_ref/libref_gather.so (sm_100a, static CUDA runtime).  TEST INFRASTRUCTURE:
// tests/test_gpu_zz_ref_gather.py runs the real kernel on the product's own descriptor grid and compares the rows
// bit for bit with what the product's fused gather wrote.
#include <cuda_runtime.h>

#include "DescriptorGather.h"

extern "C" int ref_launch_gather(const void* grid_fp16_chw, int channels, int grid_h, int grid_w, const int* cell_h,
                                 const int* cell_w, int num_keypoints, void* out_fp16) {
  superslam::launch_gather_descriptors(grid_fp16_chw, channels, grid_h, grid_w, cell_h, cell_w, num_keypoints, out_fp16,
                                       nullptr);
  const cudaError_t launch = cudaGetLastError();
  const cudaError_t sync = cudaDeviceSynchronize();
  return launch != cudaSuccess ? static_cast<int>(launch) : static_cast<int>(sync);
}
