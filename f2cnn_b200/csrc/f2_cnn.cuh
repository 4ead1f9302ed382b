// f2_cnn.cuh -- launch interface of the tcgen05 CNN forward kernels (f2_cnn.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace f2 {

cudaError_t launch_umma_selftest(const void* A, int a_rows, const void* B, int N, int K, int shift, int variant, float* D,
                                 int* status, cudaStream_t stream);

}  // namespace f2
