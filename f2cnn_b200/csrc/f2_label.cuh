// f2_label.cuh -- launch interface of the slope-label kernel (f2_label.cu).
#pragma once
#include <cuda_runtime.h>

namespace f2 {

// out[i] = (slope, intercept, r, p) of item i; see f2_label.cu.
cudaError_t launch_label_fit(const double* formant, const long long* first, const int* center, long long n_items,
                             int dots, int step, double* out, cudaStream_t stream);

}  // namespace f2
