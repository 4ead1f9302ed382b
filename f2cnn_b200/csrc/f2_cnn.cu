// f2_cnn.cu -- tcgen05 (5th-generation tensor core) forward pass of the reference's F2-direction CNN
// for `cnn eval*` (scripts/CNN/Training.py:93-114, scripts/CNN/Evaluating.py:70-87): the one dense
// contraction next to the feature-extraction hot path.
//
// See f2_umma.cuh for the operand layout all kernels here share.
#include "f2_cnn.cuh"

#include <cuda_bf16.h>

#include "f2_umma.cuh"

namespace f2 {

// ---- self-test: one 128 x N x K product through the exact descriptor conventions of the CNN kernels ---
// D[r][n] = sum_k A[shift + r][k] * B[n][k], r < 128.  A: [a_rows][K], B: [N][K] row-major bf16 in global
// memory; both are re-laid into K-major planes in shared memory, A's descriptor is advanced by `shift`
// rows (the convolution-tap trick).  variant 1 swaps the leading / stride byte offsets (must be WRONG:
// kept so that the test proves the convention rather than assuming it).
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const __nv_bfloat16* __restrict__ A, int a_rows, const __nv_bfloat16* __restrict__ B, int N, int K,
                     int shift, int variant, float* __restrict__ D, int* __restrict__ status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int planes = K / 8;
    const uint32_t plane_a = (uint32_t)a_rows * 16u, plane_b = (uint32_t)N * 16u;
    uint8_t* sa = smem;
    uint8_t* sb = smem + (size_t)planes * plane_a;
    for (int i = tid; i < a_rows * planes; i += blockDim.x) {
        const int r = i / planes, p = i % planes;
        *reinterpret_cast<uint4*>(sa + (size_t)p * plane_a + (size_t)r * 16) =
            *reinterpret_cast<const uint4*>(A + (size_t)r * K + (size_t)p * 8);
    }
    for (int i = tid; i < N * planes; i += blockDim.x) {
        const int r = i / planes, p = i % planes;
        *reinterpret_cast<uint4*>(sb + (size_t)p * plane_b + (size_t)r * 16) =
            *reinterpret_cast<const uint4*>(B + (size_t)r * K + (size_t)p * 8);
    }
    uint32_t cols = 32;
    while ((int)cols < N) cols <<= 1;
    if (warp == 0) umma::tmem_alloc(&tmem_slot, cols);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = umma::instr_desc_bf16(128, N);
        for (int kk = 0; kk < K / 16; ++kk) {
            const uint32_t a_addr = smem_u32(sa) + (uint32_t)(2 * kk) * plane_a + (uint32_t)shift * 16u;
            const uint32_t b_addr = smem_u32(sb) + (uint32_t)(2 * kk) * plane_b;
            const uint64_t da = variant == 1 ? umma::smem_desc(a_addr, 128, plane_a) : umma::smem_desc(a_addr, plane_a, 128);
            const uint64_t db = variant == 1 ? umma::smem_desc(b_addr, 128, plane_b) : umma::smem_desc(b_addr, plane_b, 128);
            umma::mma_bf16(tmem, da, db, idesc, kk > 0);
        }
        umma::mma_commit(&bar);
    }
    const bool ok = umma::mbar_wait_bounded(&bar, 0);
    umma::fence_after_sync();
    if (!ok) {
        if (tid == 0) *status = 1;  // the MMAs never completed
    } else {
        for (int c0 = 0; c0 < N; c0 += 16) {
            float v[16];
            umma::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = v[j];
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, cols);
}

cudaError_t launch_umma_selftest(const void* A, int a_rows, const void* B, int N, int K, int shift, int variant, float* D,
                                 int* status, cudaStream_t stream) {
    const size_t smem = (size_t)(K / 8) * ((size_t)a_rows + (size_t)N) * 16;
    cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    umma_selftest_kernel<<<1, 128, smem, stream>>>((const __nv_bfloat16*)A, a_rows, (const __nv_bfloat16*)B, N, K, shift,
                                                    variant, D, status);
    return cudaGetLastError();
}

}  // namespace f2
