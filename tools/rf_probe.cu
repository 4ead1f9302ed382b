// rf_probe.cu -- does register-file operand bandwidth cap packed FP32 throughput?
// Each test issues 8 independent FMA-pipe instructions per round whose source operands are
// all distinct registers (no operand-reuse-cache hits), and reports achieved lane-FMAs as a
// fraction of 148 SM x 128 lanes x clock.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

template <int T>
__global__ void __launch_bounds__(256) probe(float* out, int iters, const float* in) {
    float2 a[8], b[8], c[8];
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = make_float2(in[i] + threadIdx.x, in[8 + i]);
        b[i] = make_float2(in[16 + i], in[24 + i] + 1e-9f * threadIdx.x);
        c[i] = make_float2(in[32 + i], in[40 + i] + 1e-9f * threadIdx.x);
        s[i] = in[48 + i] + 1e-9f * threadIdx.x;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (T == 1) a[i] = __ffma2_rn(b[i], c[i], a[i]);                       // 3 pairs
                if (T == 2) a[i] = __ffma2_rn(make_float2(s[i], s[i]), c[i], a[i]);    // scalar + 2 pairs
                if (T == 3) { a[i].x = fmaf(b[i].x, c[i].x, a[i].x); a[i].y = fmaf(b[i].y, c[i].y, a[i].y); }  // 3 scalars
                if (T == 4) a[i] = __fadd2_rn(a[i], c[i]);                              // 2 pairs
                if (T == 5) a[i] = __ffma2_rn(make_float2(s[0], s[0]), c[i], a[i]);    // reused scalar + 2 pairs
                if (T == 6) a[i] = __ffma2_rn(make_float2(s[0], s[0]), make_float2(s[1], s[1]), a[i]);  // 2 reused + pair
                if (T == 7) a[i] = __ffma2_rn(make_float2(s[i], s[i]), a[(i + 3) & 7], a[i]);  // kernel-like: scalar, other state, self
                if (T == 8) {  // two independent streams sharing each coefficient back to back (reuse cache?)
                    a[i] = __ffma2_rn(make_float2(s[i], s[i]), c[i], a[i]);
                    b[i] = __ffma2_rn(make_float2(s[i], s[i]), c[(i + 1) & 7], b[i]);
                }
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y + b[i].x + b[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int T>
double run(float* out, const float* in, int blocks, int iters, double lane_fma_per_inst) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<T><<<blocks, 256>>>(out, iters, in);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        probe<T><<<blocks, 256>>>(out, iters, in);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double insts = (double)iters * 32.0 * blocks * 256.0;  // thread-level instructions (T==3 counts 2 per slot)
    return insts * lane_fma_per_inst / (best * 1e-3);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double peak = (double)p.multiProcessorCount * 128.0 * clk_khz * 1e3;  // lane-FMAs per second
    const int blocks = p.multiProcessorCount * 8, iters = 4000;
    float *out, *in;
    cudaMalloc(&out, sizeof(float) * blocks * 256);
    cudaMalloc(&in, sizeof(float) * 64);
    float h[64];
    for (int i = 0; i < 64; ++i) h[i] = 0.5f + 0.001f * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("T1 ffma2 pair,pair,pair        : %.3f of peak\n", run<1>(out, in, blocks, iters, 2.0) / peak);
    printf("T2 ffma2 scalar,pair,pair      : %.3f of peak\n", run<2>(out, in, blocks, iters, 2.0) / peak);
    printf("T3 ffma  3 distinct scalars    : %.3f of peak\n", run<3>(out, in, blocks, iters, 2.0) / peak);
    printf("T4 fadd2 pair,pair             : %.3f of peak\n", run<4>(out, in, blocks, iters, 2.0) / peak);
    printf("T5 ffma2 reused scalar,pair,pair: %.3f of peak\n", run<5>(out, in, blocks, iters, 2.0) / peak);
    printf("T6 ffma2 2 reused scalars,pair : %.3f of peak\n", run<6>(out, in, blocks, iters, 2.0) / peak);
    printf("T7 ffma2 scalar,state,self     : %.3f of peak\n", run<7>(out, in, blocks, iters, 2.0) / peak);
    printf("T8 2 streams sharing the scalar: %.3f of peak\n", run<8>(out, in, blocks, iters, 4.0) / peak);
    return 0;
}
