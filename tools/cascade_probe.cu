// cascade_probe.cu -- would two streams per thread (same channel, shared coefficients) beat one?
// Same arithmetic and dependency structure as the direct-form hot loop of f2_fused.cu (12 FFMA2 +
// 4 injection FFMA + magnitude + one-pole low-pass per channel-sample, inputs broadcast from shared
// memory, one warp per CTA), without the TMA pipeline and the stores.  NS = 1: one stream per
// thread, 16 CTAs per SM.  NS = 2: two streams per thread, interleaved at source level so that
// consecutive FFMA2s share their coefficient register (operand reuse cache), 8 or 16 CTAs per SM.
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int kTile = 256;

struct St {
    float2 y[4], q[4], up;
    float w;
};

template <int NS, int MINB, int U>
__global__ void __launch_bounds__(32, MINB) probe(const float* __restrict__ par, const float* __restrict__ inp,
                                                  float* __restrict__ out, int tiles) {
    __shared__ __align__(16) float2 sxz[NS][kTile];
    __shared__ __align__(16) float sg[NS][kTile];
    for (int i = threadIdx.x; i < kTile; i += 32)
        for (int s = 0; s < NS; ++s) {
            sxz[s][i] = make_float2(inp[(i + 7 * s) & 255], inp[256 + ((i + 11 * s) & 255)]);
            sg[s][i] = 1e-4f * inp[(i + 3 * s) & 255];
        }
    __syncwarp();
    const int c = (blockIdx.x * 32 + threadIdx.x) & 127;
    float z0, cq[4], ncy[4], zn[3], e[NS][2][4];
    z0 = par[c];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        cq[i] = par[128 * (1 + i) + c];
        ncy[i] = par[128 * (5 + i) + c];
        if (i < 3) zn[i] = par[128 * (9 + i) + c];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            e[s][0][i] = 1e-3f * par[c + i + s];
            e[s][1][i] = -1e-3f * par[c + 2 * i + s];
        }
    }
    const float lpk = 0.98f;
    St st[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
#pragma unroll
        for (int i = 0; i < 4; ++i) st[s].y[i] = st[s].q[i] = make_float2(0.f, 0.f);
        st[s].up = make_float2(0.f, 0.f);
        st[s].w = 0.f;
    }
    float sum[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) sum[s] = 0.f;
    for (int t = 0; t < tiles; ++t) {
        for (int i0 = 0; i0 < kTile; i0 += U) {
            float xv[NS][2 * U], gv[NS][U];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
#pragma unroll
                for (int j = 0; j < U / 2; ++j) {
                    const float4 v = *reinterpret_cast<const float4*>(&sxz[s][i0 + 2 * j]);
                    xv[s][4 * j] = v.x; xv[s][4 * j + 1] = v.y; xv[s][4 * j + 2] = v.z; xv[s][4 * j + 3] = v.w;
                }
#pragma unroll
                for (int j = 0; j < U / 4; ++j) {
                    const float4 v = *reinterpret_cast<const float4*>(&sg[s][i0 + 4 * j]);
                    gv[s][4 * j] = v.x; gv[s][4 * j + 1] = v.y; gv[s][4 * j + 2] = v.z; gv[s][4 * j + 3] = v.w;
                }
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                float2 in[NS], yn[NS], acc[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const float2 u = make_float2(xv[s][2 * j], xv[s][2 * j + 1]);
                    in[s] = __ffma2_rn(make_float2(z0, z0), st[s].up, u);
                    st[s].up = u;
                }
#pragma unroll
                for (int s = 0; s < NS; ++s) in[s].y = fmaf(e[s][j & 1][0], gv[s][j], in[s].y);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
#pragma unroll
                    for (int s = 0; s < NS; ++s) acc[s] = __ffma2_rn(make_float2(cq[i], cq[i]), st[s].q[i], in[s]);
#pragma unroll
                    for (int s = 0; s < NS; ++s) yn[s] = __ffma2_rn(make_float2(ncy[i], ncy[i]), st[s].y[i], acc[s]);
                    if (i < 3) {
#pragma unroll
                        for (int s = 0; s < NS; ++s) in[s] = __ffma2_rn(make_float2(zn[i], zn[i]), st[s].y[i], acc[s]);
#pragma unroll
                        for (int s = 0; s < NS; ++s) in[s].y = fmaf(e[s][j & 1][i + 1], gv[s][j], in[s].y);
                    }
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        st[s].q[i] = st[s].y[i];
                        st[s].y[i] = yn[s];
                    }
                }
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    float m;
                    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(m) : "f"(fmaf(yn[s].x, yn[s].x, yn[s].y * yn[s].y)));
                    st[s].w = fmaf(lpk, st[s].w, m);
                }
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) sum[s] += st[s].w;
        }
    }
    float tot = 0.f;
#pragma unroll
    for (int s = 0; s < NS; ++s) tot += sum[s];
    out[blockIdx.x * 32 + threadIdx.x] = tot;
}

template <int NS, int MINB, int U>
void run(const char* name, const float* par, const float* inp, float* out, int sms) {
    const int blocks = sms * MINB * 4, tiles = 64 / NS;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<NS, MINB, U><<<blocks, 32>>>(par, inp, out, tiles);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        probe<NS, MINB, U><<<blocks, 32>>>(par, inp, out, tiles);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, probe<NS, MINB, U>);
    const double cs = (double)blocks * 32.0 * NS * tiles * kTile;
    printf("%-34s regs %3d  %.3f ms  %.3e channel-samples/s  (%s)\n", name, a.numRegs, best, cs / (best * 1e-3),
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    float *par, *inp, *out;
    cudaMalloc(&par, sizeof(float) * 128 * 13);
    cudaMalloc(&inp, sizeof(float) * 512);
    cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 64 * 32 * 4);
    float h[128 * 13], hi[512];
    for (int c = 0; c < 128; ++c) {
        h[c] = -0.5f - 0.001f * c;                                       // z0
        for (int i = 0; i < 4; ++i) {
            h[128 * (1 + i) + c] = -(0.55f + 0.002f * c);                // -B2
            h[128 * (5 + i) + c] = 1.2f + 0.002f * c;                    // -B1
            if (i < 3) h[128 * (9 + i) + c] = 0.6f + 0.001f * c + 0.01f * i;
        }
    }
    for (int i = 0; i < 512; ++i) hi[i] = (float)((i * 7919) % 1000 - 500);
    cudaMemcpy(par, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMemcpy(inp, hi, sizeof(hi), cudaMemcpyHostToDevice);
    run<1, 16, 8>("1 stream, 16 CTAs/SM, U=8", par, inp, out, p.multiProcessorCount);
    run<1, 16, 16>("1 stream, 16 CTAs/SM, U=16", par, inp, out, p.multiProcessorCount);
    run<2, 8, 8>("2 streams, 8 CTAs/SM, U=8", par, inp, out, p.multiProcessorCount);
    run<2, 16, 8>("2 streams, 16 CTAs/SM, U=8", par, inp, out, p.multiProcessorCount);
    run<2, 12, 8>("2 streams, 12 CTAs/SM, U=8", par, inp, out, p.multiProcessorCount);
    run<2, 8, 4>("2 streams, 8 CTAs/SM, U=4", par, inp, out, p.multiProcessorCount);
    run<2, 16, 4>("2 streams, 16 CTAs/SM, U=4", par, inp, out, p.multiProcessorCount);
    return 0;
}
