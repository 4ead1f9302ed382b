// f2_lanes.cu -- the decimated-output hot path with warp-uniform coefficients.
//
// Same arithmetic as f2_fused.cu (see its header), different mapping.  Measured on B200
// (tools/rf_probe.cu): an FFMA2 whose three sources are distinct vector registers is capped
// by register-file operand bandwidth at 66-76 % of the FMA-pipe rate, while an FFMA2 that
// reads two register pairs plus a UNIFORM-register operand runs at 99 %.  With one thread per
// channel (f2_fused.cu) the coefficients are per-lane and every FFMA2 pays that tax.  Here
//   lane  = one stream (an utterance, or a time chunk of one),
//   warp  = one channel, whose 13 coefficients sit in uniform registers (loaded from
//           __constant__ memory with an index built from blockIdx and a compile-time warp
//           number -- the only form ptxas keeps on the uniform datapath),
//   CTA   = kLaneWarps channels x 32 streams + one producer warp.
// The producer warp streams tiles of 32 samples of each lane's (x, xi) ring and injection
// kernel into shared memory with per-lane 1-D TMA bulk copies; full/empty mbarriers make
// a kLaneStages-deep pipeline, so consumer warps never meet at a CTA-wide barrier.  Rows are
// padded by 16 bytes so that the per-lane 128-bit shared loads are conflict-free.
//
// The edge residuals e_k (zero-padded ring equation at positions n, n+1) are computed once per
// (utterance, channel) by edge_kernel and read back by every stream of that utterance.
#include "f2_lanes.cuh"

#include <stdlib.h>

namespace f2 {

#ifndef F2_LANE_UNROLL
#define F2_LANE_UNROLL 8
#endif
constexpr int kLaneU = F2_LANE_UNROLL;  // samples per inner-loop iteration (code size vs loop overhead)
constexpr int kLaneTile = 32;     // samples per tile
constexpr int kLaneStages = 4;    // pipeline depth
constexpr int kXzPitch = kLaneTile * 8 + 16;  // bytes per lane row of (x, xi)
constexpr int kGPitch = kLaneTile * 4 + 16;   // bytes per lane row of G
constexpr int kStageBytes = 32 * (kXzPitch + kGPitch);

__constant__ float c_par[13 * kMaxConstChan];  // [parameter][channel], see ChanPar

cudaError_t upload_lane_constants(const float* host_par, int c_pad, cudaStream_t stream) {
    // host_par is [kNumChanPar][c_pad]; constant layout is [13][kMaxConstChan]
    for (int k = 0; k < 13; ++k) {
        cudaError_t e = cudaMemcpyToSymbolAsync(c_par, host_par + (size_t)k * c_pad, sizeof(float) * c_pad,
                                                sizeof(float) * (size_t)k * kMaxConstChan, cudaMemcpyHostToDevice,
                                                stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- lane kernel ---------------------------------------------------------------------------
struct CoefU {  // warp-uniform
    float z[4], cq[4], ncy[4], g4;
};

struct LState {
    float2 y[4], q[4];
    float2 up;
    float l, eprev;
};

__device__ __forceinline__ float2 cascade_u(const CoefU& k, LState& s, float2 u, float g, const float (&e)[4]) {
    float2 up = s.up;
    s.up = u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 in = __ffma2_rn(make_float2(k.z[i], k.z[i]), up, u);
        in.y = fmaf(e[i], g, in.y);
        const float2 yo = s.y[i];
        float2 qn = __ffma2_rn(make_float2(k.cq[i], k.cq[i]), s.q[i], in);
        qn = __ffma2_rn(make_float2(k.ncy[i], k.ncy[i]), yo, qn);
        const float2 yn = __fadd2_rn(yo, qn);
        s.q[i] = qn;
        s.y[i] = yn;
        up = yo;
        u = yn;
    }
    return u;
}

struct LaneOut {
    float* dec;
    int next_dec;
    int t1;
    int step;
    size_t C;
    float scale;
};

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}

// One specialised copy per consumer warp so that the channel index -- blockIdx.x * kLaneWarps
// + W -- is uniform and the coefficients come from uniform registers.  __noinline__ on purpose:
// inlined into a switch over the warp index the copies are merged back into one body with a
// per-thread channel index and the coefficients land in vector registers.  The copies share
// the instruction cache, so each is kept to ONE small loop: every tile runs the same body (x
// masked to zero before t = 0, envelope and low-pass always computed, stores predicated on
// being past the warm-up).  All arguments are scalars (registers), shared memory is addressed
// with 32-bit shared-window addresses.
template <int W, bool LPF>
__device__ __noinline__ void lane_consumer(const uint32_t smem, const uint32_t full, const uint32_t empty,
                                           const float* __restrict__ edge_utt, float* __restrict__ dec_base,
                                           const int ts, const int t1, const int M, const int mB, const int step,
                                           const int dec0, const int C, const float lp_k, const float lp_b0) {
    const int lane = threadIdx.x & 31;
    const int ch = blockIdx.x * kLaneWarps + W;
    if (ch >= C) {
        // nothing to compute for a channel past the end, but the pipeline still has to turn
        for (int m = 0; m < M; ++m) {
            const int b = m % kLaneStages;
            mbar_wait_a(full + 8 * b, (uint32_t)((m / kLaneStages) & 1));
            __syncwarp();
            if (lane == 0) mbar_arrive_a(empty + 8 * b);
        }
        return;
    }
    CoefU k;
    k.g4 = c_par[P_G4 * kMaxConstChan + ch];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        k.z[i] = c_par[(P_Z + i) * kMaxConstChan + ch];
        k.cq[i] = c_par[(P_CQ + i) * kMaxConstChan + ch];
        k.ncy[i] = c_par[(P_NCY + i) * kMaxConstChan + ch];
    }
    float ee[4], eo[4];
    {
        const float4* e = reinterpret_cast<const float4*>(edge_utt + (size_t)ch * 8);
        const float4 v0 = __ldg(e), v1 = __ldg(e + 1);
        ee[0] = v0.x; ee[1] = v0.y; ee[2] = v0.z; ee[3] = v0.w;
        eo[0] = v1.x; eo[1] = v1.y; eo[2] = v1.z; eo[3] = v1.w;
    }
    const float scale = LPF ? k.g4 * lp_b0 : k.g4;
    int next_dec = 1 << 30;  // armed when the warm-up is over
    float* dec = dec_base + ch;
    LState s;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s.y[i] = make_float2(0.f, 0.f);
        s.q[i] = make_float2(0.f, 0.f);
    }
    s.up = make_float2(0.f, 0.f);
    s.l = 0.f;
    s.eprev = 0.f;
    const uint32_t row_x = smem + lane * kXzPitch;
    const uint32_t row_g = smem + 32 * kXzPitch + lane * kGPitch;

    for (int m = 0; m < M; ++m) {
        const int b = m % kLaneStages;
        mbar_wait_a(full + 8 * b, (uint32_t)((m / kLaneStages) & 1));
        const int t = ts + m * kLaneTile;
        const uint32_t sx = row_x + b * kStageBytes;
        const uint32_t sg = row_g + b * kStageBytes;
        const float xm = t >= 0 ? 1.f : 0.f;  // the real path sees zeros before the first sample
        if (t == 0) {                         // the reference's low-pass starts from zero state at t = 0
            s.l = 0.f;
            s.eprev = 0.f;
        }
        if (m == mB) next_dec = dec0;
#pragma unroll 1
        for (int i = 0; i < kLaneTile; i += kLaneU) {
            float xv[2 * kLaneU], gv[kLaneU];
#pragma unroll
            for (int j = 0; j < kLaneU / 2; ++j) {
                const float4 v = lds128(sx + 8 * i + 16 * j);
                xv[4 * j + 0] = v.x;
                xv[4 * j + 1] = v.y;
                xv[4 * j + 2] = v.z;
                xv[4 * j + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < kLaneU / 4; ++j) {
                const float4 v = lds128(sg + 4 * i + 16 * j);
                gv[4 * j + 0] = v.x;
                gv[4 * j + 1] = v.y;
                gv[4 * j + 2] = v.z;
                gv[4 * j + 3] = v.w;
            }
            float ev[kLaneU];
#pragma unroll
            for (int j = 0; j < kLaneU; ++j) {
                const float2 u = make_float2(xm * xv[2 * j], xv[2 * j + 1]);
                const float2 y = (j & 1) ? cascade_u(k, s, u, gv[j], eo) : cascade_u(k, s, u, gv[j], ee);
                float e = fast_sqrt(fmaf(y.x, y.x, y.y * y.y));
                if (LPF) {
                    s.l = fmaf(lp_k, s.l, e + s.eprev);
                    s.eprev = e;
                    e = s.l;
                }
                ev[j] = e;
            }
            while (next_dec < t + i + kLaneU) {
                const int r = next_dec - (t + i);
                float v = ev[0];
#pragma unroll
                for (int j = 1; j < kLaneU; ++j) v = (r == j) ? ev[j] : v;
                if (next_dec < t1) __stcs(dec, scale * v);
                dec += C;
                next_dec += step;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_a(empty + 8 * b);
    }
}

__global__ void __launch_bounds__((kLaneWarps + 1) * 32, 2) lane_kernel(const LaneParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_full[kLaneStages];
    __shared__ __align__(8) uint64_t s_empty[kLaneStages];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sidx = blockIdx.y * 32 + lane;
    const bool lane_valid = sidx < p.n_streams;
    LaneStream st = p.streams[lane_valid ? sidx : 0];
    if (!lane_valid) st.t1 = st.t0;  // zero-length stream on valid memory
    const UttDesc ut = p.utts[st.utt];
    const LaneGroup grp = p.groups[blockIdx.y];
    const int w_pre = grp.mB * kLaneTile;
    int len = st.t1 - st.t0;
#pragma unroll
    for (int off = 16; off; off >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, off));
    const int M = (w_pre + len + kLaneTile - 1) / kLaneTile;

    if (threadIdx.x == 0) {
        for (int b = 0; b < kLaneStages; ++b) {
            mbar_init(&s_full[b], 1);
            mbar_init(&s_empty[b], kLaneWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kLaneWarps) {
        // ---- producer: per-lane TMA bulk copies of this lane's ring tiles ----
        const float2* xz_ring = p.xz + ut.ring_off;
        const float* g_ring = p.G + ut.ring_off;
        const int mask = ut.N2 - 1;
        const int ts = st.t0 - w_pre;
        for (int m = 0; m < M; ++m) {
            const int b = m % kLaneStages;
            if (m >= kLaneStages) mbar_wait(&s_empty[b], (uint32_t)(((m / kLaneStages) - 1) & 1));
            const int tau = (ts + m * kLaneTile) & mask;
            if (lane == 0) mbar_expect_tx(&s_full[b], 32 * kLaneTile * 12);
            __syncwarp();
            tma_load_1d(smem + b * kStageBytes + lane * kXzPitch, xz_ring + tau, kLaneTile * 8, &s_full[b]);
            tma_load_1d(smem + b * kStageBytes + 32 * kXzPitch + lane * kGPitch, g_ring + tau, kLaneTile * 4,
                        &s_full[b]);
        }
        return;
    }
    // ---- consumers: one specialised copy per warp so that the channel index is uniform ----
    // per-lane scalars for the consumer
    const int ts = st.t0 - w_pre;
    const int t1 = lane_valid ? st.t1 : st.t0;
    int j0 = 0;
    if (st.t0 > p.phase) j0 = (st.t0 - p.phase + p.step - 1) / p.step;
    const int dec0 = p.phase + j0 * p.step;
    float* dec_base = p.dec + (size_t)(ut.dec_off + j0) * (size_t)p.C;
    const float* edge_utt = p.edge + (size_t)st.utt * (size_t)p.C * 8;
    const uint32_t sm = smem_u32(smem), fu = smem_u32(s_full), em = smem_u32(s_empty);
#define F2_LANE_CASE(W)                                                                                          \
    case W:                                                                                                      \
        if (p.lpf)                                                                                               \
            lane_consumer<W, true>(sm, fu, em, edge_utt, dec_base, ts, t1, M, grp.mB, p.step, dec0, p.C, p.lp_k, \
                                   p.lp_b0);                                                                     \
        else                                                                                                     \
            lane_consumer<W, false>(sm, fu, em, edge_utt, dec_base, ts, t1, M, grp.mB, p.step, dec0, p.C, p.lp_k, \
                                    p.lp_b0);                                                                    \
        break;
    switch (warp) {
        F2_LANE_CASE(0)
        F2_LANE_CASE(1)
        F2_LANE_CASE(2)
        F2_LANE_CASE(3)
#if F2_LANE_WARPS > 4
        F2_LANE_CASE(4)
        F2_LANE_CASE(5)
        F2_LANE_CASE(6)
        F2_LANE_CASE(7)
#endif
        default: break;
    }
#undef F2_LANE_CASE
}

int lane_tile_samples() { return kLaneTile; }

cudaError_t launch_lanes(const LaneParams& p, int n_groups, cudaStream_t stream) {
    if (n_groups <= 0) return cudaSuccess;
    static bool attr_done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    const int smem = kLaneStages * kStageBytes;
    if (dev >= 64 || !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        if (dev < 64) attr_done[dev] = true;
    }
    dim3 grid((p.C + kLaneWarps - 1) / kLaneWarps, n_groups);  // x: the channel blocks of one group run together (L2)
    lane_kernel<<<grid, (kLaneWarps + 1) * 32, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace f2
