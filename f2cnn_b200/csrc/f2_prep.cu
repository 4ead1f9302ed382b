// f2_prep.cu -- per-utterance pre-pass: zero-padded ring, its circular Hilbert transform
// and the edge-injection kernel.
//
// The reference takes one FFT pair per CHANNEL of the filtered signal
// (scripts/processing/EnvelopeExtraction.py:20-36 paddedHilbert -> scipy.signal.hilbert).
// The ring formulation used by the fused kernel needs only ONE Hilbert transform per
// utterance, of the zero-padded INPUT: xi = Im(hilbert(pad(x, N2))), N2 = 2^ceil(log2 n).
// It is computed here with a hand-written FP32 FFT (no cuFFT):
//   real N2-point FFT as a packed complex M = N2/2 point FFT, split radix-2 in shared
//   memory, M <= 4096 in one pass, larger M as a four-step (columns, twiddle, rows) FFT
//   with <= 4096-point legs; untangle, multiply by -i*sgn(k), tangle; inverse FFT.
// The same kernels serve the stand-alone envelope path (rows of an arbitrary matrix).
#include "f2_prep.cuh"

#include <math.h>

namespace f2 {

constexpr int kFftThreads = 256;
constexpr int kFftSmemPts = 8192;  // complex points per CTA
constexpr int kTwLog = 12;         // twiddle table covers legs up to 4096 points

__device__ float2 g_twiddle[1 << (kTwLog - 1)];  // exp(-2*pi*i*j/4096), j < 2048

__global__ void init_twiddle_kernel() {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < (1 << (kTwLog - 1))) {
        double s, c;
        sincospi(-2.0 * (double)j / (double)(1 << kTwLog), &s, &c);
        g_twiddle[j] = make_float2((float)c, (float)s);
    }
}

cudaError_t init_twiddles(cudaStream_t stream) {
    init_twiddle_kernel<<<(1 << (kTwLog - 1)) / 256, 256, 0, stream>>>();
    return cudaGetLastError();
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

template <typename T>
__device__ __forceinline__ float load_sample(const void* base, long long idx) {
    return (float)reinterpret_cast<const T*>(base)[idx];
}

__device__ __forceinline__ float load_wave(const void* base, int dtype, long long idx) {
    switch (dtype) {
        case F2_DT_I16: return load_sample<short>(base, idx);
        case F2_DT_F32: return load_sample<float>(base, idx);
        default: return load_sample<double>(base, idx);
    }
}

// ---- pack: z[m] = (x[2m], x[2m+1]) zero-padded to N2, into Z ------------------------------
__global__ void pack_kernel(PrepParams p) {
    const UttDesc ut = p.utts[blockIdx.x];
    const int M = ut.N2 >> 1;
    float2* z = reinterpret_cast<float2*>(p.Z + ut.ring_off);
    for (int m = blockIdx.y * blockDim.x + threadIdx.x; m < M; m += gridDim.y * blockDim.x) {
        const int t = 2 * m;
        const float a = t < ut.n ? load_wave(p.wave, p.wave_dtype, ut.wave_off + t) : 0.f;
        const float b = t + 1 < ut.n ? load_wave(p.wave, p.wave_dtype, ut.wave_off + t + 1) : 0.f;
        z[m] = make_float2(a, b);
    }
}

// ---- in-shared-memory radix-2 DIT FFT on `batch` arrays of length 2^logL ------------------
// Data already stored bit-reversed.  INV conjugates the twiddles (unnormalised inverse).
template <bool INV>
__device__ __forceinline__ void smem_fft(float2* s, int pitch, int logL, int batch) {
    const int halfL = 1 << (logL - 1);
    const int total = batch << (logL - 1);
    for (int st = 0; st < logL; ++st) {
        const int half = 1 << st;
        for (int bf = threadIdx.x; bf < total; bf += blockDim.x) {
            const int arr = bf >> (logL - 1);
            const int j = bf & (halfL - 1);
            const int pos = j & (half - 1);
            const int i0 = ((j >> st) << (st + 1)) + pos;
            float2* a = s + arr * pitch;
            float2 w = g_twiddle[pos << (kTwLog - 1 - st)];
            if (INV) w.y = -w.y;
            const float2 u = a[i0];
            const float2 v = cmul(a[i0 + half], w);
            a[i0] = make_float2(u.x + v.x, u.y + v.y);
            a[i0 + half] = make_float2(u.x - v.x, u.y - v.y);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ int bitrev(int v, int bits) { return bits ? (int)(__brev((unsigned)v) >> (32 - bits)) : 0; }

__host__ __device__ inline void fft_split(int log2M, int& l1, int& l2) {
    if (log2M <= kTwLog) {
        l1 = 0;
        l2 = log2M;
    } else {
        l1 = log2M / 2;
        l2 = log2M - l1;
    }
}

// ---- pass A (columns): M1-point FFTs down the columns of the M1 x M2 matrix, then the
// four-step twiddle w_M^(n2*k1); in place. -------------------------------------------------
template <bool INV>
__global__ void __launch_bounds__(kFftThreads) fft_cols_kernel(const UttDesc* utts, float* buf_base, int buf_is_xz) {
    extern __shared__ float2 s_fft[];
    const UttDesc ut = utts[blockIdx.x];
    const int log2M = ut.log2N2 - 1;
    if (log2M <= kTwLog) return;  // single-pass sizes skip the column pass
    int l1, l2;
    fft_split(log2M, l1, l2);
    const int M1 = 1 << l1, M2 = 1 << l2;
    int B = kFftSmemPts >> l1;
    if (B > M2) B = M2;
    const int c0 = blockIdx.y * B;
    if (c0 >= M2) return;
    float2* z = reinterpret_cast<float2*>(buf_base + (buf_is_xz ? 2 : 1) * ut.ring_off);
    const int pitch = M1 + 1;
    const int logB = 31 - __clz(B);
    for (int idx = threadIdx.x; idx < (B << l1); idx += blockDim.x) {
        const int n1 = idx >> logB, b = idx & (B - 1);
        s_fft[b * pitch + bitrev(n1, l1)] = z[(size_t)n1 * M2 + c0 + b];
    }
    __syncthreads();
    smem_fft<INV>(s_fft, pitch, l1, B);
    const float sgn = INV ? 2.0f : -2.0f;
    const float invM = 1.0f / (float)(1 << log2M);
    for (int idx = threadIdx.x; idx < (B << l1); idx += blockDim.x) {
        const int k1 = idx >> logB, b = idx & (B - 1);
        const int n2 = c0 + b;
        float sn, cs;
        sincospif(sgn * (float)(n2 * k1) * invM, &sn, &cs);  // n2*k1 < M <= 2^24: exact in float
        z[(size_t)k1 * M2 + n2] = cmul(s_fft[b * pitch + k1], make_float2(cs, sn));
    }
}

// ---- pass B (rows): M2-point FFTs along the rows; output X[k1 + M1*k2] to `out`. ----------
template <bool INV>
__global__ void __launch_bounds__(kFftThreads)
    fft_rows_kernel(const UttDesc* utts, float* in_base, int in_is_xz, float* out_base, int out_is_xz) {
    extern __shared__ float2 s_fft[];
    const UttDesc ut = utts[blockIdx.x];
    const int log2M = ut.log2N2 - 1;
    if (log2M < 1) return;
    int l1, l2;
    fft_split(log2M, l1, l2);
    const int M1 = 1 << l1, M2 = 1 << l2;
    int B = kFftSmemPts >> l2;
    if (B > M1) B = M1;
    const int r0 = blockIdx.y * B;
    if (r0 >= M1) return;
    const float2* zin = reinterpret_cast<const float2*>(in_base + (in_is_xz ? 2 : 1) * ut.ring_off);
    float2* zout = reinterpret_cast<float2*>(out_base + (out_is_xz ? 2 : 1) * ut.ring_off);
    const int pitch = M2 + 1;
    for (int idx = threadIdx.x; idx < (B << l2); idx += blockDim.x) {
        const int b = idx >> l2, n2 = idx & (M2 - 1);
        s_fft[b * pitch + bitrev(n2, l2)] = zin[(size_t)(r0 + b) * M2 + n2];
    }
    __syncthreads();
    smem_fft<INV>(s_fft, pitch, l2, B);
    const int logB = 31 - __clz(B);
    for (int idx = threadIdx.x; idx < (B << l2); idx += blockDim.x) {
        const int k2 = idx >> logB, b = idx & (B - 1);
        zout[(size_t)k2 * M1 + r0 + b] = s_fft[b * pitch + k2];
    }
}

// ---- untangle the packed real FFT, apply the Hilbert multiplier, tangle for the inverse ----
// Z = FFT_M(z).  X[k] = (Z[k]+conj(Z[M-k]))/2 - (i/2) e^{-2 pi i k/N} (Z[k]-conj(Z[M-k])).
// scipy.signal.hilbert keeps h[0]=h[N/2]=1, doubles 0<k<N/2, zeroes the rest: the imaginary
// part of its output has spectrum Y[k] = -i X[k] (0<k<N/2), Y[0]=Y[N/2]=0.  Packed inverse:
// W[k] = (Y[k]+conj(Y[M-k])) + i e^{+2 pi i k/N} (Y[k]-conj(Y[M-k])), scaled by 1/N.
__global__ void hilbert_mask_kernel(const UttDesc* utts, float* buf_base, int buf_is_xz) {
    const UttDesc ut = utts[blockIdx.x];
    const int M = ut.N2 >> 1;
    if (M < 1) return;
    float2* Z = reinterpret_cast<float2*>(buf_base + (buf_is_xz ? 2 : 1) * ut.ring_off);
    const float invN = 1.0f / (float)ut.N2;
    for (int k = blockIdx.y * blockDim.x + threadIdx.x; k <= M / 2; k += gridDim.y * blockDim.x) {
        if (k == 0) {
            Z[0] = make_float2(0.f, 0.f);
            continue;
        }
        const float2 a = Z[k], b = Z[M - k];
        float sn, cs;
        sincospif(-2.0f * (float)k * invN, &sn, &cs);  // e = exp(-2 pi i k / N)
        const float2 e = make_float2(cs, sn);
        // X[k]
        const float2 sp = make_float2(a.x + b.x, a.y - b.y);  // a + conj(b)
        const float2 dm = make_float2(a.x - b.x, a.y + b.y);  // a - conj(b)
        const float2 ed = cmul(e, dm);
        const float2 Xk = make_float2(0.5f * (sp.x + ed.y), 0.5f * (sp.y - ed.x));  // sp/2 - (i/2) ed
        // X[M-k] = conj(sp)/2 - (i/2) e2 * (-conj(dm)),  e2 = exp(-2 pi i (M-k)/N) = -conj(e)
        // => X[M-k] = conj(sp)/2 - (i/2) conj(e) conj(dm) = conj(sp)/2 - (i/2) conj(ed)
        const float2 Xm = make_float2(0.5f * (sp.x - ed.y), 0.5f * (-sp.y - ed.x));
        // Y = -i X
        const float2 Yk = make_float2(Xk.y, -Xk.x);
        const float2 Ym = make_float2(Xm.y, -Xm.x);
        // W[k] = (Yk + conj(Ym)) + i conj(e) (Yk - conj(Ym))
        const float2 s2 = make_float2(Yk.x + Ym.x, Yk.y - Ym.y);
        const float2 d2 = make_float2(Yk.x - Ym.x, Yk.y + Ym.y);
        const float2 ce = make_float2(e.x, -e.y);
        const float2 t2 = cmul(ce, d2);
        const float2 Wk = make_float2((s2.x - t2.y) * invN, (s2.y + t2.x) * invN);
        // W[M-k] = (Ym + conj(Yk)) + i conj(e2) (Ym - conj(Yk)), conj(e2) = -e
        //        = conj(s2) + i (-e)(-conj(d2)) = conj(s2) + i e conj(d2) = conj(s2) + i conj(t2)
        const float2 Wm = make_float2((s2.x + t2.y) * invN, (-s2.y + t2.x) * invN);
        Z[k] = Wk;
        if (k != M - k) Z[M - k] = Wm;
    }
}

// ---- finish: interleave (x, xi) into the xz ring and tabulate the injection kernel G -------
// G[tau] = h[l], l = the odd one of (tau-n) mod N2, (tau-n-1) mod N2,
// h[l] = (2/N2) cot(pi l / N2): the circular Hilbert kernel for even N2 that matches scipy's
// one-sided mask.
__global__ void finish_kernel(PrepParams p) {
    const UttDesc ut = p.utts[blockIdx.x];
    const int N2 = ut.N2;
    const float* xi = p.Z + ut.ring_off;
    float2* xz = p.xz + ut.ring_off;
    float* G = p.G ? p.G + ut.ring_off : nullptr;
    const float invN = 1.0f / (float)N2;
    for (int tau = blockIdx.y * blockDim.x + threadIdx.x; tau < N2; tau += gridDim.y * blockDim.x) {
        const float x = tau < ut.n ? load_wave(p.wave, p.wave_dtype, ut.wave_off + tau) : 0.f;
        const float im = (p.hilbert && N2 > 2) ? xi[tau] : 0.f;
        xz[tau] = make_float2(x, im);
        if (G) {
            float g = 0.f;
            if (p.hilbert && N2 > 2) {
                int l = (tau - ut.n) & (N2 - 1);
                if (!(l & 1)) l = (l - 1) & (N2 - 1);
                if (l > N2 / 2) l -= N2;  // cot is odd and pi-periodic: keep |l| <= N2/2 for accuracy
                float sn, cs;
                sincospif((float)l * invN, &sn, &cs);
                g = 2.0f * invN * cs / sn;
            }
            G[tau] = g;
        }
    }
}

static int max_blocks(const HostPrepInfo& h, bool cols) {
    int best = 1;
    for (int lg = h.min_log2N2; lg <= h.max_log2N2; ++lg) {
        const int log2M = lg - 1;
        if (log2M < 1) continue;
        int l1, l2;
        fft_split(log2M, l1, l2);
        int blocks;
        if (cols) {
            if (l1 == 0) continue;
            int B = kFftSmemPts >> l1;
            if (B > (1 << l2)) B = 1 << l2;
            blocks = (1 << l2) / B;
        } else {
            int B = kFftSmemPts >> l2;
            if (B > (1 << l1)) B = 1 << l1;
            blocks = (1 << l1) / B;
        }
        if (blocks > best) best = blocks;
    }
    return best;
}

cudaError_t launch_prep(const PrepParams& p, const HostPrepInfo& h, cudaStream_t stream) {
    if (h.n_utts <= 0) return cudaSuccess;
    if (h.max_log2N2 - 1 > 2 * kTwLog) return cudaErrorInvalidValue;
    static bool attr_done = false;
    const int smem = (kFftSmemPts + 2 * 64 + 64) * (int)sizeof(float2);
    if (!attr_done) {
        cudaFuncSetAttribute(fft_cols_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(fft_cols_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(fft_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(fft_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_done = true;
    }
    const int maxN2 = 1 << h.max_log2N2;
    const int ew_blocks = (maxN2 / 2 + 1023) / 1024 > 0 ? (maxN2 / 2 + 1023) / 1024 : 1;
    dim3 g_ew(h.n_utts, ew_blocks > 2048 ? 2048 : ew_blocks);
    float* Zf = p.Z;
    float* XZf = reinterpret_cast<float*>(p.xz);
    if (p.hilbert && h.max_log2N2 >= 2) {
        pack_kernel<<<g_ew, 256, 0, stream>>>(p);
        const bool two = (h.max_log2N2 - 1) > kTwLog;
        dim3 g_cols(h.n_utts, max_blocks(h, true));
        dim3 g_rows(h.n_utts, max_blocks(h, false));
        // forward: Z -> (cols in place) -> rows -> XZ (used as scratch)
        if (two) fft_cols_kernel<false><<<g_cols, kFftThreads, smem, stream>>>(p.utts, Zf, 0);
        fft_rows_kernel<false><<<g_rows, kFftThreads, smem, stream>>>(p.utts, Zf, 0, XZf, 1);
        hilbert_mask_kernel<<<g_ew, 256, 0, stream>>>(p.utts, XZf, 1);
        // inverse: XZ -> (cols in place) -> rows -> Z
        if (two) fft_cols_kernel<true><<<g_cols, kFftThreads, smem, stream>>>(p.utts, XZf, 1);
        fft_rows_kernel<true><<<g_rows, kFftThreads, smem, stream>>>(p.utts, XZf, 1, Zf, 0);
    }
    dim3 g_fin(h.n_utts, ew_blocks > 2048 ? 2048 : ew_blocks);
    finish_kernel<<<g_fin, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace f2
