/*
 * f2_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity checker, never the product).
 *
 * A plain-C float64 restatement of the F2CNN feature-extraction hot path, one
 * function per reference computation.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The
 * product (f2cnn_b200/) never links, imports or calls it.
 *
 * The arithmetic of the reference path lives in two third-party packages that are not
 * vendored under /root/reference: SciPy (README pins scipy 1.1.0; this container has
 * 1.18.1) and NumPy.  Their published algorithms are restated here:
 *   scipy.signal.lfilter  -> Direct-Form-II-transposed IIR        (f2o_lfilter)
 *   scipy.signal.hilbert  -> FFT analytic signal, one-sided mask  (f2o_hilbert_pow2)
 *   scipy.signal.butter   -> bilinear 1st-order Butterworth       (f2o_butter1_lowpass)
 * The restatement is pinned against the reference itself (imported from
 * /root/reference in the build container) by oracle/make_golden.py; the resulting
 * fixtures live in tests/golden/ and are checked by tests/test_oracle_golden.py.
 *
 * All file:line citations are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define EAR_Q 9.26449 /* gammatone/filters.py:35,136 */
#define MIN_BW 24.7   /* gammatone/filters.py:36,137 */

/* ---- gammatone/filters.py:21-52 erb_point, :55-71 erb_space, :74-86 centre_freqs ---- */
void f2o_erb_space(double low_freq, double high_freq, int num, double *out)
{
    const double c = EAR_Q * MIN_BW;
    for (int i = 1; i <= num; ++i) {
        double fraction = (double)i / (double)num; /* np.arange(1,num+1)/num */
        out[i - 1] = -c + exp(fraction * (-log(high_freq + c) + log(low_freq + c))) * (high_freq + c);
    }
}

void f2o_centre_freqs(double fs, int num_freqs, double cutoff, double *out)
{
    f2o_erb_space(cutoff, fs / 2.0, num_freqs, out);
}

/* ---- gammatone/filters.py:89-192 make_erb_filters ----
 * out is (C,10) row-major: [A0, A11, A12, A13, A14, A2, B0, B1, B2, gain] (:186-190). */
typedef struct { double re, im; } cplx;
static cplx c_mul(cplx a, cplx b) { cplx r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return r; }
static cplx c_sub(cplx a, cplx b) { cplx r = { a.re - b.re, a.im - b.im }; return r; }
static cplx c_scale(cplx a, double s) { cplx r = { a.re * s, a.im * s }; return r; }
static cplx c_div(cplx a, cplx b)
{
    double d = b.re * b.re + b.im * b.im;
    cplx r = { (a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d };
    return r;
}

void f2o_make_erb_filters(double fs, const double *centre_freqs, int C, double width, double *out)
{
    const double T = 1.0 / fs;
    const double rt_pos = sqrt(3.0 + pow(2.0, 1.5)); /* :154 */
    const double rt_neg = sqrt(3.0 - pow(2.0, 1.5)); /* :155 */
    for (int i = 0; i < C; ++i) {
        double cf = centre_freqs[i];
        double erb = width * (cf / EAR_Q + MIN_BW); /* :142 with order == 1 */
        double B = 1.019 * 2.0 * M_PI * erb;         /* :143 */
        double arg = 2.0 * cf * M_PI * T;            /* :145 */
        cplx vec = { cos(2.0 * arg), sin(2.0 * arg) }; /* :146 exp(2j*arg) */
        double eBT = exp(B * T);
        double B1 = -2.0 * cos(arg) / eBT;           /* :151 */
        double B2 = exp(-2.0 * B * T);               /* :152 */
        double common = -T * exp(-(B * T));          /* :157 */
        double k[4];
        k[0] = cos(arg) + rt_pos * sin(arg);         /* :162-165 */
        k[1] = cos(arg) - rt_pos * sin(arg);
        k[2] = cos(arg) + rt_neg * sin(arg);
        k[3] = cos(arg) - rt_neg * sin(arg);
        double e = exp(-B * T);
        cplx gain_arg = { e * cos(arg), e * sin(arg) }; /* :172 exp(1j*arg - B*T) */
        cplx prod = { 1.0, 0.0 };
        for (int j = 0; j < 4; ++j)
            prod = c_mul(prod, c_sub(vec, c_scale(gain_arg, k[j]))); /* :175-178 */
        /* (T*exp(BT) / (-1/exp(BT) + 1 + vec*(1-exp(BT))))**4   :179-181 */
        cplx den = { -1.0 / eBT + 1.0 + vec.re * (1.0 - eBT), vec.im * (1.0 - eBT) };
        cplx num = { T * eBT, 0.0 };
        cplx q = c_div(num, den);
        cplx q2 = c_mul(q, q);
        cplx q4 = c_mul(q2, q2);
        cplx g = c_mul(prod, q4);
        double *row = out + (size_t)i * 10;
        row[0] = T;              /* A0 :148 */
        row[1] = common * k[0];  /* A11 :167 */
        row[2] = common * k[1];
        row[3] = common * k[2];
        row[4] = common * k[3];
        row[5] = 0.0;            /* A2 :149 */
        row[6] = 1.0;            /* B0 :150 */
        row[7] = B1;
        row[8] = B2;
        row[9] = hypot(g.re, g.im); /* gain :174 */
    }
}

/* ---- scipy.signal.lfilter (Direct Form II transposed), order <= 2 ----
 * b = numerator (3 taps), a = denominator (3 taps), zero initial state.
 * Same recurrence as scipy/_sigtools `_linear_filter`: coefficients are first
 * normalised by a[0]; y = z0 + b0*x; z0 = z1 + b1*x - a1*y; z1 = b2*x - a2*y. */
void f2o_lfilter3(const double *b, const double *a, const double *x, double *y, int64_t n)
{
    double a0 = a[0];
    double b0 = b[0] / a0, b1 = b[1] / a0, b2 = b[2] / a0, a1 = a[1] / a0, a2 = a[2] / a0;
    double z0 = 0.0, z1 = 0.0;
    for (int64_t t = 0; t < n; ++t) {
        double xt = x[t];
        double yt = z0 + b0 * xt;
        z0 = z1 + b1 * xt - a1 * yt;
        z1 = b2 * xt - a2 * yt;
        y[t] = yt;
    }
}

/* 1st-order variant (lowPassFilter's butter(1) section). */
void f2o_lfilter2(const double *b, const double *a, const double *x, double *y, int64_t n)
{
    double a0 = a[0];
    double b0 = b[0] / a0, b1 = b[1] / a0, a1 = a[1] / a0;
    double z0 = 0.0;
    for (int64_t t = 0; t < n; ++t) {
        double xt = x[t];
        double yt = z0 + b0 * xt;
        z0 = b1 * xt - a1 * yt;
        y[t] = yt;
    }
}

/* ---- gammatone/filters.py:195-239 erb_filterbank ----
 * wave (n,) float64 (the reference passes int16 or float64; lfilter promotes to
 * float64), coefs (C,10) -> out (C,n) float64 row-major.  Four cascaded sections
 * per channel, numerators (A0, A1k, A2) over the shared denominator (B0,B1,B2)
 * (:233-236), divided by gain once at the end (:237). */
void f2o_erb_filterbank(const double *wave, int64_t n, const double *coefs, int C, double *out)
{
#pragma omp parallel
    {
        double *y1 = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        double *y2 = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
#pragma omp for schedule(dynamic, 1)
        for (int c = 0; c < C; ++c) {
            const double *k = coefs + (size_t)c * 10;
            double Bs[3] = { k[6], k[7], k[8] };
            double As1[3] = { k[0], k[1], k[5] }, As2[3] = { k[0], k[2], k[5] };
            double As3[3] = { k[0], k[3], k[5] }, As4[3] = { k[0], k[4], k[5] };
            double *o = out + (size_t)c * (size_t)n;
            f2o_lfilter3(As1, Bs, wave, y1, n);
            f2o_lfilter3(As2, Bs, y1, y2, n);
            f2o_lfilter3(As3, Bs, y2, y1, n);
            f2o_lfilter3(As4, Bs, y1, y2, n);
            double gain = k[9];
            for (int64_t t = 0; t < n; ++t) o[t] = y2[t] / gain;
        }
        free(y1);
        free(y2);
    }
}

/* ---- iterative radix-2 complex FFT, float64, in place; sign=-1 forward, +1 inverse
 * (unscaled).  Twiddles from a precomputed table of exp(sign*2*pi*i*k/N). ---- */
static void fft_pow2(cplx *a, int64_t N, int sign, const cplx *tw /* N/2 entries, forward */)
{
    /* bit reversal */
    for (int64_t i = 1, j = 0; i < N; ++i) {
        int64_t bit = N >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { cplx t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (int64_t len = 2; len <= N; len <<= 1) {
        int64_t half = len >> 1, step = N / len;
        for (int64_t i = 0; i < N; i += len) {
            for (int64_t k = 0; k < half; ++k) {
                cplx w = tw[k * step];
                if (sign > 0) w.im = -w.im;
                cplx u = a[i + k], v = c_mul(a[i + k + half], w);
                a[i + k].re = u.re + v.re; a[i + k].im = u.im + v.im;
                a[i + k + half].re = u.re - v.re; a[i + k + half].im = u.im - v.im;
            }
        }
    }
}

int64_t f2o_next_pow2(int64_t n)
{
    /* EnvelopeExtraction.py:29  int(2 ** ceil(log2(len(signal)))) */
    int64_t p = 1;
    while (p < n) p <<= 1;
    return p;
}

/* ---- EnvelopeExtraction.py:20-36 paddedHilbert + scipy.signal.hilbert ----
 * Zero-pad `signal` (n) to N2 = 2^ceil(log2 n), FFT, apply the one-sided mask
 * h[0]=h[N/2]=1, h[1:N/2]=2, h[N/2+1:]=0 (N even; N==1 -> h=[1]), IFFT, keep [0,n).
 * Writes the analytic signal as separate real/imag arrays (either may be NULL).
 * `work` must hold N2 cplx, `tw` N2/2 forward twiddles (see f2o_make_twiddles). */
static void padded_hilbert_work(const double *signal, int64_t n, int64_t N2, cplx *work, const cplx *tw,
                                double *out_re, double *out_im)
{
    for (int64_t t = 0; t < n; ++t) { work[t].re = signal[t]; work[t].im = 0.0; }
    for (int64_t t = n; t < N2; ++t) { work[t].re = 0.0; work[t].im = 0.0; }
    if (N2 > 1) {
        fft_pow2(work, N2, -1, tw);
        for (int64_t k = 1; k < N2 / 2; ++k) { work[k].re *= 2.0; work[k].im *= 2.0; }
        for (int64_t k = N2 / 2 + 1; k < N2; ++k) { work[k].re = 0.0; work[k].im = 0.0; }
        fft_pow2(work, N2, +1, tw);
        double s = 1.0 / (double)N2;
        for (int64_t t = 0; t < n; ++t) { work[t].re *= s; work[t].im *= s; }
    }
    if (out_re) for (int64_t t = 0; t < n; ++t) out_re[t] = work[t].re;
    if (out_im) for (int64_t t = 0; t < n; ++t) out_im[t] = work[t].im;
}

static cplx *make_twiddles(int64_t N2)
{
    int64_t h = N2 / 2 > 0 ? N2 / 2 : 1;
    cplx *tw = (cplx *)malloc(sizeof(cplx) * (size_t)h);
    for (int64_t k = 0; k < h; ++k) {
        double ang = -2.0 * M_PI * (double)k / (double)N2;
        tw[k].re = cos(ang);
        tw[k].im = sin(ang);
    }
    return tw;
}

void f2o_padded_hilbert(const double *signal, int64_t n, double *out_re, double *out_im)
{
    if (n <= 0) return;
    int64_t N2 = f2o_next_pow2(n);
    cplx *work = (cplx *)malloc(sizeof(cplx) * (size_t)N2);
    cplx *tw = make_twiddles(N2);
    padded_hilbert_work(signal, n, N2, work, tw, out_re, out_im);
    free(work);
    free(tw);
}

/* ---- scipy.signal.butter(1, Wn, 'low') restated: analog prototype 1/(s+1),
 * pre-warp with fs=2 (warped = 4*tan(pi*Wn/2)), lp2lp, bilinear.  Returns
 * b = [b0, b0], a = [1, a1].  Called as butter(1, freq/(16000/2))
 * (EnvelopeExtraction.py:47: the 8 kHz Nyquist is hard-coded). ---- */
void f2o_butter1_lowpass(double Wn, double *b, double *a)
{
    double fs = 2.0;
    double warped = 2.0 * fs * tan(M_PI * Wn / fs);
    double fs2 = 2.0 * fs;
    /* analog: zero-free, pole p = -warped, gain k = warped */
    double p = -warped, k = warped;
    double pz = (fs2 + p) / (fs2 - p);
    double kz = k * (1.0 / (fs2 - p));
    /* one zero at z = -1 added by the bilinear transform */
    b[0] = kz; b[1] = kz;
    a[0] = 1.0; a[1] = -pz;
}

/* ---- EnvelopeExtraction.py:39-48 lowPassFilter ---- */
void f2o_low_pass_filter(const double *signal, int64_t n, double freq, double *out)
{
    double b[2], a[2];
    f2o_butter1_lowpass(freq / (16000.0 / 2.0), b, a);
    f2o_lfilter2(b, a, signal, out, n);
}

/* ---- EnvelopeExtraction.py:51-67 ExtractEnvelopeFromMatrix ----
 * matrix (C,n) float64 -> envelopes (C,n) float64: abs(paddedHilbert(row)), then
 * lowPassFilter(., cutoff) iff lpf. */
void f2o_extract_envelope(const double *matrix, int C, int64_t n, int lpf, double cutoff, double *out)
{
    if (n <= 0) return;
    int64_t N2 = f2o_next_pow2(n);
    cplx *tw = make_twiddles(N2);
#pragma omp parallel
    {
        cplx *work = (cplx *)malloc(sizeof(cplx) * (size_t)N2);
        double *amp = (double *)malloc(sizeof(double) * (size_t)n);
#pragma omp for schedule(dynamic, 1)
        for (int c = 0; c < C; ++c) {
            const double *row = matrix + (size_t)c * (size_t)n;
            double *o = out + (size_t)c * (size_t)n;
            padded_hilbert_work(row, n, N2, work, tw, NULL, NULL);
            for (int64_t t = 0; t < n; ++t) amp[t] = hypot(work[t].re, work[t].im); /* numpy.abs :58 */
            if (lpf) f2o_low_pass_filter(amp, n, cutoff, o);
            else memcpy(o, amp, sizeof(double) * (size_t)n);
        }
        free(work);
        free(amp);
    }
    free(tw);
}

/* ---- InputGenerator.py:73-80 window gather for one file ----
 * env (C,n) float64; centers[m]; out (m, 2R+1, C) float32 (cast at :83).
 * entry[j, ch] = env[ch, center + STEP*(j-RADIUS)].  Python list indexing: a
 * negative index wraps once (idx += n); anything still out of range is an
 * IndexError -> returns the 1-based position of the failing window, 0 on success. */
int64_t f2o_gather_windows(const double *env, int C, int64_t n, const int64_t *centers, int64_t m,
                           int radius, int64_t step, float *out)
{
    int dots = 2 * radius + 1;
    for (int64_t i = 0; i < m; ++i) {
        for (int j = 0; j < dots; ++j) {
            int64_t idx = centers[i] + step * (int64_t)(j - radius);
            if (idx < 0) idx += n;
            if (idx < 0 || idx >= n) return i + 1;
            float *o = out + ((size_t)i * dots + j) * (size_t)C;
            for (int c = 0; c < C; ++c) o[c] = (float)env[(size_t)c * (size_t)n + idx];
        }
    }
    return 0;
}

/* ---- Evaluating.py:70-78 dense framing ----
 * nb = n - dots*STEP frames; out[i,k,ch] = env[ch, START + i + (k-RADIUS)*STEP]
 * with START = STEP*RADIUS, i.e. env[ch, i + k*STEP].  out (nb, dots, C) float64. */
void f2o_dense_frames(const double *env, int C, int64_t n, int radius, int64_t step, int64_t i0, int64_t i1,
                      double *out)
{
    int dots = 2 * radius + 1;
    int64_t start = step * radius;
#pragma omp parallel for schedule(static)
    for (int64_t i = i0; i < i1; ++i)
        for (int k = 0; k < dots; ++k) {
            int64_t idx = start + i + (int64_t)(k - radius) * step;
            double *o = out + ((size_t)(i - i0) * dots + k) * (size_t)C;
            for (int c = 0; c < C; ++c) o[c] = env[(size_t)c * (size_t)n + idx];
        }
}

/* ---- Training.py:13-28 normalizeInput on one (dots*C) frame, in place ----
 * returns 0 ok, 1 "values must all be positive" (min <= 0), 2 NaN-ordered (min > max). */
int f2o_normalize_input(double *frame, int64_t len)
{
    double mn = frame[0], mx = frame[0];
    for (int64_t i = 1; i < len; ++i) { if (frame[i] < mn) mn = frame[i]; if (frame[i] > mx) mx = frame[i]; }
    if (mn > mx) return 2;
    if (mn <= 0) return 1;
    if (mn == mx) { for (int64_t i = 0; i < len; ++i) frame[i] = 0.0; return 0; }
    double lmn = log(mn), lmx = log(mx);
    for (int64_t i = 0; i < len; ++i) { double v = log(frame[i]); v -= lmn; v /= (lmx - lmn); frame[i] = v; }
    return 0;
}

/* ---- whole hot path for one utterance, the cpu_baseline unit of work:
 * erb_filterbank -> ExtractEnvelopeFromMatrix(lpf,cutoff) -> window gather.
 * gfb/env are caller scratch of C*n doubles. ---- */
int64_t f2o_utterance(const double *wave, int64_t n, const double *coefs, int C, int lpf, double cutoff,
                      const int64_t *centers, int64_t m, int radius, int64_t step, double *gfb, double *env,
                      float *windows)
{
    f2o_erb_filterbank(wave, n, coefs, C, gfb);
    f2o_extract_envelope(gfb, C, n, lpf, cutoff, env);
    if (windows && m > 0) return f2o_gather_windows(env, C, n, centers, m, radius, step, windows);
    return 0;
}

int f2o_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void f2o_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
