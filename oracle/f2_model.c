/*
 * f2_model.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A scalar float32 CPU model of the arithmetic the CUDA kernels perform (streaming
 * ring formulation of the padded-FFT Hilbert envelope, delta-form biquads, truncated
 * warm-ups).  It exists to study precision and warm-up lengths offline, where there
 * is no GPU; it is not the oracle (that is f2_oracle.c, float64, the reference's
 * algorithm) and it is not a product path.
 *
 * Per channel (parameters prepared in float64, rounded once to float32):
 *   cq = B2, cy = 1 + B1 + B2   (pole pair, "delta" form: q = y[t]-y[t-1])
 *   z[k] = A1k/A0 (stage zero), g4 = A0^4/gain (output scale; stages run without a0)
 * Stage update (input u[t], up = u[t-1]):
 *   in = u + z*up (+ e*G[t] on the imaginary path)
 *   delta form:   q  = cq*q + in ; q = q - cy*y ; y = y + q
 *   direct form:  y  = (in - B2*y[t-2]) - B1*y[t-1]          (3 FMAs instead of 3 + 1 add)
 * `form` argument of f2m_run: 0 = delta everywhere, 1 = direct everywhere, 2 = what the kernel
 * does: direct for every group of 32 adjacent channels whose min(1+B1+B2) >= the threshold set
 * with f2m_set_direct_min_cy (the kernel: 0.035 for envelope-only runs, 0.25 when the filterbank
 * output itself is stored), delta for the others.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    float cq[4], cy[4], z[4], g4;
    float nb1[4], nb2[4]; /* direct form I: -B1, -B2 */
} chan_t;

/* Round c to float32 per stage so that the four stage values sum to 4*c as closely as
 * the float32 lattice allows: the first-order response error of the cascade depends
 * only on the sum of the per-stage denominator errors (shared poles). */
static int g_dither = 1;
static double g_direct_min_cy = 0.035;
void f2m_set_direct_min_cy(double v) { g_direct_min_cy = v; }
void f2m_set_dither(int d) { g_dither = d; }
static void dither4(double c, float *out)
{
    float r = (float)c;
    if (!g_dither) { for (int k = 0; k < 4; ++k) out[k] = r; return; }
    float lo = ((double)r <= c) ? r : nextafterf(r, -INFINITY);
    float hi = nextafterf(lo, INFINITY);
    double frac = (c - (double)lo) / ((double)hi - (double)lo);
    int m = (int)floor(4.0 * frac + 0.5);
    for (int k = 0; k < 4; ++k) out[k] = (k < m) ? hi : lo;
}

static void prep(const double *k10, chan_t *p)
{
    dither4(k10[8], p->cq);
    dither4(1.0 + k10[7] + k10[8], p->cy);
    dither4(-k10[7], p->nb1);
    dither4(-k10[8], p->nb2);
    p->g4 = (float)(k10[0] * k10[0] * k10[0] * k10[0] / k10[9]);
    for (int k = 0; k < 4; ++k) p->z[k] = (float)(k10[1 + k] / k10[0]);
}

/* one cascade step on one path.  y[4], q[4] states; u = input sample, up = previous
 * input; inj[k] = injection for stage k (0 on the real path). */
static inline float cascade_step(const chan_t *p, float *y, float *q, float u, float up, const float *inj, int form)
{
    if (form == 1) {
        /* direct form: y[k] = y[t-1], q[k] = y[t-2]; acc = in - B2*y[t-2]; y = acc - B1*y[t-1]; the next
         * section's input is formed from the same (y[t-1], acc): in' = acc + (z' - B1)*y[t-1] */
        float in = fmaf(p->z[0], up, u);
        if (inj) in = in + inj[0];
        for (int k = 0; k < 4; ++k) {
            float yp = y[k];
            float acc = fmaf(p->nb2[k], q[k], in);
            float yn = fmaf(p->nb1[k], yp, acc);
            if (k < 3) {
                in = fmaf(p->z[k + 1] + p->nb1[k], yp, acc);
                if (inj) in = in + inj[k + 1];
            }
            q[k] = yp;
            y[k] = yn;
        }
        return y[3];
    }
    for (int k = 0; k < 4; ++k) {
        float in = fmaf(p->z[k], up, u);
        if (inj) in = in + inj[k];
        float yold = y[k];
        float qn = fmaf(p->cq[k], q[k], in);
        qn = fmaf(-p->cy[k], yold, qn);
        float yn = yold + qn;
        q[k] = qn;
        y[k] = yn;
        up = yold; /* next stage: previous input = this stage's previous output */
        u = yn;
    }
    return y[3];
}

/* xf/xi: ring arrays of length N2 (x zero-padded; Im of its circular analytic signal).
 * G: ring array, G[t] = h[(t-n) mod N2] if (t-n) odd else h[(t-n-1) mod N2].
 * Outputs (C,n) float32 row-major; either may be NULL. */
void f2m_run(const float *xf, const float *xi, const float *G, int64_t n, int64_t N2, const double *coefs, int C,
             int lpf, double lp_b0, double lp_a1, int64_t W, int64_t We, int form_arg, float *out_gfb, float *out_env)
{
    const float k_lp = (float)(-lp_a1), b0_lp = (float)lp_b0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < C; ++c) {
        chan_t p;
        prep(coefs + (size_t)c * 10, &p);
        int form = form_arg;
        if (form_arg == 2) {
            double cy = 1e300;
            for (int cc = c / 32 * 32; cc < C && cc < c / 32 * 32 + 32; ++cc) {
                const double *kk = coefs + (size_t)cc * 10;
                if (1.0 + kk[7] + kk[8] < cy) cy = 1.0 + kk[7] + kk[8];
            }
            form = cy >= g_direct_min_cy;
        }
        /* E phase: real path over the last We samples from zero state -> edge residuals */
        float y[4] = { 0, 0, 0, 0 }, q[4] = { 0, 0, 0, 0 };
        int64_t t0 = n - We > 0 ? n - We : 0;
        float up = t0 > 0 ? xf[t0 - 1] : 0.0f;
        for (int64_t t = t0; t < n; ++t) { cascade_step(&p, y, q, xf[t], up, NULL, form); up = xf[t]; }
        /* e0[k] at ring index n, e1[k] at n+1:
         *   e0 = b1*y[n-1] + b2*y[n-2] - a1k*yprev_stage[n-1] = (cy-1)*y - cq*q - a1k*u
         *   e1 = b2*y[n-1] = cq*y */
        float e0[4], e1[4];
        for (int k = 0; k < 4; ++k) {
            float uprev = (k == 0) ? (n > 0 ? xf[n - 1] : 0.0f) : y[k - 1];
            if (form == 1) { /* b1*y[n-1] + b2*y[n-2] - z*u[n-1] ; b2*y[n-1] */
                e0[k] = fmaf(-p.nb1[k], y[k], -p.nb2[k] * q[k]) - p.z[k] * uprev;
                e1[k] = -p.nb2[k] * y[k];
                continue;
            }
            e0[k] = fmaf(p.cy[k] - 1.0f, y[k], -p.cq[k] * q[k]) - p.z[k] * uprev;
            e1[k] = p.cq[k] * y[k];
        }
        /* imaginary path: periodic steady state by a W-sample warm-up on the ring */
        float vy[4] = { 0, 0, 0, 0 }, vq[4] = { 0, 0, 0, 0 };
        float inj[4];
        float vup = 0.0f;
        if (N2 > 2) {
            int64_t r = ((-W) % N2 + N2) % N2;
            vup = xi[(r - 1 + N2) % N2];
            for (int64_t t = -W; t < 0; ++t) {
                int64_t rr = ((t % N2) + N2) % N2;
                int odd = (int)(((t - n) % 2 + 2) % 2);
                for (int k = 0; k < 4; ++k) inj[k] = (odd ? e0[k] : e1[k]) * G[rr];
                cascade_step(&p, vy, vq, xi[rr], vup, inj, form);
                vup = xi[rr];
            }
        }
        /* main pass */
        for (int k = 0; k < 4; ++k) { y[k] = 0; q[k] = 0; }
        up = 0.0f;
        float w = 0.0f, wprev = 0.0f;
        for (int64_t t = 0; t < n; ++t) {
            float yr = cascade_step(&p, y, q, xf[t], up, NULL, form);
            up = xf[t];
            float yi = 0.0f;
            if (N2 > 2) {
                int odd = (int)(((t - n) % 2 + 2) % 2);
                for (int k = 0; k < 4; ++k) inj[k] = (odd ? e0[k] : e1[k]) * G[t];
                yi = cascade_step(&p, vy, vq, xi[t], vup, inj, form);
                vup = xi[t];
            }
            if (out_gfb) out_gfb[(size_t)c * n + t] = p.g4 * yr;
            if (out_env) {
                float e = sqrtf(fmaf(yr, yr, yi * yi));
                if (lpf) {
                    /* one-pole state w = e + k*w ; out = b0*(w[t] + w[t-1]) */
                    w = fmaf(k_lp, w, e);
                    out_env[(size_t)c * n + t] = (p.g4 * b0_lp) * (w + wprev);
                    wprev = w;
                } else out_env[(size_t)c * n + t] = p.g4 * e;
            }
        }
    }
}
