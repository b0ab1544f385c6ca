/*
 * oracle/ssimu2_oracle.c — CPU restatement of SSIMULACRA2 v2.1 for the call at
 * /root/reference/src/tq.zig:37.  TEST INFRASTRUCTURE ONLY; "parity unpinned" versus
 * fssimu2 0.1.1 (see ssimu2_oracle.h for the full status note).
 *
 * Every function cites what it restates.  "v2.1 §n" refers to the stage numbering of
 * SURVEY.md Appendix A (the published algorithm: libjxl tools/ssimulacra2.cc and
 * lib/jxl/gauss_blur.cc); "tq.zig:37" is the reference call site whose contract
 * (two RGB8 buffers of w*h*3 bytes, channels = 3, one f64 result) this file honours.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: every fused multiply-add below is
 * an explicit fmaf so the rounding sequence is fixed by the source, not the compiler).
 */
#include "ssimu2_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------ */
/* v2.1 §7 — the 108 weights, order: for c in XYB, for scale 0..5, for n in {1-norm,4-norm}:
 * ssim, artifact, detail_lost.                                                          */
static const double kWeights[108] = {
    0.0, 0.0007376606707406586, 0.0, 0.0, 0.0007793481682867309, 0.0, 0.0,
    0.0004371155730107379, 0.0, 1.1041726426657346, 0.00066284834129271,
    0.00015231632783718752, 0.0, 0.0016406437456599754, 0.0, 1.8422455520539298,
    11.441172603757666, 0.0, 0.0007989109436015163, 0.000176816438078653, 0.0,
    1.8787594979546387, 10.94906990605142, 0.0, 0.0007289346991508072,
    0.9677937080626833, 0.0, 0.00014003424285435884, 0.9981766977854967,
    0.00031949755934435053, 0.0004550992113792063, 0.0, 0.0, 0.0013648766163243398,
    0.0, 0.0, 0.0, 0.0, 0.0, 7.466890328078848, 0.0, 17.445833984131262,
    0.0006235601634041466, 0.0, 0.0, 6.683678146179332, 0.00037724407979611296,
    1.027889937768264, 225.20515300849274, 0.0, 0.0, 19.213238186143016,
    0.0011401524586618361, 0.001237755635509985, 176.39317598450694, 0.0, 0.0,
    24.43300999870476, 0.28520802612117757, 0.0004485436923833408, 0.0, 0.0, 0.0,
    34.77906344483772, 44.835625328877896, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0,
    0.0008680556573291698, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0005313191874358747, 0.0,
    0.00016533814161379112, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0004179171803251336,
    0.0017290828234722833, 0.0, 0.0020827005846636437, 0.0, 0.0, 8.826982764996862,
    23.19243343998926, 0.0, 95.1080498811086, 0.9863978034400682, 0.9834382792465353,
    0.0012286405048278493, 171.2667255897307, 0.9807858872435379, 0.0, 0.0, 0.0,
    0.0005130064588990679, 0.0, 0.00010854057858411537};

const double *oracle_weights(void) { return kWeights; }

/* ------------------------------------------------------------------------------------ */
/* v2.1 §1 — sRGB transfer function on v = u8/255.                                       */
int oracle_get_variant(void);
void oracle_srgb_lut(float lut[256])
{
    if (oracle_get_variant() & ORACLE_VARIANT_F32_TRANSFER) {   /* the transfer function evaluated in binary32 */
        for (int i = 0; i < 256; ++i) {
            const float v = (float)i / 255.0f;
            lut[i] = (v <= 0.04045f) ? v / 12.92f : powf((v + 0.055f) / 1.055f, 2.4f);
        }
        return;
    }
    for (int i = 0; i < 256; ++i) {
        double v = (double)i / 255.0;
        double l = (v <= 0.04045) ? v / 12.92 : pow((v + 0.055) / 1.055, 2.4);
        lut[i] = (float)l;
    }
}

/* tq.zig:37 passes interleaved RGB8; the scorer works on planar linear f32 (v2.1 §1). */
void oracle_rgb8_to_linear(const uint8_t *rgb, int w, int h, int stride, float *planes)
{
    float lut[256];
    oracle_srgb_lut(lut);
    const size_t n = (size_t)w * h;
    for (int y = 0; y < h; ++y) {
        const uint8_t *row = rgb + (size_t)y * stride;
        float *r = planes + (size_t)y * w, *g = r + n, *b = g + n;
        for (int x = 0; x < w; ++x) {
            r[x] = lut[row[3 * x + 0]];
            g[x] = lut[row[3 * x + 1]];
            b[x] = lut[row[3 * x + 2]];
        }
    }
}

/* v2.1 §2 — Downsample(in, 2, 2): sum over iy then ix with clamped coordinates, then
 * multiply by 1/4.  The addition order ((p00 + p01) + p10) + p11 is kept.               */
void oracle_downsample2x(const float *in, int w, int h, float *out)
{
    const int ow = (w + 1) / 2, oh = (h + 1) / 2;
    for (int oy = 0; oy < oh; ++oy) {
        const int y0 = 2 * oy, y1 = (2 * oy + 1 < h) ? 2 * oy + 1 : h - 1;
        const float *r0 = in + (size_t)y0 * w, *r1 = in + (size_t)y1 * w;
        float *o = out + (size_t)oy * ow;
        for (int ox = 0; ox < ow; ++ox) {
            const int x0 = 2 * ox, x1 = (2 * ox + 1 < w) ? 2 * ox + 1 : w - 1;
            float s = 0.0f;
            s += r0[x0];
            s += r0[x1];
            s += r1[x0];
            s += r1[x1];
            o[ox] = s * 0.25f;
        }
    }
}

/* v2.1 §3 — opsin absorbance (FMA chain, bias innermost), cube root minus cbrt(bias),
 * X/Y/B mix, then MakePositiveXYB.                                                       */
static const float kM00 = 0.30f, kM02 = 0.078f;
static const float kM10 = 0.23f, kM12 = 0.078f;
static const float kM20 = 0.24342268924547819f, kM21 = 0.20476744424496821f;
static const float kOpsinBias = 0.0037930732552754493f;

/* Cube root.  The published implementation does not call libm here either: it uses its own
 * fast approximation (polynomial seed + Newton steps, ~1e-6 relative).  This restatement fixes
 * one explicit sequence of IEEE-754 binary32 operations (integer seed for x^(-1/3), two
 * multiply-only Newton steps, c = x*y*y, one fused correction; max error 0.77 ulp measured over
 * [0.0037, 1.2]) so that any conforming host or device reproduces it bit for bit.
 * oracle_set_libm_cbrt(1) switches to libm's cbrtf to measure how much that choice matters. */
static int g_use_libm_cbrt = 0;
void oracle_set_libm_cbrt(int on) { g_use_libm_cbrt = on; }

/* Plausible-variant switches (ORACLE_VARIANT_*): readings of the published code this restatement cannot
 * settle without an authoritative copy.  Used by scripts/variant_envelope.py to bound how far "the"
 * SSIMULACRA2 score moves between them — the honest error bar on "matches fssimu2".                    */
static int g_variant = 0;
void oracle_set_variant(int flags) { g_variant = flags; }
int oracle_get_variant(void) { return g_variant; }

float oracle_cbrtf(float x)
{
    if (g_use_libm_cbrt) return cbrtf(x);
    if (!(x > 0.0f)) return 0.0f;
    uint32_t ix;
    memcpy(&ix, &x, 4);
    const uint32_t iy = 0x54a2fa8cu - ix / 3u;
    float y;
    memcpy(&y, &iy, 4);
    const float third = 0.333333343f;
    for (int k = 0; k < 2; ++k) {
        const float y3 = y * y * y;
        const float t = fmaf(-x, y3, 4.0f);
        y = y * t * third;
    }
    const float y2 = y * y;
    float c = x * y2;
    const float r = fmaf(-(c * c), c, x);
    c = fmaf(r, y2 * third, c);
    return c;
}

void oracle_linear_to_xyb(const float *lin, int n, float *xyb)
{
    const float m01 = 1.0f - kM02 - kM00; /* 0.622 */
    const float m11 = 1.0f - kM12 - kM10; /* 0.692 */
    const float m22 = 1.0f - kM20 - kM21; /* 0.5518... */
    const float neg_cb = -oracle_cbrtf(kOpsinBias);
    const float *r = lin, *g = lin + n, *b = lin + 2 * (size_t)n;
    float *X = xyb, *Y = xyb + n, *B = xyb + 2 * (size_t)n;
    for (int i = 0; i < n; ++i) {
        float m0 = fmaf(kM00, r[i], fmaf(m01, g[i], fmaf(kM02, b[i], kOpsinBias)));
        float m1 = fmaf(kM10, r[i], fmaf(m11, g[i], fmaf(kM12, b[i], kOpsinBias)));
        float m2 = fmaf(kM20, r[i], fmaf(kM21, g[i], fmaf(m22, b[i], kOpsinBias)));
        m0 = m0 < 0.0f ? 0.0f : m0;
        m1 = m1 < 0.0f ? 0.0f : m1;
        m2 = m2 < 0.0f ? 0.0f : m2;
        const float L = oracle_cbrtf(m0) + neg_cb;
        const float M = oracle_cbrtf(m1) + neg_cb;
        const float S = oracle_cbrtf(m2) + neg_cb;
        const float x = 0.5f * (L - M);
        const float yv = 0.5f * (L + M);
        /* MakePositiveXYB: B first (uses the un-offset Y), then X, then Y. */
        B[i] = (S - yv) + 0.55f;
        X[i] = x * 14.0f + 0.42f;
        Y[i] = yv + 0.01f;
    }
}

/* ------------------------------------------------------------------------------------ */
/* v2.1 §4 — CreateRecursiveGaussian(sigma): Charalampidis 2016, truncated cosines
 * k in {1,3,5}.  Equation numbers are the paper's.                                       */
static void inv3x3(double a[9])
{
    double m[9];
    memcpy(m, a, sizeof m);
    const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
                       m[2] * (m[3] * m[7] - m[4] * m[6]);
    const double id = 1.0 / det;
    a[0] = (m[4] * m[8] - m[5] * m[7]) * id;
    a[1] = (m[2] * m[7] - m[1] * m[8]) * id;
    a[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    a[3] = (m[5] * m[6] - m[3] * m[8]) * id;
    a[4] = (m[0] * m[8] - m[2] * m[6]) * id;
    a[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    a[6] = (m[3] * m[7] - m[4] * m[6]) * id;
    a[7] = (m[1] * m[6] - m[0] * m[7]) * id;
    a[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

void oracle_rg_coeffs(double sigma, double n2[3], double d1[3], int *radius)
{
    const double kPi = 3.141592653589793238;
    const double N = (double)roundf((float)(3.2795 * sigma + 0.2546)); /* (57) */
    const double pi_div_2r = kPi / (2.0 * N);
    const double omega[3] = {pi_div_2r, 3.0 * pi_div_2r, 5.0 * pi_div_2r};
    /* (37) */
    const double p1 = +1.0 / tan(0.5 * omega[0]);
    const double p3 = -1.0 / tan(0.5 * omega[1]);
    const double p5 = +1.0 / tan(0.5 * omega[2]);
    /* (44) */
    const double r1 = +p1 * p1 / sin(omega[0]);
    const double r3 = -p3 * p3 / sin(omega[1]);
    const double r5 = +p5 * p5 / sin(omega[2]);
    /* (50) */
    const double neg_half_sigma2 = -0.5 * sigma * sigma;
    double rho[3];
    for (int i = 0; i < 3; ++i) rho[i] = exp(neg_half_sigma2 * omega[i] * omega[i]) / N;
    /* (52) */
    const double D13 = p1 * r3 - r1 * p3;
    const double D35 = p3 * r5 - r3 * p5;
    const double D51 = p5 * r1 - r5 * p1;
    const double zeta15 = D35 / D13;
    const double zeta35 = D51 / D13;
    double A[9] = {p1, p3, p5, r1, r3, r5, zeta15, zeta35, 1.0}; /* (56) */
    inv3x3(A);
    const double gamma[3] = {1.0, N * N - sigma * sigma, /* (55) */
                             zeta15 * rho[0] + zeta35 * rho[1] + rho[2]};
    double beta[3];
    for (int i = 0; i < 3; ++i) /* (53) */
        beta[i] = A[3 * i] * gamma[0] + A[3 * i + 1] * gamma[1] + A[3 * i + 2] * gamma[2];
    for (int i = 0; i < 3; ++i) {
        n2[i] = -beta[i] * cos(omega[i] * (N + 1.0)); /* (33) */
        d1[i] = -2.0 * cos(omega[i]);                 /* (35) */
    }
    *radius = (int)N;
}

/* The recursion fed with a unit impulse, in double: its response is exactly finite,
 * 2N-1 taps (the two injections at n-N-1 and n+N-1 start and cancel each oscillator).    */
int oracle_fir_taps(double sigma, double *taps, int max_taps)
{
    double n2[3], d1[3];
    int N;
    oracle_rg_coeffs(sigma, n2, d1, &N);
    const int len = 4 * N + 1, c = 2 * N; /* impulse at c */
    const int ntaps = 2 * N - 1;
    if (max_taps < ntaps) return -1;
    double prev[3] = {0, 0, 0}, prev2[3] = {0, 0, 0};
    for (int n = -N + 1; n < len; ++n) {
        const int l = n - N - 1, r = n + N - 1;
        const double lv = (l == c) ? 1.0 : 0.0, rv = (r == c) ? 1.0 : 0.0;
        double o = 0.0;
        for (int k = 0; k < 3; ++k) {
            const double ok = n2[k] * (lv + rv) - d1[k] * prev[k] - prev2[k];
            prev2[k] = prev[k];
            prev[k] = ok;
            o += ok;
        }
        const int t = n - (c - (N - 1));
        if (t >= 0 && t < ntaps) taps[t] = o;
    }
    return ntaps;
}

typedef struct {
    float n2[3], d1[3];
    int N;
    float taps[16];
    double taps64[16];
    int ntaps;
} blur_consts;

static const blur_consts *get_blur_consts(void)
{
    static blur_consts bc;
    static int init = 0;
    if (!init) {
        double n2[3], d1[3];
        oracle_rg_coeffs(1.5, n2, d1, &bc.N);
        for (int i = 0; i < 3; ++i) {
            bc.n2[i] = (float)n2[i];
            bc.d1[i] = (float)d1[i];
        }
        bc.ntaps = oracle_fir_taps(1.5, bc.taps64, 16);
        for (int i = 0; i < bc.ntaps; ++i) bc.taps[i] = (float)bc.taps64[i];
        init = 1;
    }
    return &bc;
}

/* v2.1 §4 — FastGaussian1D, scalar form: out_k = n2_k*(l+r) - d1_k*prev_k - prev2_k,
 * evaluated as sum*mul_in, then MulAdd(-1, prev2, .), then MulAdd(-d1, prev, .).        */
static void iir_row(const blur_consts *bc, const float *in, int len, float *out)
{
    const int N = bc->N;
    float p[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
    for (int n = -N + 1; n < len; ++n) {
        const int l = n - N - 1, r = n + N - 1;
        const float lv = l >= 0 ? in[l] : 0.0f;
        const float rv = r < len ? in[r] : 0.0f;
        const float sum = lv + rv;
        float o[3];
        for (int k = 0; k < 3; ++k) {
            float ok = sum * bc->n2[k];
            ok = fmaf(-1.0f, p2[k], ok);
            ok = fmaf(-bc->d1[k], p[k], ok);
            p2[k] = p[k];
            p[k] = ok;
            o[k] = ok;
        }
        if (n >= 0) out[n] = o[0] + o[1] + o[2];
    }
}

/* Vertical pass: the same recursion per column, marched row by row so x vectorises.     */
static void iir_cols(const blur_consts *bc, const float *in, int w, int h, float *out)
{
    const int N = bc->N;
    float *st = (float *)calloc((size_t)w * 6, sizeof(float));
    float *p[3] = {st, st + w, st + 2 * (size_t)w};
    float *p2[3] = {st + 3 * (size_t)w, st + 4 * (size_t)w, st + 5 * (size_t)w};
    for (int n = -N + 1; n < h; ++n) {
        const int l = n - N - 1, r = n + N - 1;
        const float *lrow = l >= 0 ? in + (size_t)l * w : NULL;
        const float *rrow = r < h ? in + (size_t)r * w : NULL;
        float *orow = n >= 0 ? out + (size_t)n * w : NULL;
        const int vorder = g_variant & ORACLE_VARIANT_VERTICAL_ORDER;
        for (int x = 0; x < w; ++x) {
            const float sum = (lrow ? lrow[x] : 0.0f) + (rrow ? rrow[x] : 0.0f);
            float acc = 0.0f;
            for (int k = 0; k < 3; ++k) {
                float ok;
                if (vorder) {
                    /* lib/jxl/gauss_blur.cc VerticalBlock as recalled: MulAdd(n2, sum, NegMulSub(d1, y[n-1], y[n-2])),
                     * i.e. the product n2*sum is NOT rounded on its own and -d1*y[n-1] - y[n-2] is formed first */
                    ok = fmaf(bc->n2[k], sum, fmaf(-bc->d1[k], p[k][x], -p2[k][x]));
                } else {
                    ok = sum * bc->n2[k];
                    ok = fmaf(-1.0f, p2[k][x], ok);
                    ok = fmaf(-bc->d1[k], p[k][x], ok);
                }
                p2[k][x] = p[k][x];
                p[k][x] = ok;
                acc = (k == 0) ? ok : acc + ok;
            }
            if (orow) orow[x] = acc;
        }
    }
    free(st);
}

/* FIR forms: out[i] = sum_t taps[t] * in[i + t - (N-1)], zero outside, accumulated left
 * to right with fmaf starting from taps[0]*in (f32) or in double (FIR64).               */
static void fir_pass(const blur_consts *bc, const float *in, int w, int h, int vertical,
                     int f64acc, float *out)
{
    const int nt = bc->ntaps, c = nt / 2;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float acc = 0.0f;
            double acc64 = 0.0;
            for (int t = 0; t < nt; ++t) {
                const int xx = vertical ? x : x + t - c;
                const int yy = vertical ? y + t - c : y;
                const float v = (xx >= 0 && xx < w && yy >= 0 && yy < h) ? in[(size_t)yy * w + xx] : 0.0f;
                if (f64acc)
                    acc64 += bc->taps64[t] * (double)v;
                else
                    acc = (t == 0) ? bc->taps[0] * v : fmaf(bc->taps[t], v, acc);
            }
            out[(size_t)y * w + x] = f64acc ? (float)acc64 : acc;
        }
}

void oracle_blur(const float *in, int w, int h, int mode, float *tmp, float *out)
{
    const blur_consts *bc = get_blur_consts();
    if (mode == ORACLE_BLUR_IIR) {
        for (int y = 0; y < h; ++y) iir_row(bc, in + (size_t)y * w, w, tmp + (size_t)y * w);
        iir_cols(bc, tmp, w, h, out);
    } else {
        fir_pass(bc, in, w, h, 0, mode == ORACLE_BLUR_FIR64, tmp);
        fir_pass(bc, tmp, w, h, 1, mode == ORACLE_BLUR_FIR64, out);
    }
}

/* ------------------------------------------------------------------------------------ */
/* v2.1 §5 — SSIMMap for one channel: float products, double 1-x, double sums.            */
static void ssim_map(const float *m1, const float *m2, const float *s11, const float *s22,
                     const float *s12, size_t n, double sums[2])
{
    const float kC2 = 0.0009f;
    double a = 0.0, b = 0.0;
    for (size_t i = 0; i < n; ++i) {
        const float mu1 = m1[i], mu2 = m2[i];
        const float mu11 = mu1 * mu1, mu22 = mu2 * mu2, mu12 = mu1 * mu2;
        const float dm = (mu1 - mu2) * (mu1 - mu2);
        const float num_m = (float)(1.0 - (double)dm);
        const float num_s = 2.0f * (s12[i] - mu12) + kC2;
        const float denom_s = (s11[i] - mu11) + (s22[i] - mu22) + kC2;
        double d = 1.0 - (double)(num_m * num_s / denom_s);
        d = d > 0.0 ? d : 0.0;
        a += d;
        d *= d;
        d *= d;
        b += d;
    }
    sums[0] = a;
    sums[1] = b;
}

/* v2.1 §6 — EdgeDiffMap for one channel: float differences, double ratio.                */
static void edge_diff_map(const float *i1, const float *m1, const float *i2, const float *m2,
                          size_t n, double sums[4])
{
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (size_t i = 0; i < n; ++i) {
        const double d1 = (1.0 + (double)fabsf(i2[i] - m2[i])) / (1.0 + (double)fabsf(i1[i] - m1[i])) - 1.0;
        double art = d1 > 0.0 ? d1 : 0.0;
        double det = d1 < 0.0 ? -d1 : 0.0;
        s0 += art;
        art *= art;
        art *= art;
        s1 += art;
        s2 += det;
        det *= det;
        det *= det;
        s3 += det;
    }
    sums[0] = s0;
    sums[1] = s1;
    sums[2] = s2;
    sums[3] = s3;
}

/* ORACLE_VARIANT_F32_MAPS: both maps of one channel with binary32 values and per-row binary32 accumulators. */
static void maps_f32(const float *i1, const float *m1, const float *i2, const float *m2, const float *s11,
                     const float *s22, const float *s12, int w, int h, double ss[2], double ed[4])
{
    const float kC2 = 0.0009f;
    double tot[6] = {0, 0, 0, 0, 0, 0};
    for (int y = 0; y < h; ++y) {
        float acc[6] = {0, 0, 0, 0, 0, 0};
        for (int x = 0; x < w; ++x) {
            const size_t i = (size_t)y * w + x;
            const float mu1 = m1[i], mu2 = m2[i];
            const float mu11 = mu1 * mu1, mu22 = mu2 * mu2, mu12 = mu1 * mu2;
            const float dm = (mu1 - mu2) * (mu1 - mu2);
            const float num_m = 1.0f - dm;
            const float num_s = 2.0f * (s12[i] - mu12) + kC2;
            const float denom_s = (s11[i] - mu11) + (s22[i] - mu22) + kC2;
            float d = 1.0f - num_m * num_s / denom_s;
            d = d > 0.0f ? d : 0.0f;
            acc[0] += d;
            d *= d;
            acc[1] += d * d;
            const float d1 = (1.0f + fabsf(i2[i] - m2[i])) / (1.0f + fabsf(i1[i] - m1[i])) - 1.0f;
            float art = d1 > 0.0f ? d1 : 0.0f, det = d1 < 0.0f ? -d1 : 0.0f;
            acc[2] += art;
            art *= art;
            acc[3] += art * art;
            acc[4] += det;
            det *= det;
            acc[5] += det * det;
        }
        for (int k = 0; k < 6; ++k) tot[k] += (double)acc[k];
    }
    ss[0] = tot[0]; ss[1] = tot[1];
    ed[0] = tot[2]; ed[1] = tot[3]; ed[2] = tot[4]; ed[3] = tot[5];
}

/* v2.1 §7 — Msssim::Score().                                                             */
double oracle_final_score(int n_scales, const double avg_ssim[][6], const double avg_edgediff[][12])
{
    double ssim = 0.0;
    int i = 0;
    /* Two readings for images with fewer than six scales (identical at six):
     *   default — a fixed six-scale weight table, absent scales contribute zero;
     *   ORACLE_VARIANT_CONTIGUOUS_WEIGHTS — libjxl Msssim::Score() as recalled: `scale < scales.size()` with a
     *   running i++, so the weights are consumed contiguously over the scales present.                      */
    const int nloop = (g_variant & ORACLE_VARIANT_CONTIGUOUS_WEIGHTS) ? n_scales : 6;
    for (int c = 0; c < 3; ++c)
        for (int scale = 0; scale < nloop; ++scale)
            for (int n = 0; n < 2; ++n) {
                if (scale >= n_scales) {
                    i += 3;
                    continue;
                }
                ssim += kWeights[i++] * fabs(avg_ssim[scale][c * 2 + n]);
                ssim += kWeights[i++] * fabs(avg_edgediff[scale][c * 4 + n]);
                ssim += kWeights[i++] * fabs(avg_edgediff[scale][c * 4 + n + 2]);
            }
    ssim = ssim * 0.9562382616834844;
    ssim = 2.326765642916932 * ssim - 0.020884521182843837 * ssim * ssim +
           6.248496625763138e-05 * ssim * ssim * ssim;
    if (ssim > 0.0)
        ssim = 100.0 - 10.0 * pow(ssim, 0.6276336467831387);
    else
        ssim = 100.0;
    return ssim;
}

/* ------------------------------------------------------------------------------------ */
/* Whole path — ComputeSSIMULACRA2 as called from tq.zig:37 (RGB8 pair, no alpha).        */
int oracle_ssimu2_rgb8(const uint8_t *ref, int ref_stride, const uint8_t *dist, int dist_stride,
                       int w, int h, int blur_mode, double *score, oracle_detail *detail)
{
    if (!ref || !dist || w <= 0 || h <= 0 || !score) return -1;
    if (ref_stride < 3 * w || dist_stride < 3 * w) return -1;
    oracle_detail local;
    oracle_detail *D = detail ? detail : &local;
    memset(D, 0, sizeof *D);

    const size_t n0 = (size_t)w * h;
    float *lin1 = (float *)malloc(n0 * 3 * sizeof(float));
    float *lin2 = (float *)malloc(n0 * 3 * sizeof(float));
    float *xyb1 = (float *)malloc(n0 * 3 * sizeof(float));
    float *xyb2 = (float *)malloc(n0 * 3 * sizeof(float));
    float *buf = (float *)malloc(n0 * 7 * sizeof(float)); /* mul,tmp,mu1,mu2,s11,s22,s12 */
    float *dn = (float *)malloc(((size_t)(w + 1) / 2) * ((h + 1) / 2) * 3 * sizeof(float));
    if (!lin1 || !lin2 || !xyb1 || !xyb2 || !buf || !dn) {
        free(lin1); free(lin2); free(xyb1); free(xyb2); free(buf); free(dn);
        return -2;
    }
    oracle_rgb8_to_linear(ref, w, h, ref_stride, lin1);
    oracle_rgb8_to_linear(dist, w, h, dist_stride, lin2);

    int cw = w, ch = h;
    for (int scale = 0; scale < ORACLE_MAX_SCALES; ++scale) {
        if (cw < 8 || ch < 8) break;
        if (scale) {
            const int nw = (cw + 1) / 2, nh = (ch + 1) / 2;
            const size_t on = (size_t)cw * ch, nn = (size_t)nw * nh;
            for (int c = 0; c < 3; ++c) oracle_downsample2x(lin1 + c * on, cw, ch, dn + c * nn);
            memcpy(lin1, dn, nn * 3 * sizeof(float));
            for (int c = 0; c < 3; ++c) oracle_downsample2x(lin2 + c * on, cw, ch, dn + c * nn);
            memcpy(lin2, dn, nn * 3 * sizeof(float));
            cw = nw;
            ch = nh;
            /* NB: as published, the size test above looks at the PREVIOUS scale's
             * dimensions, so a scale whose own size is below 8 is still evaluated once
             * (e.g. 100x100 -> scales 100,50,25,13,7).                                   */
        }
        const size_t n = (size_t)cw * ch;
        oracle_linear_to_xyb(lin1, (int)n, xyb1);
        oracle_linear_to_xyb(lin2, (int)n, xyb2);
        float *mul = buf, *tmp = buf + n, *mu1 = buf + 2 * n, *mu2 = buf + 3 * n;
        float *s11 = buf + 4 * n, *s22 = buf + 5 * n, *s12 = buf + 6 * n;
        for (int c = 0; c < 3; ++c) {
            const float *a = xyb1 + c * n, *b = xyb2 + c * n;
            for (size_t i = 0; i < n; ++i) mul[i] = a[i] * a[i];
            oracle_blur(mul, cw, ch, blur_mode, tmp, s11);
            for (size_t i = 0; i < n; ++i) mul[i] = b[i] * b[i];
            oracle_blur(mul, cw, ch, blur_mode, tmp, s22);
            for (size_t i = 0; i < n; ++i) mul[i] = a[i] * b[i];
            oracle_blur(mul, cw, ch, blur_mode, tmp, s12);
            oracle_blur(a, cw, ch, blur_mode, tmp, mu1);
            oracle_blur(b, cw, ch, blur_mode, tmp, mu2);
            double ss[2], ed[4];
            if (g_variant & ORACLE_VARIANT_F32_MAPS) {
                maps_f32(a, mu1, b, mu2, s11, s22, s12, cw, ch, ss, ed);
            } else {
                ssim_map(mu1, mu2, s11, s22, s12, n, ss);
                edge_diff_map(a, mu1, b, mu2, n, ed);
            }
            const double opp = 1.0 / (double)n;
            D->sums[scale][c * 6 + 0] = ss[0];
            D->sums[scale][c * 6 + 1] = ss[1];
            D->sums[scale][c * 6 + 2] = ed[0];
            D->sums[scale][c * 6 + 3] = ed[1];
            D->sums[scale][c * 6 + 4] = ed[2];
            D->sums[scale][c * 6 + 5] = ed[3];
            D->avg_ssim[scale][c * 2 + 0] = opp * ss[0];
            D->avg_ssim[scale][c * 2 + 1] = sqrt(sqrt(opp * ss[1]));
            D->avg_edgediff[scale][c * 4 + 0] = opp * ed[0];
            D->avg_edgediff[scale][c * 4 + 1] = sqrt(sqrt(opp * ed[1]));
            D->avg_edgediff[scale][c * 4 + 2] = opp * ed[2];
            D->avg_edgediff[scale][c * 4 + 3] = sqrt(sqrt(opp * ed[3]));
        }
        D->w[scale] = cw;
        D->h[scale] = ch;
        D->n_scales = scale + 1;
    }
    D->score = oracle_final_score(D->n_scales, D->avg_ssim, D->avg_edgediff);
    *score = D->score;
    free(lin1); free(lin2); free(xyb1); free(xyb2); free(buf); free(dn);
    return 0;
}

int oracle_xyb_at_scale(const uint8_t *rgb, int stride, int w, int h, int scale, float *xyb,
                        int *ws, int *hs)
{
    if (!rgb || w <= 0 || h <= 0 || scale < 0 || scale >= ORACLE_MAX_SCALES) return -1;
    const size_t n0 = (size_t)w * h;
    float *lin = (float *)malloc(n0 * 3 * sizeof(float));
    float *dn = (float *)malloc(n0 * 3 * sizeof(float));
    if (!lin || !dn) { free(lin); free(dn); return -2; }
    oracle_rgb8_to_linear(rgb, w, h, stride, lin);
    int cw = w, ch = h;
    for (int s = 0; s < scale; ++s) {
        const int nw = (cw + 1) / 2, nh = (ch + 1) / 2;
        const size_t on = (size_t)cw * ch, nn = (size_t)nw * nh;
        for (int c = 0; c < 3; ++c) oracle_downsample2x(lin + c * on, cw, ch, dn + c * nn);
        memcpy(lin, dn, nn * 3 * sizeof(float));
        cw = nw;
        ch = nh;
    }
    oracle_linear_to_xyb(lin, cw * ch, xyb);
    *ws = cw;
    *hs = ch;
    free(lin);
    free(dn);
    return 0;
}
