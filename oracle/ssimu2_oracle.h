/*
 * oracle/ssimu2_oracle.h — CPU restatement of the scoring path of oavif's target-quality loop.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and there only as the checker / the timed CPU baseline.  The shipped path is the CUDA
 * library behind include/oavif_ssimu2.h and it never links or calls this file.
 *
 * PARITY STATUS: *** parity unpinned *** for the SSIMULACRA2 scorer.
 *   The arithmetic the reference executes at /root/reference/src/tq.zig:37
 *   (`fssimu2.computeSsimu2(allocator, e.rgb, decoded_rgb, e.w, e.h, 3, null)`) lives in the
 *   un-vendored Zig dependency gianni-rosato/fssimu2 tag 0.1.1
 *   (/root/reference/build.zig.zon:7-10, wired at build.zig:30-33,65).  Its source is not in
 *   /root/reference, there is no network and no zig toolchain, and the reference has no
 *   tests / golden vectors.  This file therefore restates the PUBLISHED
 *   algorithm fssimu2 derives from — SSIMULACRA2 v2.1 (libjxl tools/ssimulacra2.cc +
 *   lib/jxl/gauss_blur.cc, Charalampidis 2016 recursive Gaussian) — and is anchored on the
 *   reference's call site (two tight interleaved RGB8 buffers, channels = 3, one f64 out)
 *   and on analytic known answers (tests/test_oracle.py).
 *   The decoded-YUV -> RGB8 step (yuv2rgb_oracle.c) IS pinned: bit-exact against libavif
 *   1.4.1's avifImageYUVToRGB, the function the reference calls at src/io.zig:478.
 */
#ifndef SSIMU2_ORACLE_H
#define SSIMU2_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_MAX_SCALES 6

/* Blur variants (all compute the same sigma = 1.5 filter):
 *   0  IIR  f32 — libjxl FastGaussian recursion, the faithful default
 *   1  FIR  f32 — the exactly-equivalent 9-tap kernel, zero padded (what the GPU runs)
 *   2  FIR  f64 accumulate — rounding-free yardstick                                      */
enum { ORACLE_BLUR_IIR = 0, ORACLE_BLUR_FIR = 1, ORACLE_BLUR_FIR64 = 2 };

typedef struct {
    int n_scales;                       /* scales actually evaluated (w,h >= 8)            */
    int w[ORACLE_MAX_SCALES];
    int h[ORACLE_MAX_SCALES];
    /* raw pooled sums per scale: [c*6 + {0:sum d,1:sum d^4,2:sum art,3:sum art^4,4:sum det,5:sum det^4}] */
    double sums[ORACLE_MAX_SCALES][18];
    double avg_ssim[ORACLE_MAX_SCALES][6];      /* [2c+n]  n=0: 1-norm, n=1: 4-norm        */
    double avg_edgediff[ORACLE_MAX_SCALES][12]; /* [4c+{0,1}] artifact, [4c+{2,3}] detail  */
    double score;
} oracle_detail;

/* ---- stages (each usable on its own from the tests) -------------------------------- */

/* sRGB u8 -> linear f32, 256-entry table; evaluated in double, rounded once. */
void oracle_srgb_lut(float lut[256]);

/* interleaved RGB8 (row stride in bytes) -> planar linear f32 (3 planes of w*h). */
void oracle_rgb8_to_linear(const uint8_t *rgb, int w, int h, int stride, float *planes);

/* 2x2 box mean, out = ceil(w/2) x ceil(h/2), source coords clamped (one plane). */
void oracle_downsample2x(const float *in, int w, int h, float *out);

/* the fixed-sequence binary32 cube root used by the XYB stage (see ssimu2_oracle.c) */
float oracle_cbrtf(float x);
void oracle_set_libm_cbrt(int on);

/* Variant switches for the envelope study (process-wide, not thread-safe; default 0 = the restatement the
 * GPU path is held to bit for bit). */
enum {
    ORACLE_VARIANT_CONTIGUOUS_WEIGHTS = 1, /* final sum: running weight index over the scales present     */
    ORACLE_VARIANT_VERTICAL_ORDER = 2,     /* vertical recursion: fma(n2, sum, fma(-d1, y1, -y2)) per step */
    ORACLE_VARIANT_F32_TRANSFER = 8,       /* sRGB transfer function evaluated in binary32 (powf) instead of binary64 */
    ORACLE_VARIANT_F32_MAPS = 4            /* error maps and their powers in binary32, one binary32 accumulator per
                                              image ROW folded into a binary64 total (what a vectorised f32 scorer
                                              plausibly does) instead of binary64 per pixel                    */
};
void oracle_set_variant(int flags);
int oracle_get_variant(void);

/* planar linear RGB -> planar "positive" XYB (X,Y,B order), n pixels per plane. */
void oracle_linear_to_xyb(const float *lin, int n, float *xyb);

/* Recursive-Gaussian coefficients for sigma (n2[3], d1[3], radius), derived in double. */
void oracle_rg_coeffs(double sigma, double n2[3], double d1[3], int *radius);

/* The equivalent FIR taps (2*radius-1 of them, centre at [radius-1]); returns tap count. */
int oracle_fir_taps(double sigma, double *taps, int max_taps);

/* Separable blur of one plane, horizontal then vertical, zero padded. tmp: w*h floats. */
void oracle_blur(const float *in, int w, int h, int mode, float *tmp, float *out);

/* 108-weight sum + nonlinear map.  Missing scales contribute zero (or, under
 * ORACLE_VARIANT_CONTIGUOUS_WEIGHTS, the weights run contiguously over the scales present). */
double oracle_final_score(int n_scales, const double avg_ssim[][6], const double avg_edgediff[][12]);

/* the 108 weights, in the order the final loop consumes them */
const double *oracle_weights(void);

/* ---- whole path ---------------------------------------------------------------------- */

/* Mirrors fssimu2.computeSsimu2(ref, dist, w, h, channels=3): tight or strided RGB8 pairs.
 * Returns 0 on success, <0 on bad arguments / allocation failure. */
int oracle_ssimu2_rgb8(const uint8_t *ref, int ref_stride, const uint8_t *dist, int dist_stride,
                       int w, int h, int blur_mode, double *score, oracle_detail *detail);

/* Debug taps: planar XYB of one RGB8 image at `scale` (caller provides 3*ws*hs floats). */
int oracle_xyb_at_scale(const uint8_t *rgb, int stride, int w, int h, int scale, float *xyb,
                        int *ws, int *hs);

/* ---- decoded-YUV -> RGB8 (libavif/libyuv integer path; yuv2rgb_oracle.c) ---------------- */

/* matrix: AV1 matrix_coefficients (1 = BT.709, 9 = BT.2020-NCL, 2/5/6 = BT.601).
 * depth: 8 (planes are u8) or 10 (planes are u16).  Full range only.
 * rgba_path: nonzero selects the arithmetic libavif runs when the image carries an alpha
 *   plane (io.zig:473 -> AVIF_RGB_FORMAT_RGBA); it differs from the RGB path for depth 10.
 * strides are in BYTES.  out: tight interleaved RGB8.  Returns 0 or <0 (unsupported). */
int oracle_yuv444_to_rgb8(const void *y, const void *u, const void *v,
                          int y_stride, int u_stride, int v_stride,
                          int w, int h, int depth, int matrix, int rgba_path, uint8_t *out);

/* Image.toRGB8 (src/io.zig:57-133): channels 1..4, 8- or 16-bit (native endian) -> RGB8. */
int oracle_to_rgb8(const void *data, int w, int h, int channels, int hbd, uint8_t *out);

/* encodeAvifToBuffer's per-pass sample conversions (src/io.zig:562-609) over n samples: 8 -> 10 bit
 * (v*1023+127)/255, 16 -> 10 bit v >> 6, 16 -> 8 bit v >> 8.  Returns 0, or -1 for 8 -> 8 (no conversion). */
int oracle_source_samples(const void *data, size_t n, int hbd, int out_depth, void *out);

#ifdef __cplusplus
}
#endif
#endif
