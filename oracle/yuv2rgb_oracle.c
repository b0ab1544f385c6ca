/*
 * oracle/yuv2rgb_oracle.c — CPU restatement of the decoded-YUV -> RGB8 step that feeds the
 * scorer.  TEST INFRASTRUCTURE ONLY (see ssimu2_oracle.h).
 *
 * Reference path: /root/reference/src/io.zig:452-482 (decodeAvifCommon: avifRGBImageSetDefaults,
 * rgb.depth = 8, format RGB or RGBA by presence of an alpha plane, avifImageYUVToRGB) followed by
 * io.zig:638-666 (decodeAvifToRgb: drop alpha, tight RGB8).  avifImageYUVToRGB is libavif, a
 * system C library that is not in /root/reference; with default settings on YUV444 full-range
 * input it dispatches to libyuv's fixed-point row functions.  This file restates that integer
 * arithmetic.  PINNED: tests/test_yuv_oracle.py checks it bit-for-bit against libavif 1.4.1
 * (the shared object Pillow bundles) and against tests/golden/yuv2rgb_*.npz generated from it.
 */
#include "ssimu2_oracle.h"

#include <stddef.h>

/* libyuv YuvConstants, full range ("J"/"F"/"V2020" sets): YG, YB, UB, UG, VG, VR. */
typedef struct { int yg, yb, ub, ug, vg, vr; } yuv_consts;

static int pick_consts(int matrix, yuv_consts *k)
{
    switch (matrix) {
    case 2: /* unspecified: libavif falls back to BT.601 */
    case 5:
    case 6:
        *k = (yuv_consts){16320, 32, 113, 22, 46, 90};
        return 0;
    case 1:
        *k = (yuv_consts){16320, 32, 119, 12, 30, 101};
        return 0;
    case 9:
        *k = (yuv_consts){16320, 32, 120, 11, 37, 94};
        return 0;
    default:
        return -1;
    }
}

static inline uint8_t clamp255(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

/* libyuv CALC_RGB16 + >>6 store: y32 is the 16-bit-expanded luma, u8/v8 the 8-bit chroma. */
static inline void yuv_pixel(uint32_t y32, int u8, int v8, const yuv_consts *k, uint8_t *rgb)
{
    const int y1 = (int)((y32 * (uint32_t)k->yg) >> 16) + k->yb;
    const int ui = u8 - 128, vi = v8 - 128;
    const int b16 = y1 + ui * k->ub;
    const int g16 = y1 - (ui * k->ug + vi * k->vg);
    const int r16 = y1 + vi * k->vr;
    rgb[0] = clamp255(r16 >> 6);
    rgb[1] = clamp255(g16 >> 6);
    rgb[2] = clamp255(b16 >> 6);
}

int oracle_yuv444_to_rgb8(const void *y, const void *u, const void *v,
                          int y_stride, int u_stride, int v_stride,
                          int w, int h, int depth, int matrix, int rgba_path, uint8_t *out)
{
    yuv_consts k;
    if (!y || !u || !v || !out || w <= 0 || h <= 0) return -1;
    if (pick_consts(matrix, &k) != 0) return -2;
    if (depth != 8 && depth != 10) return -3;
    for (int j = 0; j < h; ++j) {
        uint8_t *o = out + (size_t)j * w * 3;
        if (depth == 8) {
            const uint8_t *py = (const uint8_t *)y + (size_t)j * y_stride;
            const uint8_t *pu = (const uint8_t *)u + (size_t)j * u_stride;
            const uint8_t *pv = (const uint8_t *)v + (size_t)j * v_stride;
            /* I444ToRGB24Matrix / I444AlphaToARGBMatrix share YuvPixel: y32 = y * 0x0101 */
            for (int i = 0; i < w; ++i) yuv_pixel((uint32_t)py[i] * 0x0101u, pu[i], pv[i], &k, o + 3 * i);
        } else {
            const uint16_t *py = (const uint16_t *)((const uint8_t *)y + (size_t)j * y_stride);
            const uint16_t *pu = (const uint16_t *)((const uint8_t *)u + (size_t)j * u_stride);
            const uint16_t *pv = (const uint16_t *)((const uint8_t *)v + (size_t)j * v_stride);
            for (int i = 0; i < w; ++i) {
                const uint32_t Y = py[i] & 1023u, U = pu[i] & 1023u, V = pv[i] & 1023u;
                if (rgba_path) {
                    /* I410AlphaToARGBMatrix (YuvPixel10): luma stays 10-bit wide */
                    const uint32_t y32 = (Y << 6) | (Y >> 4);
                    const int u8 = (int)(U >> 2) > 255 ? 255 : (int)(U >> 2);
                    const int v8 = (int)(V >> 2) > 255 ? 255 : (int)(V >> 2);
                    yuv_pixel(y32, u8, v8, &k, o + 3 * i);
                } else {
                    /* no 10-bit -> RGB24 kernel in libyuv: libavif downshifts every plane
                     * to 8 bits (Convert16To8Plane, scale 16384 == >> 2) and runs the 8-bit path */
                    yuv_pixel((Y >> 2) * 0x0101u, (int)(U >> 2), (int)(V >> 2), &k, o + 3 * i);
                }
            }
        }
    }
    return 0;
}

/* Image.toRGB8, /root/reference/src/io.zig:57-133: 16-bit samples >> 8, alpha dropped,
 * gray replicated.  (The 8-bit RGB case is an alias of the input, main.zig:86.)            */
int oracle_to_rgb8(const void *data, int w, int h, int channels, int hbd, uint8_t *out)
{
    if (!data || !out || w <= 0 || h <= 0 || channels < 1 || channels > 4) return -1;
    const size_t n = (size_t)w * h;
    const uint8_t *s8 = (const uint8_t *)data;
    const uint16_t *s16 = (const uint16_t *)data;
    for (size_t i = 0; i < n; ++i) {
        uint8_t c[3];
        for (int k = 0; k < 3; ++k) {
            const size_t idx = i * channels + (channels >= 3 ? k : 0);
            c[k] = hbd ? (uint8_t)(s16[idx] >> 8) : s8[idx];
        }
        out[3 * i + 0] = c[0];
        out[3 * i + 1] = c[1];
        out[3 * i + 2] = c[2];
    }
    return 0;
}

/* encodeAvifToBuffer's sample conversions, /root/reference/src/io.zig:562-609, over n = w*h*channels samples:
 * 8-bit source at depth 10: (v*1023 + 127)/255 in usize (io.zig:572); 16-bit source at depth 10: v >> 6
 * (io.zig:587); 16-bit source at depth 8: v >> 8 (io.zig:602).  An 8-bit source at depth 8 is passed through
 * (io.zig:611-613): not a conversion, -1 here.  out: uint16_t for depth 10, uint8_t for depth 8.          */
int oracle_source_samples(const void *data, size_t n, int hbd, int out_depth, void *out)
{
    if (!data || !out) return -1;
    const uint8_t *s8 = (const uint8_t *)data;
    const uint16_t *s16 = (const uint16_t *)data;
    if (!hbd && out_depth == 10) {
        for (size_t i = 0; i < n; ++i) ((uint16_t *)out)[i] = (uint16_t)(((size_t)s8[i] * 1023 + 127) / 255);
    } else if (hbd && out_depth == 10) {
        for (size_t i = 0; i < n; ++i) ((uint16_t *)out)[i] = (uint16_t)(s16[i] >> 6);
    } else if (hbd && out_depth == 8) {
        for (size_t i = 0; i < n; ++i) ((uint8_t *)out)[i] = (uint8_t)(s16[i] >> 8);
    } else {
        return -1;
    }
    return 0;
}
