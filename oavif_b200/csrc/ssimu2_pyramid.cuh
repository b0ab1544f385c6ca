// ssimu2_pyramid.cuh — K0+K1+K2+K3 in one launch: decoded pixels -> six-scale XYB pyramid.
//
// Replaces, per image: avifImageYUVToRGB + repack (src/io.zig:470-478, 654-663), then the
// scorer's sRGB->linear, 2x2 downsample chain and linear->XYB at every scale (SURVEY.md
// Appendix A §1-§3).  One CTA owns a 64x64 scale-0 tile; because the 2x2 box pyramid of a
// 64-aligned tile only ever looks inside the tile (edge clamping replicates the tile's own
// last valid row/column), the CTA can walk all six scales without leaving registers/shared
// memory: linear RGB is never written to HBM.  HBM traffic per scale-0 pixel: 3..6 B read,
// 12 B * 1.333 written.
#pragma once

#include "ssimu2_common.cuh"
#include "ssimu2_tma.cuh"

namespace oavif {

constexpr int kPyrInline = 16;

struct PyrArgs {
    Geom g;
    const void *const *planes;  // device table, 3 pointers per image (RGB8 uses the first) — or null:
    const void *inl[3 * kPyrInline];  // up to kPyrInline images carry their pointers in the launch arguments
    long long stride[3];        // bytes per input row
    float *out;                 // pyramid of image 0
    long long out_stride;       // floats between consecutive images' pyramids
    const float *lut;           // sRGB->linear table, each of the 256 entries repeated 32 times (one copy per lane)
    YuvK k;
    int channels, hbd;          // IN_PIXELS only: 1..4 interleaved channels, 16-bit samples if hbd
    float one, neg_one;         // 1.0f / -1.0f as run-time values (Unit2, ssimu2_common.cuh)
};

// Four pixels of one row, as fetched: 3 words for the 8-bit layouts (RGB8: the 12 bytes r0 g0 b0 r1 ...;
// YUV8: y[0..3], u[0..3], v[0..3]), 6 words for 10-bit YUV (two 16-bit samples per word: y01 y23 u01 u23
// v01 v23).  fetch4 only loads (coordinates clamped to the image), decode4 only computes: the kernel
// issues all of a tile's loads before it waits for anything else.
template <int KIND>
struct PyrRaw {
    static constexpr int N = (KIND == IN_YUV10_RGB || KIND == IN_YUV10_RGBA) ? 6 : 3;
};

template <int KIND>
__device__ __forceinline__ void fetch4(const PyrArgs &a, const void *p0, const void *p1, const void *p2, int x0, int y,
                                       uint32_t *raw)
{
    const int w = a.g.w[0];
    const bool inside = (x0 + 3 < w);
    if (KIND == IN_RGB8) {
        const uint8_t *row = (const uint8_t *)p0 + (long long)y * a.stride[0];
        const uint8_t *q = row + 3 * x0;
        if (inside && ((reinterpret_cast<uintptr_t>(q) & 3) == 0)) {
            raw[0] = __ldg((const uint32_t *)q);
            raw[1] = __ldg((const uint32_t *)q + 1);
            raw[2] = __ldg((const uint32_t *)q + 2);
        } else {
            uint32_t b[12];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int x = min(x0 + i, w - 1);
                b[3 * i] = __ldg(row + 3 * x);
                b[3 * i + 1] = __ldg(row + 3 * x + 1);
                b[3 * i + 2] = __ldg(row + 3 * x + 2);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) raw[k] = b[4 * k] | (b[4 * k + 1] << 8) | (b[4 * k + 2] << 16) | (b[4 * k + 3] << 24);
        }
    } else if (KIND == IN_PIXELS) {
        // the loaders' layouts (io.zig:57-133): 16-bit samples keep their high byte, gray is replicated, alpha dropped
        const uint8_t *row = (const uint8_t *)p0 + (long long)y * a.stride[0];
        const int ch = a.channels, gstep = ch >= 3 ? 1 : 0, bstep = ch >= 3 ? 2 : 0;
        uint32_t b[12];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int x = min(x0 + i, w - 1);
            if (a.hbd) {
                const uint16_t *px = (const uint16_t *)row + (long long)x * ch;
                b[3 * i] = __ldg(px) >> 8;
                b[3 * i + 1] = __ldg(px + gstep) >> 8;
                b[3 * i + 2] = __ldg(px + bstep) >> 8;
            } else {
                const uint8_t *px = row + (long long)x * ch;
                b[3 * i] = __ldg(px);
                b[3 * i + 1] = __ldg(px + gstep);
                b[3 * i + 2] = __ldg(px + bstep);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) raw[k] = b[4 * k] | (b[4 * k + 1] << 8) | (b[4 * k + 2] << 16) | (b[4 * k + 3] << 24);
    } else if (KIND == IN_YUV8) {
        const uint8_t *r[3] = {(const uint8_t *)p0 + (long long)y * a.stride[0], (const uint8_t *)p1 + (long long)y * a.stride[1],
                               (const uint8_t *)p2 + (long long)y * a.stride[2]};
        const bool al = ((reinterpret_cast<uintptr_t>(r[0] + x0) | reinterpret_cast<uintptr_t>(r[1] + x0) |
                          reinterpret_cast<uintptr_t>(r[2] + x0)) & 3) == 0;
        if (inside && al) {
#pragma unroll
            for (int k = 0; k < 3; ++k) raw[k] = __ldg((const uint32_t *)(r[k] + x0));
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                uint32_t v = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) v |= (uint32_t)__ldg(r[k] + min(x0 + i, w - 1)) << (8 * i);
                raw[k] = v;
            }
        }
    } else {
        const uint16_t *r[3] = {(const uint16_t *)((const uint8_t *)p0 + (long long)y * a.stride[0]),
                                (const uint16_t *)((const uint8_t *)p1 + (long long)y * a.stride[1]),
                                (const uint16_t *)((const uint8_t *)p2 + (long long)y * a.stride[2])};
        const bool al = ((reinterpret_cast<uintptr_t>(r[0] + x0) | reinterpret_cast<uintptr_t>(r[1] + x0) |
                          reinterpret_cast<uintptr_t>(r[2] + x0)) & 7) == 0;
        if (inside && al) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint2 v = __ldg((const uint2 *)(r[k] + x0));
                raw[2 * k] = v.x;
                raw[2 * k + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                uint32_t s4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) s4[i] = __ldg(r[k] + min(x0 + i, w - 1));
                raw[2 * k] = s4[0] | (s4[1] << 16);
                raw[2 * k + 1] = s4[2] | (s4[3] << 16);
            }
        }
    }
}

template <int KIND>
__device__ __forceinline__ void decode4(const PyrArgs &a, const uint32_t *raw, int rgb[4][3])
{
    if (KIND == IN_RGB8 || KIND == IN_PIXELS) {
        const uint32_t w0 = raw[0], w1 = raw[1], w2 = raw[2];
        rgb[0][0] = w0 & 255; rgb[0][1] = (w0 >> 8) & 255; rgb[0][2] = (w0 >> 16) & 255;
        rgb[1][0] = w0 >> 24; rgb[1][1] = w1 & 255; rgb[1][2] = (w1 >> 8) & 255;
        rgb[2][0] = (w1 >> 16) & 255; rgb[2][1] = w1 >> 24; rgb[2][2] = w2 & 255;
        rgb[3][0] = (w2 >> 8) & 255; rgb[3][1] = (w2 >> 16) & 255; rgb[3][2] = w2 >> 24;
    } else if (KIND == IN_YUV8) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            yuv_to_rgb8<KIND>((raw[0] >> (8 * i)) & 255, (raw[1] >> (8 * i)) & 255, (raw[2] >> (8 * i)) & 255, a.k,
                              rgb[i][0], rgb[i][1], rgb[i][2]);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int sh = 16 * (i & 1), wi = i >> 1;
            yuv_to_rgb8<KIND>((raw[wi] >> sh) & 0xffff, (raw[2 + wi] >> sh) & 0xffff, (raw[4 + wi] >> sh) & 0xffff, a.k,
                              rgb[i][0], rgb[i][1], rgb[i][2]);
        }
    }
}

// grid = (tiles_x, tiles_y, n_images), block = 256 (16x16 threads, 4x4 scale-0 pixels each).
template <int KIND>
__global__ void __launch_bounds__(256, 4) k_pyramid(const __grid_constant__ PyrArgs a)
{
    // The sRGB table once per LANE: entry v of lane l lives at [v][l], so a warp's 32 gathers hit 32 different banks
    // whatever the pixel values are (a single 256-entry copy made 44 % of the kernel's shared-memory wavefronts
    // bank conflicts, profiles/r1_final_ncu_recursive.txt).  It arrives as one 32 KB bulk copy (no thread moves it).
    __shared__ __align__(128) float s_lut[256][32];
    __shared__ uint64_t s_lut_bar;
    // scale-2 linear RGB of this tile, rows of 16: thread t stores word t (no conflict), and the scale-3 step reads
    // it as 8-byte pairs with even and odd rows split between the two half-warps (see there)
    __shared__ __align__(8) float s_l2[3][16][16];
    __shared__ float s_l3[3][8][9];
    __shared__ float s_l4[3][4][5];

    const int tid = threadIdx.x;
    const Geom &g = a.g;
    const int tx = tid & 15, ty = tid >> 4;
    const int bx = blockIdx.x, by = blockIdx.y, img = blockIdx.z;
    const void *p0 = a.planes ? a.planes[3 * img + 0] : a.inl[3 * img + 0];
    const void *p1 = a.planes ? a.planes[3 * img + 1] : a.inl[3 * img + 1];
    const void *p2 = a.planes ? a.planes[3 * img + 2] : a.inl[3 * img + 2];
    float *pyr = a.out + (long long)img * a.out_stride;
    const XybConst kx = xyb_consts();
    const Unit2 un = unit2(a.one, a.neg_one);

    // ---- scale 0: 4x4 pixels per thread ------------------------------------------------
    // all of the thread's pixel loads are issued first, then the table: the two latencies overlap
    const int x0 = bx * 64 + tx * 4, y0 = by * 64 + ty * 4;
    uint32_t raw[4][PyrRaw<KIND>::N];
#pragma unroll
    for (int j = 0; j < 4; ++j) fetch4<KIND>(a, p0, p1, p2, x0, min(y0 + j, g.h[0] - 1), raw[j]);
    if (tid == 0) {
        mbar_init(&s_lut_bar, 1);
        mbar_init_fence();
    }
    __syncthreads();
    if (tid == 0) {   // the table is already laid out per lane in global memory (32 KB, L2-resident)
        mbar_arrive_expect_tx(&s_lut_bar, sizeof s_lut);
        bulk_load(&s_lut[0][0], a.lut, sizeof s_lut, &s_lut_bar);
    }
    mbar_wait(&s_lut_bar, 0);
    float lin[4][4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int rgb[4][3];
        decode4<KIND>(a, raw[j], rgb);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            lin[j][i][0] = s_lut[rgb[i][0]][tid & 31];
            lin[j][i][1] = s_lut[rgb[i][1]][tid & 31];
            lin[j][i][2] = s_lut[rgb[i][2]][tid & 31];
        }
    }
    {
        float *X = pyr + g.off[0], *Y = X + g.plane[0], *B = Y + g.plane[0];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float4 vx, vy, vb;
            f32x2 px, py, pb;   // pixels (0, 1) and (2, 3) of the row as packed pairs
            linear_to_xyb2(kx, un, pk2(lin[j][0][0], lin[j][1][0]), pk2(lin[j][0][1], lin[j][1][1]),
                           pk2(lin[j][0][2], lin[j][1][2]), px, py, pb);
            unpk2(px, vx.x, vx.y); unpk2(py, vy.x, vy.y); unpk2(pb, vb.x, vb.y);
            linear_to_xyb2(kx, un, pk2(lin[j][2][0], lin[j][3][0]), pk2(lin[j][2][1], lin[j][3][1]),
                           pk2(lin[j][2][2], lin[j][3][2]), px, py, pb);
            unpk2(px, vx.z, vx.w); unpk2(py, vy.z, vy.w); unpk2(pb, vb.z, vb.w);
            const long long o = (long long)(y0 + j) * g.pitch[0] + x0;
            *reinterpret_cast<float4 *>(X + o) = vx;
            *reinterpret_cast<float4 *>(Y + o) = vy;
            *reinterpret_cast<float4 *>(B + o) = vb;
        }
    }
    if (g.n_scales < 2) return;

    // ---- scale 1: 2x2 per thread, from registers -----------------------------------------
    float l1[2][2][3];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                l1[j][i][c] = box4(lin[2 * j][2 * i][c], lin[2 * j][2 * i + 1][c], lin[2 * j + 1][2 * i][c],
                                   lin[2 * j + 1][2 * i + 1][c]);
    {
        float *X = pyr + g.off[1], *Y = X + g.plane[1], *B = Y + g.plane[1];
        const int x1 = bx * 32 + tx * 2, y1 = by * 32 + ty * 2;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float2 vx, vy, vb;
            f32x2 px, py, pb;
            linear_to_xyb2(kx, un, pk2(l1[j][0][0], l1[j][1][0]), pk2(l1[j][0][1], l1[j][1][1]),
                           pk2(l1[j][0][2], l1[j][1][2]), px, py, pb);
            unpk2(px, vx.x, vx.y); unpk2(py, vy.x, vy.y); unpk2(pb, vb.x, vb.y);
            const long long o = (long long)(y1 + j) * g.pitch[1] + x1;
            *reinterpret_cast<float2 *>(X + o) = vx;
            *reinterpret_cast<float2 *>(Y + o) = vy;
            *reinterpret_cast<float2 *>(B + o) = vb;
        }
    }
    if (g.n_scales < 3) return;

    // ---- scale 2: 1 per thread; scale-1 coordinates clamp inside the thread's 2x2 --------
    {
        const int X2 = bx * 16 + tx, Y2 = by * 16 + ty;
        const int ib = (2 * X2 + 1 < g.w[1]) ? 1 : 0;
        const int jb = (2 * Y2 + 1 < g.h[1]) ? 1 : 0;
        float l2[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float p00 = l1[0][0][c];
            const float p01 = ib ? l1[0][1][c] : l1[0][0][c];
            const float p10 = jb ? l1[1][0][c] : l1[0][0][c];
            const float p11 = jb ? (ib ? l1[1][1][c] : l1[1][0][c]) : (ib ? l1[0][1][c] : l1[0][0][c]);
            l2[c] = box4(p00, p01, p10, p11);
            s_l2[c][ty][tx] = l2[c];
        }
        float vx, vy, vb;
        linear_to_xyb(kx, l2[0], l2[1], l2[2], vx, vy, vb);
        const long long o = (long long)Y2 * g.pitch[2] + X2;
        float *X = pyr + g.off[2];
        X[o] = vx;
        X[o + g.plane[2]] = vy;
        X[o + 2 * g.plane[2]] = vb;
    }
    if (g.n_scales < 4) return;
    __syncthreads();

    // ---- scale 3: 8x8 per tile, from shared memory; clamp to the tile-local last valid ---
    // A cell reads its two scale-2 rows as 8-byte pairs.  Rows of 16 words put even rows in banks 0..15 and odd rows
    // in 16..31, so within a half-warp (two cell rows cy, eight cx each) the cells of even cy read their upper row
    // first and those of odd cy their lower row first: every access covers all 32 banks once.
    if (tid < 64) {
        const int cx = tid & 7, cy = tid >> 3;
        const bool two_x = 2 * cx + 1 <= g.w[2] - 1 - bx * 16;      // else the cell's right column is clamped onto its left
        const int ly = max(min(2 * cy + 1, g.h[2] - 1 - by * 16), 0);
        const bool swap = cy & 1;
        const int ra = swap ? ly : 2 * cy, rb = swap ? 2 * cy : ly;
        float l3[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float2 va = *reinterpret_cast<const float2 *>(&s_l2[c][ra][2 * cx]);
            const float2 vb = *reinterpret_cast<const float2 *>(&s_l2[c][rb][2 * cx]);
            const float2 top = swap ? vb : va, bot = swap ? va : vb;
            l3[c] = box4(top.x, two_x ? top.y : top.x, bot.x, two_x ? bot.y : bot.x);
            s_l3[c][cy][cx] = l3[c];
        }
        float vx, vy, vb;
        linear_to_xyb(kx, l3[0], l3[1], l3[2], vx, vy, vb);
        const long long o = (long long)(by * 8 + cy) * g.pitch[3] + bx * 8 + cx;
        float *X = pyr + g.off[3];
        X[o] = vx;
        X[o + g.plane[3]] = vy;
        X[o + 2 * g.plane[3]] = vb;
    }
    if (g.n_scales < 5) return;
    __syncthreads();

    if (tid < 16) {
        const int cx = tid & 3, cy = tid >> 2;
        const int lx = max(min(2 * cx + 1, g.w[3] - 1 - bx * 8), 0);
        const int ly = max(min(2 * cy + 1, g.h[3] - 1 - by * 8), 0);
        float l4[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            l4[c] = box4(s_l3[c][2 * cy][2 * cx], s_l3[c][2 * cy][lx], s_l3[c][ly][2 * cx], s_l3[c][ly][lx]);
            s_l4[c][cy][cx] = l4[c];
        }
        float vx, vy, vb;
        linear_to_xyb(kx, l4[0], l4[1], l4[2], vx, vy, vb);
        const long long o = (long long)(by * 4 + cy) * g.pitch[4] + bx * 4 + cx;
        float *X = pyr + g.off[4];
        X[o] = vx;
        X[o + g.plane[4]] = vy;
        X[o + 2 * g.plane[4]] = vb;
    }
    if (g.n_scales < 6) return;
    __syncthreads();

    if (tid < 4) {
        const int cx = tid & 1, cy = tid >> 1;
        const int lx = max(min(2 * cx + 1, g.w[4] - 1 - bx * 4), 0);
        const int ly = max(min(2 * cy + 1, g.h[4] - 1 - by * 4), 0);
        float l5[3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
            l5[c] = box4(s_l4[c][2 * cy][2 * cx], s_l4[c][2 * cy][lx], s_l4[c][ly][2 * cx], s_l4[c][ly][lx]);
        float vx, vy, vb;
        linear_to_xyb(kx, l5[0], l5[1], l5[2], vx, vy, vb);
        const long long o = (long long)(by * 2 + cy) * g.pitch[5] + bx * 2 + cx;
        float *X = pyr + g.off[5];
        X[o] = vx;
        X[o + g.plane[5]] = vy;
        X[o + 2 * g.plane[5]] = vb;
    }
}

// decodeAvifToRgb's pixel work alone (io.zig:470-478, 654-663): planes -> tight RGB8.
struct Yuv2RgbArgs {
    const void *y, *u, *v;
    long long stride[3];
    int w, h;
    uint8_t *out;
    YuvK k;
};

template <int KIND>
__global__ void __launch_bounds__(256) k_yuv_to_rgb8(const __grid_constant__ Yuv2RgbArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= a.w || y >= a.h) return;
    uint32_t Y, U, V;
    if (KIND == IN_YUV8) {
        Y = ((const uint8_t *)a.y + (long long)y * a.stride[0])[x];
        U = ((const uint8_t *)a.u + (long long)y * a.stride[1])[x];
        V = ((const uint8_t *)a.v + (long long)y * a.stride[2])[x];
    } else {
        Y = ((const uint16_t *)((const uint8_t *)a.y + (long long)y * a.stride[0]))[x];
        U = ((const uint16_t *)((const uint8_t *)a.u + (long long)y * a.stride[1]))[x];
        V = ((const uint16_t *)((const uint8_t *)a.v + (long long)y * a.stride[2]))[x];
    }
    int r, g, b;
    yuv_to_rgb8<KIND>(Y, U, V, a.k, r, g, b);
    uint8_t *o = a.out + ((long long)y * a.w + x) * 3;
    o[0] = (uint8_t)r;
    o[1] = (uint8_t)g;
    o[2] = (uint8_t)b;
}

// The sample array encodeAvifToBuffer builds for avifImageRGBToYUV (io.zig:562-609), from the staged source pixels:
// MODE 0: 8-bit -> 10-bit, (v * 1023 + 127) / 255 (io.zig:572); MODE 1: 16-bit -> 10-bit, v >> 6 (io.zig:587);
// MODE 2: 16-bit -> 8-bit, v >> 8 (io.zig:602).  Every channel is kept; input rows `stride` bytes apart, output tight.
struct SamplesArgs {
    const void *in;
    void *out;
    long long stride;      // input bytes per row
    int row_samples, h;    // w * channels
};

template <int MODE>
__global__ void __launch_bounds__(256) k_source_samples(const __grid_constant__ SamplesArgs a)
{
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;   // four samples per thread
    const int y = blockIdx.y;
    if (x >= a.row_samples) return;
    const int n = min(4, a.row_samples - x);
    const uint8_t *row = (const uint8_t *)a.in + (long long)y * a.stride;
    const long long o = (long long)y * a.row_samples + x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (i >= n) break;
        if (MODE == 0) ((uint16_t *)a.out)[o + i] = (uint16_t)(((uint32_t)row[x + i] * 1023u + 127u) / 255u);
        else if (MODE == 1) ((uint16_t *)a.out)[o + i] = (uint16_t)(((const uint16_t *)row)[x + i] >> 6);
        else ((uint8_t *)a.out)[o + i] = (uint8_t)(((const uint16_t *)row)[x + i] >> 8);
    }
}

}  // namespace oavif
