// ssimu2_fir.cuh — K4+K5 fused, FIR form: blur of {a, b, a^2, b^2, ab} + SSIM / edge-diff maps +
// 1-norm / 4-norm pooling for one 64x32 tile of one XYB channel, with no intermediate plane in HBM.
//
// The sigma = 1.5 recursive Gaussian of SSIMULACRA2 v2.1 has an exactly finite impulse response
// (2N-1 = 9 taps); this kernel evaluates that 9-tap filter directly,
// horizontal then vertical, zero padded, accumulating left->right / top->bottom with fmaf — the
// same sequence as the oracle's FIR mode.  HBM traffic: the two XYB planes once (plus halo, served
// by L2), six doubles per CTA out.
#pragma once

#include "ssimu2_common.cuh"

namespace oavif {

constexpr int kFirTW = 64, kFirTH = 32, kFirR = 4;          // tile and filter radius
constexpr int kFirIW = kFirTW + 2 * kFirR;                  // 72 staged columns
constexpr int kFirIH = kFirTH + 2 * kFirR;                  // 40 staged rows
constexpr int kFirThreads = 256;
// shared memory: staged (a, b) interleaved per pixel [40][72][2]; row-filtered pairs (a,b) and (a*a,b*b)
// interleaved [2][40][64][2] and a*b [40][64]; the reduction scratch
constexpr size_t kFirSmemBytes =
    (size_t)(2 * kFirIH * kFirIW + 5 * kFirIH * kFirTW) * sizeof(float) + 8 * 6 * sizeof(double);

struct BlurArgs {
    Geom g;
    const float *src;        // source pyramid (shared by all candidates)
    const float *dist;       // pyramid of candidate 0
    long long dist_stride;   // floats between candidates
    double *partials;        // [candidate][cta][6]
    long long partials_stride;  // doubles between candidates
    int first_cta[kMaxScales + 1];  // CTA index ranges per scale: 3 channels x tiles each
    int tiles_x[kMaxScales], tiles_y[kMaxScales];
    float taps[9];
    float one, neg_one;      // 1.0f / -1.0f as run-time values (Unit2, ssimu2_common.cuh)
};

__device__ __forceinline__ float fir9(const float *v, const float *t)
{
    float acc = t[0] * v[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) acc = fmaf(t[k], v[k], acc);
    return acc;
}

// the same accumulation on a packed pair (two quantities of one pixel)
__device__ __forceinline__ f32x2 fir9x2(const f32x2 *v, const f32x2 *t)
{
    f32x2 acc = mul2(t[0], v[0]);
#pragma unroll
    for (int k = 1; k < 9; ++k) acc = fma2(t[k], v[k], acc);
    return acc;
}

// grid = (total CTAs over scales/channels/tiles, n_candidates), block = 256, dynamic smem.
//
// The two planes are staged INTERLEAVED, (a, b) per pixel, so that every later load delivers aligned
// register pairs: the horizontal pass filters the pairs (a, b) and (a*a, b*b) with packed FFMA2 — each
// half is the scalar fmaf chain — and a*b alone; the vertical pass does the same on the row-filtered
// pairs; the maps take the five values of a pixel as they come.
__global__ void __launch_bounds__(kFirThreads, 3) k_fir_fused(const __grid_constant__ BlurArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sab = reinterpret_cast<float *>(smem_raw);      // [40][72][2]  (a, b)
    float *shp = sab + 2 * kFirIH * kFirIW;                // [2][40][64][2]  rows pass of (a, b), (a*a, b*b)
    float *shs = shp + 4 * kFirIH * kFirTW;                // [40][64]        rows pass of a*b
    double *sred = reinterpret_cast<double *>(shs + kFirIH * kFirTW);

    const int tid = threadIdx.x;
    const int cta = blockIdx.x, cand = blockIdx.y;
    int s = 0;
#pragma unroll
    for (int i = 1; i < kMaxScales; ++i)
        if (i < a.g.n_scales && cta >= a.first_cta[i]) s = i;
    const int local = cta - a.first_cta[s];
    const int ntile = a.tiles_x[s] * a.tiles_y[s];
    const int c = local / ntile;
    const int t = local - c * ntile;
    const int tyi = t / a.tiles_x[s], txi = t - tyi * a.tiles_x[s];
    const int x0 = txi * kFirTW, y0 = tyi * kFirTH;
    const int w = a.g.w[s], h = a.g.h[s], pitch = a.g.pitch[s];
    const float *pa = a.src + a.g.off[s] + (long long)c * a.g.plane[s];
    const float *pb = a.dist + (long long)cand * a.dist_stride + a.g.off[s] + (long long)c * a.g.plane[s];

    float taps[9];
    f32x2 taps2[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        taps[k] = a.taps[k];
        taps2[k] = splat2(taps[k]);
    }

    // Shared-memory rows are stored in 16-byte chunks (two interleaved pixels); on ODD rows adjacent chunks
    // are swapped (physical chunk = logical chunk ^ 1).  With eight-lane groups made of four consecutive
    // chunk pairs on an even row and the same four on the next row, the eight 16-byte accesses of a phase
    // then land on eight different bank groups (even chunks on the even row, odd ones on the odd row); without
    // the swap every 128-bit access of the interleaved layout is a 2-way bank conflict.  The swap costs no
    // instructions: even and odd logical chunks just use two base pointers, `+odd` and `-odd` chunks away.

    // ---- stage a, b with a 4-pixel zero-padded halo (float4 granules), interleaved per pixel -------
    // items: 20 row pairs x 5 groups x 8 lanes (4 granules x 2 rows); 18 granules per row, so the last group is half empty
    for (int idx = tid; idx < (kFirIH / 2) * 5 * 8; idx += kFirThreads) {
        const int l8 = idx & 7, j = idx >> 3;
        const int grp = j % 5, rowpair = j / 5;
        const int c4 = 4 * grp + (l8 & 3), odd = l8 >> 2, row = 2 * rowpair + odd;
        if (c4 >= kFirIW / 4) continue;
        const int gx = x0 - kFirR + 4 * c4, gy = y0 - kFirR + row;
        float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
        if (gy >= 0 && gy < h && gx >= 0 && gx < w) {
            const long long o = (long long)gy * pitch + gx;
            va = __ldg(reinterpret_cast<const float4 *>(pa + o));
            vb = __ldg(reinterpret_cast<const float4 *>(pb + o));
            if (gx + 3 >= w) {  // right edge: samples at x >= w are padding, not data
                if (gx + 1 >= w) { va.y = 0.f; vb.y = 0.f; }
                if (gx + 2 >= w) { va.z = 0.f; vb.z = 0.f; }
                va.w = 0.f; vb.w = 0.f;
            }
        }
        float4 *dst = reinterpret_cast<float4 *>(sab + 2 * (row * kFirIW + 4 * c4));   // logical chunks 2*c4, 2*c4 + 1
        dst[odd] = make_float4(va.x, vb.x, va.y, vb.y);
        dst[1 - odd] = make_float4(va.z, vb.z, va.w, vb.w);
    }
    __syncthreads();

    // ---- horizontal pass: 4 outputs per item, 12-sample window of (a, b) pairs -----------------------
    // items: 20 row pairs x 4 groups x 8 lanes (4 output granules x 2 rows)
    for (int idx = tid; idx < kFirIH * (kFirTW / 4); idx += kFirThreads) {
        const int l8 = idx & 7, j = idx >> 3;
        const int g4 = 4 * (j & 3) + (l8 & 3), odd = l8 >> 2, row = 2 * (j >> 2) + odd;
        f32x2 v[12], q[12];
        float p[12];
        const float4 *r4 = reinterpret_cast<const float4 *>(sab + 2 * (row * kFirIW + 4 * g4));
        const float4 *re = r4 + odd, *ro = r4 - odd;   // even / odd logical chunks of this row
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const float4 x = (k & 1) ? ro[k] : re[k];
            v[2 * k] = pk2(x.x, x.y);
            v[2 * k + 1] = pk2(x.z, x.w);
            p[2 * k] = x.x * x.y;
            p[2 * k + 1] = x.z * x.w;
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) q[k] = mul2(v[k], v[k]);
        float4 *d0 = reinterpret_cast<float4 *>(shp + 2 * (row * kFirTW + 4 * g4));
        float4 *d1 = reinterpret_cast<float4 *>(shp + 2 * (kFirIH * kFirTW + row * kFirTW + 4 * g4));
        f32x2 o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = fir9x2(v + i, taps2);
        float4 w4;
        unpk2(o[0], w4.x, w4.y); unpk2(o[1], w4.z, w4.w); d0[odd] = w4;
        unpk2(o[2], w4.x, w4.y); unpk2(o[3], w4.z, w4.w); d0[1 - odd] = w4;
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = fir9x2(q + i, taps2);
        unpk2(o[0], w4.x, w4.y); unpk2(o[1], w4.z, w4.w); d1[odd] = w4;
        unpk2(o[2], w4.x, w4.y); unpk2(o[3], w4.z, w4.w); d1[1 - odd] = w4;
        w4.x = fir9(p, taps); w4.y = fir9(p + 1, taps); w4.z = fir9(p + 2, taps); w4.w = fir9(p + 3, taps);
        *reinterpret_cast<float4 *>(shs + row * kFirTW + 4 * g4) = w4;
    }
    __syncthreads();

    // ---- vertical pass + maps: one column, 8 output rows per thread ---------------------------
    const int col = tid & (kFirTW - 1), rg = tid >> 6;  // 4 row groups of 8
    f32x2 om[8], os[8];   // (mu1, mu2), (s11, s22)
    float ox[8];          // s12
    {
        f32x2 v[16];
        // window row k of this thread is tile row rg*8 + k: its parity is that of k.  Pixel `col` lives in
        // logical chunk col >> 1; on odd rows the chunk is the neighbouring one.
        const float *p = shp + 2 * ((rg * 8) * kFirTW) + 4 * (col >> 1) + 2 * (col & 1);
        const int swap = (col & 2) ? -4 : 4;   // floats from a chunk to its partner
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = lds2(p + 2 * k * kFirTW + ((k & 1) ? swap : 0));
#pragma unroll
        for (int r = 0; r < 8; ++r) om[r] = fir9x2(v + r, taps2);
        p += 2 * kFirIH * kFirTW;
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = lds2(p + 2 * k * kFirTW + ((k & 1) ? swap : 0));
#pragma unroll
        for (int r = 0; r < 8; ++r) os[r] = fir9x2(v + r, taps2);
        float u[16];
        const float *ps = shs + (rg * 8) * kFirTW + col;
#pragma unroll
        for (int k = 0; k < 16; ++k) u[k] = ps[k * kFirTW];
#pragma unroll
        for (int r = 0; r < 8; ++r) ox[r] = fir9(u + r, taps);
    }
    // maps on packed pairs of rows (r, r+1) of this column; a row below the image enters as all-zero
    // inputs, which pool to exactly zero
    const Unit2 un = unit2(a.one, a.neg_one);
    f32x2 acc[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) acc[j] = splat2(0.0f);
    const int gx = x0 + col;
#pragma unroll
    for (int r = 0; r < 8; r += 2) {
        const int gy = y0 + rg * 8 + r;
        if (gx < w && gy < h) {
            // staged row rg*8 + r + 4: parity of r (r is even here, r + 1 odd); pixel col + 4
            const int px = col + kFirR, sw = (px & 2) ? -4 : 4;
            const float2 ab0 = *reinterpret_cast<const float2 *>(sab + 2 * ((rg * 8 + r + kFirR) * kFirIW) + 4 * (px >> 1) +
                                                                 2 * (px & 1));
            const float2 ab1 = *reinterpret_cast<const float2 *>(sab + 2 * ((rg * 8 + r + 1 + kFirR) * kFirIW) +
                                                                 4 * (px >> 1) + 2 * (px & 1) + sw);
            float m1a, m2a, m1b, m2b, s1a, s2a, s1b, s2b;
            unpk2(om[r], m1a, m2a);
            unpk2(om[r + 1], m1b, m2b);
            unpk2(os[r], s1a, s2a);
            unpk2(os[r + 1], s1b, s2b);
            const bool two = gy + 1 < h;
            error_maps2(un, pk2(ab0.x, two ? ab1.x : 0.f), pk2(ab0.y, two ? ab1.y : 0.f), pk2(m1a, two ? m1b : 0.f),
                        pk2(m2a, two ? m2b : 0.f), pk2(s1a, two ? s1b : 0.f), pk2(s2a, two ? s2b : 0.f),
                        pk2(ox[r], two ? ox[r + 1] : 0.f), acc);
        }
    }
    double dacc[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        float lo, hi;
        unpk2(acc[j], lo, hi);
        dacc[j] = (double)lo + (double)hi;
    }
    block_reduce6<kFirThreads / 32>(dacc, sred,
                                    a.partials + (long long)cand * a.partials_stride + (long long)cta * 6);
}

}  // namespace oavif
