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
};

__device__ __forceinline__ float fir9(const float *v, const float *t)
{
    float acc = t[0] * v[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) acc = fmaf(t[k], v[k], acc);
    return acc;
}

// grid = (total CTAs over scales/channels/tiles, n_candidates), block = 256, dynamic smem.
__global__ void __launch_bounds__(kFirThreads, 3) k_fir_fused(const __grid_constant__ BlurArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sa = reinterpret_cast<float *>(smem_raw);       // [40][72]
    float *sb = sa + kFirIH * kFirIW;                      // [40][72]
    float *sh = sb + kFirIH * kFirIW;                      // [5][40][64]
    double *sred = reinterpret_cast<double *>(sh + 5 * kFirIH * kFirTW);

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
#pragma unroll
    for (int k = 0; k < 9; ++k) taps[k] = a.taps[k];

    // ---- stage a, b with a 4-pixel zero-padded halo (float4 granules) ----------------------
    for (int idx = tid; idx < kFirIH * (kFirIW / 4); idx += kFirThreads) {
        const int row = idx / (kFirIW / 4), c4 = idx - row * (kFirIW / 4);
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
        *reinterpret_cast<float4 *>(sa + row * kFirIW + 4 * c4) = va;
        *reinterpret_cast<float4 *>(sb + row * kFirIW + 4 * c4) = vb;
    }
    __syncthreads();

    // ---- horizontal pass: 4 outputs per item, 12-sample register window ---------------------
    for (int idx = tid; idx < kFirIH * (kFirTW / 4); idx += kFirThreads) {
        const int row = idx / (kFirTW / 4), g4 = idx - row * (kFirTW / 4);
        float va[12], vb[12], q[12];
        const float4 *ra = reinterpret_cast<const float4 *>(sa + row * kFirIW + 4 * g4);
        const float4 *rb = reinterpret_cast<const float4 *>(sb + row * kFirIW + 4 * g4);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float4 x = ra[k], y = rb[k];
            va[4 * k] = x.x; va[4 * k + 1] = x.y; va[4 * k + 2] = x.z; va[4 * k + 3] = x.w;
            vb[4 * k] = y.x; vb[4 * k + 1] = y.y; vb[4 * k + 2] = y.z; vb[4 * k + 3] = y.w;
        }
        float4 o;
        float *dst = sh + row * kFirTW + 4 * g4;
        o.x = fir9(va, taps); o.y = fir9(va + 1, taps); o.z = fir9(va + 2, taps); o.w = fir9(va + 3, taps);
        *reinterpret_cast<float4 *>(dst) = o;
        o.x = fir9(vb, taps); o.y = fir9(vb + 1, taps); o.z = fir9(vb + 2, taps); o.w = fir9(vb + 3, taps);
        *reinterpret_cast<float4 *>(dst + kFirIH * kFirTW) = o;
#pragma unroll
        for (int k = 0; k < 12; ++k) q[k] = va[k] * va[k];
        o.x = fir9(q, taps); o.y = fir9(q + 1, taps); o.z = fir9(q + 2, taps); o.w = fir9(q + 3, taps);
        *reinterpret_cast<float4 *>(dst + 2 * kFirIH * kFirTW) = o;
#pragma unroll
        for (int k = 0; k < 12; ++k) q[k] = vb[k] * vb[k];
        o.x = fir9(q, taps); o.y = fir9(q + 1, taps); o.z = fir9(q + 2, taps); o.w = fir9(q + 3, taps);
        *reinterpret_cast<float4 *>(dst + 3 * kFirIH * kFirTW) = o;
#pragma unroll
        for (int k = 0; k < 12; ++k) q[k] = va[k] * vb[k];
        o.x = fir9(q, taps); o.y = fir9(q + 1, taps); o.z = fir9(q + 2, taps); o.w = fir9(q + 3, taps);
        *reinterpret_cast<float4 *>(dst + 4 * kFirIH * kFirTW) = o;
    }
    __syncthreads();

    // ---- vertical pass + maps: one column, 8 output rows per thread ---------------------------
    const int col = tid & (kFirTW - 1), rg = tid >> 6;  // 4 row groups of 8
    float out[5][8];
#pragma unroll
    for (int qn = 0; qn < 5; ++qn) {
        float v[16];
        const float *p = sh + qn * kFirIH * kFirTW + (rg * 8) * kFirTW + col;
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = p[k * kFirTW];
#pragma unroll
        for (int r = 0; r < 8; ++r) out[qn][r] = fir9(v + r, taps);
    }
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int gx = x0 + col;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int gy = y0 + rg * 8 + r;
        if (gx < w && gy < h) {
            const float av = sa[(rg * 8 + r + kFirR) * kFirIW + col + kFirR];
            const float bv = sb[(rg * 8 + r + kFirR) * kFirIW + col + kFirR];
            error_maps(av, bv, out[0][r], out[1][r], out[2][r], out[3][r], out[4][r], acc);
        }
    }
    double dacc[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) dacc[j] = (double)acc[j];
    block_reduce6<kFirThreads / 32>(dacc, sred,
                                    a.partials + (long long)cand * a.partials_stride + (long long)cta * 6);
}

}  // namespace oavif
