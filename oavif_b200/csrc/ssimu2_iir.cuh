// ssimu2_iir.cuh — K4+K5, RECURSIVE form: the sigma = 1.5 recursive Gaussian of SSIMULACRA2 v2.1
// (Charalampidis 2016, three undamped oscillators k in {1,3,5}, radius N = 5; SURVEY.md Appendix A
// §4) evaluated exactly as published — every row and every column is ONE serial binary32 chain
// from n = -N+1, because the recursion's round-off is not negligible at the metric's C2 = 9e-4
// scale and is therefore part of the published result.
//
//   k_iir_rows  : rows.  One CTA = 32 rows of one channel; warp q owns quantity q of
//                 {a, b, a^2, b^2, ab}; lane = row.  Pixels arrive as 32x32 tiles via cp.async into
//                 a padded shared-memory ring (coalesced 128-byte row reads, conflict-free
//                 transposed reads), results leave through a per-warp transposed staging tile.
//   k_iir_cols  : columns + error maps + pooling.  One CTA = 32 columns; warp q owns quantity q;
//                 lane = column, so every global access is a 128-byte row segment.  The five
//                 filtered values of a pixel meet in shared memory every 20 rows and go straight
//                 into the SSIM / edge-diff maps — blurred planes are never written to HBM.
//
// HBM traffic per scale pixel and channel: rows pass reads 8 B, writes 20 B; columns pass reads
// 20 B + 8 B.  No tensor cores (nothing here is a contraction).
#pragma once

#include "ssimu2_common.cuh"

namespace oavif {

struct IirCoef {
    float n2[3], d1[3];
};

struct IirState {
    float p[3], p2[3];
};

// One step of FastGaussian1D: out_k = n2_k*(l+r) - prev2_k - d1_k*prev_k, evaluated as
// sum*n2, minus prev2, fma(-d1, prev, .); output = (o1 + o3) + o5.
__device__ __forceinline__ float iir_step(const IirCoef &k, IirState &s, float l, float r)
{
    const float sum = l + r;
    float o[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float ok = sum * k.n2[i];
        ok = ok - s.p2[i];
        ok = fmaf(-k.d1[i], s.p[i], ok);
        s.p2[i] = s.p[i];
        s.p[i] = ok;
        o[i] = ok;
    }
    return (o[0] + o[1]) + o[2];
}

__device__ __forceinline__ void cp_async_f32(float *smem_dst, const float *gmem_src, bool valid)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 4 : 0;  // src-size 0 => the 4 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem_src), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

constexpr int kIirRows = 64;    // rows per CTA in the rows pass (two per lane: 6 independent chains)
constexpr int kIirChunk = 32;   // columns per staged tile
constexpr int kIirSlots = 4;    // tile ring: t-1, t in use while t+1, t+2 land
constexpr int kIirThreads = 160;
constexpr int kIirVCols = 32;   // columns per CTA in the columns pass
constexpr int kIirVBatch = 20;  // rows between two map phases (20*32 px = 4 per thread)

struct IirArgs {
    Geom g;
    IirCoef k;
    const float *src;
    const float *dist;
    long long dist_stride;      // floats between candidates' pyramids
    float *hplanes;             // [candidate][quantity][pyramid layout]
    long long hplanes_stride;   // floats between candidates (= 5 * q_stride)
    long long q_stride;         // floats between quantities
    double *partials;
    long long partials_stride;
    int first_cta[kMaxScales + 1];
    int blocks[kMaxScales];     // row blocks (rows pass) or column blocks (columns pass) per channel
};

__device__ __forceinline__ void decode_cta(const IirArgs &a, int cta, int &s, int &c, int &blk)
{
    s = 0;
#pragma unroll
    for (int i = 1; i < kMaxScales; ++i)
        if (i < a.g.n_scales && cta >= a.first_cta[i]) s = i;
    const int local = cta - a.first_cta[s];
    c = local / a.blocks[s];
    blk = local - c * a.blocks[s];
}

// ------------------------------------------------------------------------------------------------
// rows pass.  grid = (sum over scales of 3 * ceil(h/64), n_candidates), block = 160, dynamic smem.
//
// Chunk t emits outputs n = 32t-4 .. 32t+27, so its right taps (n+4) are exactly tile t and its left
// taps (n-6) fall in tiles t-1 and t: two tiles are in use while tiles t+1 and t+2 are in flight
// (cp.async, two chunks of slack).  Chunk 0's first four outputs are the recursion's warm-up steps
// n = -4..-1 and are dropped.
struct IirRowsSmem {
    float ta[kIirSlots][kIirRows][kIirChunk + 1];  // source channel tiles, ring
    float tb[kIirSlots][kIirRows][kIirChunk + 1];  // distorted channel tiles, ring
    float to[5][kIirRows][kIirChunk + 1];          // per-warp output staging (transposed write-out)
};

// quantity q of {a, b, a*a, b*b, a*b} = x * (y*m + o) with warp-uniform tile pointers for x, y and
// (m, o) = (0, 1) for the two plain planes, (1, 0) for the products.  y*0+1 and x*1 are exact, so
// every quantity is bit-identical to the direct expression, with no divergent code.
__device__ __forceinline__ float pick_quantity(const float *px, const float *py, int j, float m, float o)
{
    return px[j] * fmaf(py[j], m, o);
}

__global__ void __launch_bounds__(kIirThreads, 2) k_iir_rows(const __grid_constant__ IirArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IirRowsSmem &sm = *reinterpret_cast<IirRowsSmem *>(smem_raw);

    int s, c, rb;
    decode_cta(a, blockIdx.x, s, c, rb);
    const int cand = blockIdx.y;
    const int w = a.g.w[s], h = a.g.h[s], pitch = a.g.pitch[s];
    const int y0 = rb * kIirRows;
    const long long poff = a.g.off[s] + (long long)c * a.g.plane[s];
    const float *pa = a.src + poff;
    const float *pb = a.dist + (long long)cand * a.dist_stride + poff;
    const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
    float *ph = a.hplanes + (long long)cand * a.hplanes_stride + (long long)q * a.q_stride + poff;
    const int nch = (w + kIirChunk - 1) / kIirChunk;          // input tiles
    const int nout = (w + 4 + kIirChunk - 1) / kIirChunk;     // output chunks
    const IirCoef k = a.k;
    const bool x_is_b = (q == 1) || (q == 3);   // x: b for {b, b*b}, a otherwise
    const bool y_is_a = (q == 2);                 // y: a for a*a, b for {b*b, a*b}; unused for q < 2
    const float ym = q < 2 ? 0.0f : 1.0f, yo = q < 2 ? 1.0f : 0.0f;

    // tile t of both planes -> ring slot t & 3.  Warp q stages rows q, q+5, ... of both planes.
    // Anything outside the image (t < 0, t >= nch, row >= h, column >= w) lands as zeros: that IS
    // the filter's zero padding.
    const int rows_here = min(kIirRows, h - y0);
    auto issue_tile = [&](int t) {
        const int slot = t & (kIirSlots - 1);
        const int gx = t * kIirChunk + lane;
        const bool col_ok = (t >= 0) && (gx < w);
        const long long o0 = (long long)y0 * pitch + gx;
#pragma unroll 1
        for (int row = q; row < kIirRows; row += 5) {
            const bool valid = col_ok && (row < rows_here);
            const long long o = valid ? o0 + (long long)row * pitch : 0;
            cp_async_f32(&sm.ta[slot][row][lane], pa + o, valid);
            cp_async_f32(&sm.tb[slot][row][lane], pb + o, valid);
        }
        cp_async_commit();
    };

#pragma unroll 1
    for (int t = -1; t <= 1; ++t) issue_tile(t);
    cp_async_wait_all();
    __syncthreads();

    IirState st0, st1;  // rows lane and lane + 32
#pragma unroll
    for (int i = 0; i < 3; ++i) st0.p[i] = st0.p2[i] = st1.p[i] = st1.p2[i] = 0.0f;

    for (int t = 0; t < nout; ++t) {
        issue_tile(t + 2);
        const int cur = t & (kIirSlots - 1), prev = (t + kIirSlots - 1) & (kIirSlots - 1);
        const float(*tx)[kIirRows][kIirChunk + 1] = x_is_b ? sm.tb : sm.ta;
        const float(*ty)[kIirRows][kIirChunk + 1] = y_is_a ? sm.ta : sm.tb;
        const float *x0 = &tx[cur][lane][0], *y0p = &ty[cur][lane][0];
        const float *x1 = &tx[cur][lane + 32][0], *y1p = &ty[cur][lane + 32][0];
        const float *px0 = &tx[prev][lane][0], *py0 = &ty[prev][lane][0];
        const float *px1 = &tx[prev][lane + 32][0], *py1 = &ty[prev][lane + 32][0];
        float *o0 = &sm.to[q][lane][0], *o1 = &sm.to[q][lane + 32][0];
        float r0[kIirChunk], r1[kIirChunk];
#pragma unroll
        for (int j = 0; j < kIirChunk; ++j) {
            r0[j] = pick_quantity(x0, y0p, j, ym, yo);
            r1[j] = pick_quantity(x1, y1p, j, ym, yo);
        }
#pragma unroll
        for (int j = 0; j < kIirChunk; ++j) {
            float l0, l1;
            if (j >= 10) {
                l0 = r0[j - 10];
                l1 = r1[j - 10];
            } else {
                l0 = pick_quantity(px0, py0, j + 22, ym, yo);
                l1 = pick_quantity(px1, py1, j + 22, ym, yo);
            }
            o0[j] = iir_step(k, st0, l0, r0[j]);
            o1[j] = iir_step(k, st1, l1, r1[j]);
        }
        __syncwarp();
        // transposed write-out: 64 row segments of 32 floats starting at column 32t - 4
        const int n = t * kIirChunk - 4 + lane;
        if (n >= 0 && n < nch * kIirChunk) {
            float *dst = ph + (long long)y0 * pitch + n;
            const float *srcp = &sm.to[q][0][lane];
#pragma unroll 4
            for (int rr = 0; rr < rows_here; ++rr) dst[(long long)rr * pitch] = srcp[rr * (kIirChunk + 1)];
        }
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncthreads();  // tile t+1 visible to all; every warp is past its reads of tile t-1
    }
}

// ------------------------------------------------------------------------------------------------
// columns pass + maps + pooling.  grid = (sum over scales of 3 * ceil(w/32), n_candidates), block = 160.
//
// Each warp streams its quantity's row-filtered plane down the image, 20 rows per batch; the next
// batch's 20 loads and this batch's a/b samples are issued before the current batch is computed, so
// DRAM latency is covered by a full batch of arithmetic.  The exchange buffer is double-buffered:
// one barrier per batch.
__global__ void __launch_bounds__(kIirThreads) k_iir_cols(const __grid_constant__ IirArgs a)
{
    __shared__ float ex[2][5][kIirVBatch][kIirVCols];
    __shared__ double sred[5 * 6];

    int s, c, cb;
    decode_cta(a, blockIdx.x, s, c, cb);
    const int cand = blockIdx.y;
    const int w = a.g.w[s], h = a.g.h[s], pitch = a.g.pitch[s];
    const long long poff = a.g.off[s] + (long long)c * a.g.plane[s];
    const float *pa = a.src + poff;
    const float *pb = a.dist + (long long)cand * a.dist_stride + poff;
    const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
    const float *ph = a.hplanes + (long long)cand * a.hplanes_stride + (long long)q * a.q_stride + poff +
                      cb * kIirVCols + lane;
    const IirCoef k = a.k;
    constexpr int kPx = (kIirVBatch * kIirVCols) / kIirThreads;  // 4 map pixels per thread and batch

    IirState st;
#pragma unroll
    for (int i = 0; i < 3; ++i) st.p[i] = st.p2[i] = 0.0f;
    // n = -4..-1: right taps are rows 0..3, nothing emitted
#pragma unroll
    for (int n = -4; n < 0; ++n) {
        const int rr = n + 4;
        (void)iir_step(k, st, 0.0f, rr < h ? __ldg(ph + (long long)rr * pitch) : 0.0f);
    }
    // carry[j] = input row (n0 + j - 6) for j < 10, i.e. the previous batch's R[j + 10].
    float carry[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        const int rr = j - 6;
        carry[j] = (rr >= 0 && rr < h) ? __ldg(ph + (long long)rr * pitch) : 0.0f;
    }
    float R[kIirVBatch], Rn[kIirVBatch];
#pragma unroll
    for (int j = 0; j < kIirVBatch; ++j) {
        const int rr = j + 4;
        R[j] = rr < h ? __ldg(ph + (long long)rr * pitch) : 0.0f;
    }

    double dacc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    int buf = 0;
    for (int n0 = 0; n0 < h; n0 += kIirVBatch, buf ^= 1) {
        // prefetch: next batch's right taps, and this batch's a/b samples for the maps
#pragma unroll
        for (int j = 0; j < kIirVBatch; ++j) {
            const int rr = n0 + kIirVBatch + j + 4;
            Rn[j] = rr < h ? __ldg(ph + (long long)rr * pitch) : 0.0f;
        }
        float av[kPx], bv[kPx];
#pragma unroll
        for (int i = 0; i < kPx; ++i) {
            const int idx = threadIdx.x + i * kIirThreads;
            const int gy = n0 + (idx >> 5), gx = cb * kIirVCols + (idx & 31);
            const bool ok = gy < h && gx < w;
            const long long o = ok ? (long long)gy * pitch + gx : 0;
            av[i] = __ldg(pa + o);
            bv[i] = __ldg(pb + o);
        }
#pragma unroll
        for (int j = 0; j < kIirVBatch; ++j) {
            const float l = (j < 10) ? carry[j] : R[j - 10];
            ex[buf][q][j][lane] = iir_step(k, st, l, R[j]);
        }
#pragma unroll
        for (int j = 0; j < 10; ++j) carry[j] = R[j + 10];
#pragma unroll
        for (int j = 0; j < kIirVBatch; ++j) R[j] = Rn[j];
        __syncthreads();

        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < kPx; ++i) {
            const int idx = threadIdx.x + i * kIirThreads;
            const int row = idx >> 5, col = idx & 31;
            if (n0 + row < h && cb * kIirVCols + col < w)
                error_maps(av[i], bv[i], ex[buf][0][row][col], ex[buf][1][row][col], ex[buf][2][row][col],
                           ex[buf][3][row][col], ex[buf][4][row][col], acc);
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) dacc[j] += (double)acc[j];
        // no second barrier: the next batch writes the other half of ex
    }
    block_reduce6<kIirThreads / 32>(dacc, sred,
                                    a.partials + (long long)cand * a.partials_stride + (long long)blockIdx.x * 6);
}

// ------------------------------------------------------------------------------------------------
// Plain single-plane filters for oavif_ssimu2_debug_blur (filter-only tests; not on the scored path).
__global__ void k_plain_rows(const float *in, float *out, int w, int h, int pitch, int fir, IirCoef k,
                             const float *taps)
{
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= h) return;
    const float *r = in + (long long)y * pitch;
    float *o = out + (long long)y * pitch;
    if (fir) {
        for (int x = 0; x < w; ++x) {
            float acc = 0.f;
            for (int t = 0; t < 9; ++t) {
                const int xx = x + t - 4;
                const float v = (xx >= 0 && xx < w) ? r[xx] : 0.f;
                acc = t ? fmaf(taps[t], v, acc) : taps[0] * v;
            }
            o[x] = acc;
        }
    } else {
        IirState st;
        for (int i = 0; i < 3; ++i) st.p[i] = st.p2[i] = 0.f;
        for (int n = -4; n < w; ++n) {
            const float l = n - 6 >= 0 ? r[n - 6] : 0.f, rv = n + 4 < w ? r[n + 4] : 0.f;
            const float v = iir_step(k, st, l, rv);
            if (n >= 0) o[n] = v;
        }
    }
}

__global__ void k_plain_cols(const float *in, float *out, int w, int h, int pitch, int fir, IirCoef k,
                             const float *taps)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    if (fir) {
        for (int y = 0; y < h; ++y) {
            float acc = 0.f;
            for (int t = 0; t < 9; ++t) {
                const int yy = y + t - 4;
                const float v = (yy >= 0 && yy < h) ? in[(long long)yy * pitch + x] : 0.f;
                acc = t ? fmaf(taps[t], v, acc) : taps[0] * v;
            }
            out[(long long)y * pitch + x] = acc;
        }
    } else {
        IirState st;
        for (int i = 0; i < 3; ++i) st.p[i] = st.p2[i] = 0.f;
        for (int n = -4; n < h; ++n) {
            const float l = n - 6 >= 0 ? in[(long long)(n - 6) * pitch + x] : 0.f;
            const float rv = n + 4 < h ? in[(long long)(n + 4) * pitch + x] : 0.f;
            const float v = iir_step(k, st, l, rv);
            if (n >= 0) out[(long long)n * pitch + x] = v;
        }
    }
}

// ---- host-side launch helpers ------------------------------------------------------------------
inline long long iir_hplane_floats(long long pyr_floats) { return 5 * pyr_floats; }

inline cudaError_t iir_configure()
{
    return cudaFuncSetAttribute(k_iir_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IirRowsSmem));
}

inline cudaError_t launch_iir_blur(const Geom &g, const IirCoef &k, const float *src, const float *dist,
                                   long long pyr_stride, float *hplanes, long long hplanes_stride,
                                   double *partials, long long partials_stride, const int *first_cta_cols,
                                   const int *col_blocks, int n, cudaStream_t st, cudaEvent_t between,
                                   int *launches)
{
    IirArgs a{};
    a.g = g;
    a.k = k;
    a.src = src;
    a.dist = dist;
    a.dist_stride = pyr_stride;
    a.hplanes = hplanes;
    a.hplanes_stride = hplanes_stride;
    a.q_stride = pyr_stride;
    a.partials = partials;
    a.partials_stride = partials_stride;
    int acc = 0;
    for (int s = 0; s < g.n_scales; ++s) {
        a.blocks[s] = (g.h[s] + kIirRows - 1) / kIirRows;
        a.first_cta[s] = acc;
        acc += 3 * a.blocks[s];
    }
    for (int s = g.n_scales; s <= kMaxScales; ++s) a.first_cta[s] = acc;
    k_iir_rows<<<dim3(acc, n), kIirThreads, sizeof(IirRowsSmem), st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (between) cudaEventRecord(between, st);
    for (int s = 0; s <= kMaxScales; ++s) a.first_cta[s] = first_cta_cols[s];
    for (int s = 0; s < kMaxScales; ++s) a.blocks[s] = col_blocks[s];
    k_iir_cols<<<dim3(first_cta_cols[kMaxScales], n), kIirThreads, 0, st>>>(a);
    *launches = 2;
    return cudaGetLastError();
}

inline cudaError_t launch_debug_blur(bool fir, const float *taps, const IirCoef &k, const float *d_in, float *d_tmp,
                                     float *d_out, int w, int h, int pitch, cudaStream_t st)
{
    float *d_taps = nullptr;
    cudaError_t e = cudaMalloc(&d_taps, 9 * sizeof(float));
    if (e != cudaSuccess) return e;
    cudaMemcpyAsync(d_taps, taps, 9 * sizeof(float), cudaMemcpyHostToDevice, st);
    k_plain_rows<<<(h + 63) / 64, 64, 0, st>>>(d_in, d_tmp, w, h, pitch, fir ? 1 : 0, k, d_taps);
    k_plain_cols<<<(w + 63) / 64, 64, 0, st>>>(d_tmp, d_out, w, h, pitch, fir ? 1 : 0, k, d_taps);
    e = cudaGetLastError();
    cudaStreamSynchronize(st);
    cudaFree(d_taps);
    return e;
}

}  // namespace oavif
