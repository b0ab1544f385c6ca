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

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Work decomposition (v4): small independent tasks so that the hardware scheduler balances the
// 4 x 148 sub-partitions by itself, and explicit software pipelining inside each task because a
// sub-partition only ever hosts one or two of these warps (latency must be hidden by ILP, not TLP).
//   rows pass   : task = (32 rows, 1 channel, 1 quantity); CTA = 1 warp; lane = row
//   columns pass: task = (32 columns, 1 channel); CTA = 2 warps: warp 0 runs the five recursions,
//                 warp 1 evaluates the maps one 5-row batch behind it; lane = column
constexpr int kIirRows = 32;     // rows per rows-pass task
constexpr int kIirChunk = 32;    // columns per staged tile
constexpr int kIirPitch = 36;    // smem tile pitch in floats: 16-byte rows, conflict-free 128-bit access
constexpr int kIirSlots = 3;     // tile ring: t-1, t in use while t+1 lands (L2-prefetched 3 tiles ahead)
constexpr int kIirVCols = 32;    // columns per columns-pass task
constexpr int kIirVBatch = 5;    // rows per exchange batch in the columns pass

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
    int blocks[kMaxScales];     // tasks per channel: 5 * ceil(h/32) (rows pass) or ceil(w/32) (columns pass)
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

__device__ __forceinline__ void cp_async_16(float *smem_dst, const float *gmem_src, int src_bytes)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    // bytes beyond src_bytes (0, 4, 8, 12 or 16) are zero-filled: the filter's zero padding
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }

// The recursion, software-pipelined by hand.  iir_step() above is the definition; this form performs
// the SAME operations on the SAME values (sum*n2, minus prev2, fma(-d1, prev, .), (o1+o3)+o5) but
// issues "sum*n2 - prev2" of step n+1 — which only needs the output of step n-1 — next to the fused
// multiply-add of step n, so the dependent chain per step is one FFMA instead of five operations.
struct IirPipe {
    float p1[3];  // outputs of the previous step, per oscillator
    float u[3];   // sum*n2 - prev2 of the CURRENT step, already evaluated
};

__device__ __forceinline__ void pipe_begin(const IirCoef &k, IirPipe &P, const IirState &st, float sum0)
{
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        P.p1[i] = st.p[i];
        P.u[i] = sum0 * k.n2[i] - st.p2[i];
    }
}

// finishes the current step (returns its output) and pre-evaluates the next one from `sum_next`
__device__ __forceinline__ float pipe_step(const IirCoef &k, IirPipe &P, float sum_next)
{
    float nw[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) nw[i] = fmaf(-k.d1[i], P.p1[i], P.u[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        P.u[i] = sum_next * k.n2[i] - P.p1[i];
        P.p1[i] = nw[i];
    }
    return (nw[0] + nw[1]) + nw[2];
}

// last step of a run: also hands (prev, prev2) back as an IirState
__device__ __forceinline__ float pipe_end(const IirCoef &k, IirPipe &P, IirState &st)
{
    float nw[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        nw[i] = fmaf(-k.d1[i], P.p1[i], P.u[i]);
        st.p2[i] = P.p1[i];
        st.p[i] = nw[i];
    }
    return (nw[0] + nw[1]) + nw[2];
}

// ------------------------------------------------------------------------------------------------
// rows pass.  grid = (sum over scales of 3 * 5 * ceil(h/32), n_candidates), block = 32, dynamic smem.
//
// Chunk t emits outputs n = 32t-4 .. 32t+27, so its right taps (n+4) are exactly tile t and its left
// taps (n-6) fall in tiles t-1 and t.  Tiles arrive by cp.async (16 bytes per lane) one chunk ahead of
// use and are pulled into L2 four chunks ahead; results leave as 16-byte row segments through a
// staging tile that reuses the slot of tile t-1 (dead once the chunk's samples are in registers).
// Chunk 0's first four outputs are the recursion's warm-up steps n = -4..-1 and are dropped.
// Quantity q of {a, b, a*a, b*b, a*b} = x * (y*m + o) with warp-uniform tile pointers for x, y and
// (m, o) = (0, 1) for the two plain planes, (1, 0) for the products: y*0+1 and x*1 are exact, so every
// quantity is bit-identical to the direct expression, with one code path.
struct IirRowsSmem {
    float tile[2][kIirSlots][kIirRows][kIirPitch];  // [plane a|b][ring slot][row][column]
};

__global__ void __launch_bounds__(32) k_iir_rows(const __grid_constant__ IirArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IirRowsSmem &sm = *reinterpret_cast<IirRowsSmem *>(smem_raw);

    int s, c, blk;
    decode_cta(a, blockIdx.x, s, c, blk);
    const int q = blk % 5, rb = blk / 5;
    const int cand = blockIdx.y;
    const int w = a.g.w[s], h = a.g.h[s], pitch = a.g.pitch[s];
    const int y0 = rb * kIirRows;
    const int rows_here = min(kIirRows, h - y0);
    const long long poff = a.g.off[s] + (long long)c * a.g.plane[s] + (long long)y0 * pitch;
    const float *pa = a.src + poff;
    const float *pb = a.dist + (long long)cand * a.dist_stride + poff;
    const int lane = threadIdx.x;
    float *ph = a.hplanes + (long long)cand * a.hplanes_stride + (long long)q * a.q_stride + poff;
    const int nch = (w + kIirChunk - 1) / kIirChunk;          // input tiles
    const int nout = (w + 4 + kIirChunk - 1) / kIirChunk;     // output chunks
    const IirCoef k = a.k;
    const bool need_a = (q != 1) && (q != 3), need_b = (q != 0) && (q != 2);
    const int xp = (q == 1 || q == 3) ? 1 : 0;    // x: b for {b, b*b}, a otherwise
    const int yp = (q == 0 || q == 2) ? 0 : 1;    // y: a for a*a, b for {b*b, a*b}; for q < 2 any LOADED plane (y*0+1)
    const float ym = q < 2 ? 0.0f : 1.0f, yo = q < 2 ? 1.0f : 0.0f;

    // Stage tile t: lane l copies 16 bytes = columns 4*(l&7)..+3 of rows (l>>3) + 4i, i = 0..7.
    const int sub_row = lane >> 3, sub_col = (lane & 7) * 4;
    auto issue_tile = [&](int t) {
        const int slot = (t + kIirSlots) % kIirSlots;
        const int gx = t * kIirChunk + sub_col;
        int bytes = 0;
        if (t >= 0 && gx < w) bytes = min(4, w - gx) * 4;
#pragma unroll
        for (int i = 0; i < kIirRows / 4; ++i) {
            const int row = sub_row + 4 * i;
            const int nb = row < rows_here ? bytes : 0;
            const long long o = nb ? (long long)row * pitch + gx : 0;
            if (need_a) cp_async_16(&sm.tile[0][slot][row][sub_col], pa + o, nb);
            if (need_b) cp_async_16(&sm.tile[1][slot][row][sub_col], pb + o, nb);
        }
        cp_async_commit();
    };
    // Pull tile t into L2: lane = row, one 128-byte line per plane.
    auto prefetch_tile = [&](int t) {
        if (t < nch && lane < rows_here) {
            const long long o = (long long)lane * pitch + t * kIirChunk;
            if (need_a) prefetch_l2(pa + o);
            if (need_b) prefetch_l2(pb + o);
        }
    };

#pragma unroll 1
    for (int t = 2; t <= 4; ++t) prefetch_tile(t);
#pragma unroll 1
    for (int t = -1; t <= 1; ++t) issue_tile(t);
    cp_async_wait_all();
    __syncwarp();

    IirState st;
#pragma unroll
    for (int i = 0; i < 3; ++i) st.p[i] = st.p2[i] = 0.0f;

#pragma unroll 1
    for (int t = 0; t < nout; ++t) {
        const int cur = t % kIirSlots, prev = (t + kIirSlots - 1) % kIirSlots;
        const float4 *x4 = reinterpret_cast<const float4 *>(&sm.tile[xp][cur][lane][0]);
        const float4 *y4 = reinterpret_cast<const float4 *>(&sm.tile[yp][cur][lane][0]);
        const float4 *px4 = reinterpret_cast<const float4 *>(&sm.tile[xp][prev][lane][0]);
        const float4 *py4 = reinterpret_cast<const float4 *>(&sm.tile[yp][prev][lane][0]);
        float r[kIirChunk], lp[12];
#pragma unroll
        for (int j4 = 0; j4 < 3; ++j4) {  // previous tile's columns 20..31 (22..31 are the left taps)
            const float4 xv = px4[5 + j4], yv = py4[5 + j4];
            lp[4 * j4 + 0] = xv.x * fmaf(yv.x, ym, yo);
            lp[4 * j4 + 1] = xv.y * fmaf(yv.y, ym, yo);
            lp[4 * j4 + 2] = xv.z * fmaf(yv.z, ym, yo);
            lp[4 * j4 + 3] = xv.w * fmaf(yv.w, ym, yo);
        }
#pragma unroll
        for (int j4 = 0; j4 < kIirChunk / 4; ++j4) {
            const float4 xv = x4[j4], yv = y4[j4];
            r[4 * j4 + 0] = xv.x * fmaf(yv.x, ym, yo);
            r[4 * j4 + 1] = xv.y * fmaf(yv.y, ym, yo);
            r[4 * j4 + 2] = xv.z * fmaf(yv.z, ym, yo);
            r[4 * j4 + 3] = xv.w * fmaf(yv.w, ym, yo);
        }
        // sums l + r of the 32 steps (left tap of step j: column j+22 of the previous tile, or r[j-10])
        float sum[kIirChunk];
#pragma unroll
        for (int j = 0; j < kIirChunk; ++j) sum[j] = ((j >= 10) ? r[j - 10] : lp[j + 2]) + r[j];
        // tile t-1 is dead from here on (its samples are in registers): its slot is the staging tile
        float4 *o4 = reinterpret_cast<float4 *>(&sm.tile[xp][prev][lane][0]);
        IirPipe P;
        pipe_begin(k, P, st, sum[0]);
#pragma unroll
        for (int j4 = 0; j4 < kIirChunk / 4; ++j4) {
            float o[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int j = 4 * j4 + jj;
                o[jj] = (j + 1 < kIirChunk) ? pipe_step(k, P, sum[j + 1]) : pipe_end(k, P, st);
            }
            o4[j4] = make_float4(o[0], o[1], o[2], o[3]);
        }
        __syncwarp();
        // transposed write-out: 16 bytes per lane, 4 rows per instruction, starting at column 32t - 4
        {
            const int n = t * kIirChunk - 4 + sub_col;
            if (n >= 0 && n < nch * kIirChunk) {
#pragma unroll
                for (int i = 0; i < kIirRows / 4; ++i) {
                    const int row = sub_row + 4 * i;
                    if (row < rows_here)
                        *reinterpret_cast<float4 *>(ph + (long long)row * pitch + n) =
                            *reinterpret_cast<const float4 *>(&sm.tile[xp][prev][row][sub_col]);
                }
            }
        }
        __syncwarp();            // staging consumed: the slot may be overwritten
        issue_tile(t + 2);       // lands in the slot tile t-1 / the staging tile just vacated
        prefetch_tile(t + 5);
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncwarp();            // tile t+1 (issued one chunk ago) is visible to every lane
    }
}

// ------------------------------------------------------------------------------------------------
// columns pass + maps + pooling.  grid = (sum over scales of 3 * ceil(w/32), n_candidates), block = 64.
//
// Warp 0 streams all five row-filtered planes of its 32 columns down the image: 15 recursions per lane,
// left taps in a 10-deep circular register delay line per quantity, loads issued one 5-row batch ahead.
// It drops the five filtered values of each pixel into a double-buffered shared-memory batch; warp 1,
// one batch behind, loads the pixel's own XYB samples and evaluates the SSIM / edge-diff maps and the
// six pooled sums.  One block barrier per 5 rows; the two warps sit on different sub-partitions.
struct ColsBatch {
    float r[5][kIirVBatch];  // [quantity][row in batch]: right taps = row-filtered input rows
};

__device__ __forceinline__ void cols_load(ColsBatch &B, const float *ph, long long q_stride, int pitch, int h,
                                          int n_first)
{
    // right taps of outputs n_first..n_first+4 are input rows n_first+4..n_first+8
#pragma unroll
    for (int j = 0; j < kIirVBatch; ++j) {
        const int rr = n_first + j + 4;
        const bool ok = rr < h;
        const long long o = (long long)rr * pitch;
        // predicated loads (no select on the loaded value: nothing may depend on it until it is used)
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            B.r[q][j] = 0.0f;
            if (ok) B.r[q][j] = __ldg(ph + q * q_stride + o);
        }
    }
}

// five outputs n_first .. n_first+4 of every quantity into ex[q][j][lane]; PHASE (0 or 5) is
// n_first mod 10: it makes every delay-line index a compile-time constant.
template <int PHASE>
__device__ __forceinline__ void cols_compute(const ColsBatch &B, const IirCoef &k, IirState st[5], float d[5][10],
                                             float (*ex)[kIirVBatch][kIirVCols], int lane)
{
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        float sum[kIirVBatch];
#pragma unroll
        for (int j = 0; j < kIirVBatch; ++j) {
            const int slot = (PHASE + j + 4) % 10;  // row (n+4) mod 10 == row (n-6) mod 10
            sum[j] = d[q][slot] + B.r[q][j];
            d[q][slot] = B.r[q][j];
        }
        IirPipe P;
        pipe_begin(k, P, st[q], sum[0]);
#pragma unroll
        for (int j = 0; j < kIirVBatch; ++j)
            ex[q][j][lane] = (j + 1 < kIirVBatch) ? pipe_step(k, P, sum[j + 1]) : pipe_end(k, P, st[q]);
    }
}

__global__ void __launch_bounds__(64) k_iir_cols(const __grid_constant__ IirArgs a)
{
    __shared__ float ex[2][5][kIirVBatch][kIirVCols];

    int s, c, cb;
    decode_cta(a, blockIdx.x, s, c, cb);
    const int cand = blockIdx.y;
    const int w = a.g.w[s], h = a.g.h[s], pitch = a.g.pitch[s];
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    const int gx = cb * kIirVCols + lane;
    const bool col_ok = gx < w;
    const long long poff = a.g.off[s] + (long long)c * a.g.plane[s] + gx;
    const int nbatch = (h + kIirVBatch - 1) / kIirVBatch;  // producer runs batches 0..nbatch-1, consumer one behind

    if (role == 0) {
        // ---------------- producer: the five column recursions ----------------
        const float *ph = a.hplanes + (long long)cand * a.hplanes_stride + poff;
        const long long qs = a.q_stride;
        const IirCoef k = a.k;
        IirState st[5];
        float d[5][10];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
#pragma unroll
            for (int i = 0; i < 3; ++i) st[q].p[i] = st[q].p2[i] = 0.0f;
#pragma unroll
            for (int i = 0; i < 10; ++i) d[q][i] = 0.0f;
        }
        // n = -4..-1: right taps are rows 0..3 (kept at delay slots 0..3), left taps are padding
#pragma unroll
        for (int n = -4; n < 0; ++n) {
            const int rr = n + 4;
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                float v = 0.0f;
                if (rr < h) v = __ldg(ph + q * qs + (long long)rr * pitch);
                (void)iir_step(k, st[q], 0.0f, v);
                d[q][rr] = v;
            }
        }
        ColsBatch B0, B1;
        cols_load(B0, ph, qs, pitch, h, 0);
#pragma unroll 1
        for (int b = 0; b < nbatch; b += 2) {
            cols_load(B1, ph, qs, pitch, h, (b + 1) * kIirVBatch);
            cols_compute<0>(B0, k, st, d, ex[0], lane);
            __syncthreads();  // batch b published; consumer finished batch b-1 (other buffer)
            cols_load(B0, ph, qs, pitch, h, (b + 2) * kIirVBatch);
            cols_compute<5>(B1, k, st, d, ex[1], lane);
            __syncthreads();  // batch b+1 published
        }
        __syncthreads();      // consumer's last batch
    } else {
        // ---------------- consumer: maps + pooling, one batch behind ----------------
        const float *pa = a.src + poff;
        const float *pb = a.dist + (long long)cand * a.dist_stride + poff;
        double dacc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        float av[kIirVBatch], bv[kIirVBatch], an[kIirVBatch], bn[kIirVBatch];
        auto load_ab = [&](float *A, float *Bv, int n_first) {
#pragma unroll
            for (int j = 0; j < kIirVBatch; ++j) {
                const int n = n_first + j;
                A[j] = 0.0f;
                Bv[j] = 0.0f;
                if (n < h) {
                    A[j] = __ldg(pa + (long long)n * pitch);
                    Bv[j] = __ldg(pb + (long long)n * pitch);
                }
            }
        };
        auto maps = [&](const float *A, const float *Bv, int buf, int n_first) {
            float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < kIirVBatch; ++j)
                if (col_ok && n_first + j < h)
                    error_maps(A[j], Bv[j], ex[buf][0][j][lane], ex[buf][1][j][lane], ex[buf][2][j][lane],
                               ex[buf][3][j][lane], ex[buf][4][j][lane], acc);
#pragma unroll
            for (int j = 0; j < 6; ++j) dacc[j] += (double)acc[j];
        };
        load_ab(av, bv, 0);
#pragma unroll 1
        for (int b = 0; b < nbatch; b += 2) {
            load_ab(an, bn, (b + 1) * kIirVBatch);
            __syncthreads();                                   // batch b is in ex[0]
            maps(av, bv, 0, b * kIirVBatch);
            load_ab(av, bv, (b + 2) * kIirVBatch);
            __syncthreads();                                   // batch b+1 is in ex[1]
            maps(an, bn, 1, (b + 1) * kIirVBatch);
        }
        __syncthreads();
        // fixed shuffle tree over the 32 columns, lane 0 writes the task's six sums
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            double x = dacc[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            dacc[j] = x;
        }
        if (lane == 0) {
            double *out = a.partials + (long long)cand * a.partials_stride + (long long)blockIdx.x * 6;
#pragma unroll
            for (int j = 0; j < 6; ++j) out[j] = dacc[j];
        }
    }
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
        a.blocks[s] = 5 * ((g.h[s] + kIirRows - 1) / kIirRows);
        a.first_cta[s] = acc;
        acc += 3 * a.blocks[s];
    }
    for (int s = g.n_scales; s <= kMaxScales; ++s) a.first_cta[s] = acc;
    k_iir_rows<<<dim3(acc, n), 32, sizeof(IirRowsSmem), st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (between) cudaEventRecord(between, st);
    for (int s = 0; s <= kMaxScales; ++s) a.first_cta[s] = first_cta_cols[s];
    for (int s = 0; s < kMaxScales; ++s) a.blocks[s] = col_blocks[s];
    k_iir_cols<<<dim3(first_cta_cols[kMaxScales], n), 64, 0, st>>>(a);
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
