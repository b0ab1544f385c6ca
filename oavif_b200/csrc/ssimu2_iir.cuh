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

// Work decomposition (v5): small independent tasks so that the hardware scheduler balances the
// 4 x 148 sub-partitions by itself, and explicit software pipelining inside each task because a
// sub-partition only ever hosts one or two of these warps (latency must be hidden by ILP, not TLP).
//   rows pass   : task = (32 rows, 1 channel, 1 quantity); CTA = 1 warp; lane = row
//   columns pass: task = (32 columns, 1 channel); CTA = 2 warps: warp 0 runs the five recursions,
//                 warp 1 evaluates the maps one 5-row batch behind it; lane = column
constexpr int kIirRows = 32;     // rows per rows-pass task
constexpr int kIirChunk = 32;    // columns per staged tile
constexpr int kIirPitch = 36;    // smem tile pitch in floats: 16-byte rows, conflict-free 128-bit access
constexpr int kIirSlots = 4;     // tile ring: t-1, t, t+1 in use while t+2 lands
constexpr int kIirVCols = 32;    // columns per columns-pass task

struct IirArgs {
    Geom g;
    IirCoef k;
    const float *src;
    const float *dist;
    long long dist_stride;      // floats between candidates' pyramids
    // row-filtered planes, pyramid layout each.  a and a*a only depend on the source: they live in a
    // per-source cache (candidate stride 0) written once by set_source; b, b*b, a*b are per candidate.
    float *hq[5];
    long long hq_cand_stride[5];
    int nq, qlist[4];           // rows pass, single-plane class: the quantities this launch computes
    double *partials;
    long long partials_stride;
    int first_cta[kMaxScales + 1];  // CTA ranges per scale
    int blocks[kMaxScales];     // tasks per channel and scale
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

__device__ __forceinline__ void cp_async_4(float *smem_dst, const float *gmem_src, bool valid)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 4 : 0;  // src-size 0 => the destination is zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem_src), "r"(bytes));
}

template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

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
// rows pass.  Two task classes, launched side by side on two streams because they need different
// amounts of shared memory: NPLANES = 1 for the quantities made from one plane {a, b, a*a, b*b}
// (grid = sum over scales of 3 * 4 * ceil(h/32)), NPLANES = 2 for a*b (3 * ceil(h/32)).  block = 32.
//
// Chunk t emits the 128-byte-aligned outputs n = 32t .. 32t+31: right taps (n+4) come from tiles t and
// t+1, left taps (n-6) from tiles t-1 and t.  Tiles arrive by cp.async (16 bytes per lane) two chunks
// ahead of use; results leave as whole 128-byte lines (streaming stores) through a staging tile that
// reuses the slot of tile t-1, dead once the chunk's samples are in registers.
// The quantity kind (plane, square, product) is uniform per task, so the sample loads branch on it
// without divergence.
// VARIANT is a profiling aid (oavif_ssimu2_debug_time_rows): bit 0 drops the stores, bit 1 the tile
// loads after the first ones.  0 is the product kernel.
template <int NPLANES>
struct IirRowsSmem {
    float tile[NPLANES][kIirSlots][kIirRows][kIirPitch];  // [plane][ring slot][row][column]
};

template <int NPLANES, int VARIANT>
__global__ void __launch_bounds__(32) k_iir_rows(const __grid_constant__ IirArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IirRowsSmem<NPLANES> &sm = *reinterpret_cast<IirRowsSmem<NPLANES> *>(smem_raw);

    int s, c, blk;
    decode_cta(a, blockIdx.x, s, c, blk);
    const int q = NPLANES == 2 ? 4 : a.qlist[blk % a.nq];
    const int rb = NPLANES == 2 ? blk : blk / a.nq;
    const int cand = blockIdx.y;
    const int w = a.g.w[s], h = a.g.h[s], pitch = a.g.pitch[s];
    const int y0 = rb * kIirRows;
    const int rows_here = min(kIirRows, h - y0);
    const long long poff = a.g.off[s] + (long long)c * a.g.plane[s] + (long long)y0 * pitch;
    const float *pa = a.src + poff;
    const float *pb = a.dist + (long long)cand * a.dist_stride + poff;
    const int lane = threadIdx.x;
    float *ph = a.hq[q] + (long long)cand * a.hq_cand_stride[q] + poff;
    const int nch = (w + kIirChunk - 1) / kIirChunk;
    const IirCoef k = a.k;
    // plane 0 of the ring holds a for {a, a*a, a*b} and b for {b, b*b}; plane 1 (a*b only) holds b
    const float *p0 = (NPLANES == 1 && (q & 1)) ? pb : pa;
    const float *p1 = pb;

    // Stage tile t: lane l copies 16 bytes = columns 4*(l&7)..+3 of rows (l>>3) + 4i, i = 0..7.
    // All addresses are base + 32-bit element offsets (one IMAD.WIDE each); rows beyond the image are
    // clamped to a valid address and zero-filled through the copy's source size.
    const int sub_row = lane >> 3, sub_col = (lane & 7) * 4;
    const float *g0 = p0 + sub_col, *g1 = p1 + sub_col;
    unsigned row_off[kIirRows / 4];
    bool row_ok[kIirRows / 4];
#pragma unroll
    for (int i = 0; i < kIirRows / 4; ++i) {
        const int row = sub_row + 4 * i;
        row_ok[i] = row < rows_here;
        row_off[i] = (unsigned)(min(row, rows_here - 1) * pitch);
    }
    auto issue_tile = [&](int t) {
        const int slot = (t + kIirSlots) & (kIirSlots - 1);
        const int gx = t * kIirChunk + sub_col;
        int bytes = 0;
        if (t >= 0 && gx < w) bytes = min(4, w - gx) * 4;
        const unsigned col_off = bytes ? (unsigned)(t * kIirChunk) : 0u;
#pragma unroll
        for (int i = 0; i < kIirRows / 4; ++i) {
            const int nb = row_ok[i] ? bytes : 0;
            cp_async_16(&sm.tile[0][slot][sub_row + 4 * i][sub_col], g0 + (row_off[i] + col_off), nb);
            if (NPLANES == 2)
                cp_async_16(&sm.tile[NPLANES - 1][slot][sub_row + 4 * i][sub_col], g1 + (row_off[i] + col_off), nb);
        }
        cp_async_commit();
    };
    // four samples of the quantity at columns 4*j4..4*j4+3 of a ring slot.  KIND is warp-uniform
    // (one warp per CTA): 0 = the plane itself, 1 = its square, 2 = the product of both planes.
    const int kind = NPLANES == 2 ? 2 : (q < 2 ? 0 : 1);
    auto sample4 = [&](int slot, int j4, float *out) {
        const float4 xv = *reinterpret_cast<const float4 *>(&sm.tile[0][slot][lane][4 * j4]);
        if (NPLANES == 2) {
            const float4 yv = *reinterpret_cast<const float4 *>(&sm.tile[NPLANES - 1][slot][lane][4 * j4]);
            out[0] = xv.x * yv.x; out[1] = xv.y * yv.y; out[2] = xv.z * yv.z; out[3] = xv.w * yv.w;
        } else if (kind == 1) {
            out[0] = xv.x * xv.x; out[1] = xv.y * xv.y; out[2] = xv.z * xv.z; out[3] = xv.w * xv.w;
        } else {
            out[0] = xv.x; out[1] = xv.y; out[2] = xv.z; out[3] = xv.w;
        }
    };

#pragma unroll 1
    for (int t = -1; t <= 1; ++t) issue_tile(t);
    issue_tile(2);
    cp_async_wait<1>();
    __syncwarp();

    IirState st;
#pragma unroll
    for (int i = 0; i < 3; ++i) st.p[i] = st.p2[i] = 0.0f;
    {   // n = -4 .. -1: right taps are columns 0..3 of tile 0, left taps are padding, nothing emitted
        float w4[4];
        sample4(0, 0, w4);
#pragma unroll
        for (int i = 0; i < 4; ++i) (void)iir_step(k, st, 0.0f, w4[i]);
    }

#pragma unroll 1
    for (int t = 0; t < nch; ++t) {
        const int cur = t & (kIirSlots - 1), prev = (t + kIirSlots - 1) & (kIirSlots - 1),
                  next = (t + 1) & (kIirSlots - 1);
        // samples v[i] = quantity at column 32t - 8 + i, i = 0..43 (columns 32t-8 .. 32t+35)
        float v[44];
        sample4(prev, 6, v);
        sample4(prev, 7, v + 4);
#pragma unroll
        for (int j4 = 0; j4 < kIirChunk / 4; ++j4) sample4(cur, j4, v + 8 + 4 * j4);
        sample4(next, 0, v + 40);
        // step j (output column 32t + j): left tap column 32t + j - 6 = v[j+2], right tap 32t + j + 4 = v[j+12]
        float sum[kIirChunk];
#pragma unroll
        for (int j = 0; j < kIirChunk; ++j) sum[j] = v[j + 2] + v[j + 12];
        // tile t-1 is dead from here on (its samples are in registers): its slot is the staging tile
        float4 *o4 = reinterpret_cast<float4 *>(&sm.tile[0][prev][lane][0]);
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
        // transposed write-out of output tile t: whole 128-byte lines, 4 rows per instruction
        if (!(VARIANT & 1) || t == nch - 1) {
            float *dst = ph + sub_col;
            const unsigned col_off = (unsigned)(t * kIirChunk);
#pragma unroll
            for (int i = 0; i < kIirRows / 4; ++i)
                if (row_ok[i])
                    __stcs(reinterpret_cast<float4 *>(dst + (row_off[i] + col_off)),
                           *reinterpret_cast<const float4 *>(&sm.tile[0][prev][sub_row + 4 * i][sub_col]));
        }
        __syncwarp();                 // staging consumed: the slot may be overwritten
        if (!(VARIANT & 2)) issue_tile(t + 3);  // lands in the slot of tile t-1 / the staging tile
        else cp_async_commit();
        cp_async_wait<1>();           // tile t+2 (issued one chunk ago) has landed
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// columns pass + maps + pooling.  grid = (sum over scales of 3 * ceil(w/32), n_candidates), block = 288.
//
// One CTA owns 32 columns of one channel; lane = column in the arithmetic, so the recursions of a warp
// are 32 independent columns.  Warps 0..4 (producers) each stream ONE row-filtered plane down the image
// through a private shared-memory ring.  The ring is fed by 16-byte cp.async: one instruction moves
// four whole 128-byte row segments (8 lanes per row), RCAP-18 rows ahead of use; both taps of the
// recursion are read back from the ring (no register delay line).  Each producer drops its filtered
// values into a double-buffered 8-row batch.  Warps 5..8 (consumers), one batch behind, stream the
// pixel's own XYB samples the same way and evaluate the SSIM / edge-diff maps and the six pooled sums
// for two rows of the batch each (a producer step costs ~17 instructions, a map pixel ~70, so 5 + 4
// warps are balanced).  Nine warps per task give a sub-partition enough independent work to hide
// latencies when a single 4K pair is all the GPU has.  One block barrier per 8 rows.
constexpr int kIirVThreads = 288;   // 5 producer warps + 4 consumer warps

template <int RCAP, int B>
struct IirColsSmem {
    float ring[5][RCAP][kIirVCols];            // producer input rows, row r at [r & (RCAP-1)]
    float ab[2][32][kIirVCols];                // consumer rows of the two XYB planes
    float ex[2][5][B][kIirVCols];              // filtered values, double-buffered
    double red[4][6];
};

// B rows per exchange batch (16 with the deep ring, 8 with the shallow one): the per-batch overhead
// (barrier, copy issue, address set-up) is paid once per B rows by each of the nine warps.
template <int RCAP, int B>
__global__ void __launch_bounds__(kIirVThreads) k_iir_cols(const __grid_constant__ IirArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IirColsSmem<RCAP, B> &sm = *reinterpret_cast<IirColsSmem<RCAP, B> *>(smem_raw);
    constexpr int D = ((RCAP - B - 10) / B) * B;   // rows of look-ahead: D + B + 10 <= RCAP, D % B == 0
    constexpr int DA = 16;                         // consumer look-ahead (ring of 32 rows)
    constexpr int CR = B / 4;                      // rows per consumer warp and batch
    static_assert(D >= B && DA % B == 0 && DA + B <= 32 && B % 8 == 0 && RCAP % B == 0, "ring geometry");

    int s, c, cb;
    decode_cta(a, blockIdx.x, s, c, cb);
    const int cand = blockIdx.y;
    const int w = a.g.w[s], h = a.g.h[s], pitch = a.g.pitch[s];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gx = cb * kIirVCols + lane;
    const bool col_ok = gx < w;
    // copy role of a lane: row (lane >> 3) of a 4-row group, columns 4*(lane & 7)..+3 of the 32
    const int crow = lane >> 3, ccol = (lane & 7) * 4;
    const long long poff = a.g.off[s] + (long long)c * a.g.plane[s] + cb * kIirVCols;
    const int nbatch = (h + B - 1) / B;

    if (warp < 5) {
        // ---------------- producer: the column recursion of quantity `warp` ----------------
        const int q = warp;
        const float *ph = a.hq[q] + (long long)cand * a.hq_cand_stride[q] + poff + ccol;
        const IirCoef k = a.k;
        float *ring = &sm.ring[q][0][0];
        auto issue_rows4 = [&](int r0) {   // rows r0..r0+3 (zeros beyond h): one 16-byte copy per lane
            const int rr = r0 + crow;
            cp_async_16(ring + (rr & (RCAP - 1)) * kIirVCols + ccol, ph + (unsigned)(min(rr, h - 1) * pitch),
                        rr < h ? 16 : 0);
        };
        // rows -6..-1 are padding: their ring slots hold zeros until real rows wrap around to them
#pragma unroll
        for (int j = 1; j <= 6; ++j) ring[(RCAP - j) * kIirVCols + lane] = 0.0f;
        for (int r0 = 0; r0 < 4 + D; r0 += 4) issue_rows4(r0);   // everything before the first batch's request
        cp_async_commit();
        IirState st;
#pragma unroll
        for (int i = 0; i < 3; ++i) st.p[i] = st.p2[i] = 0.0f;
        cp_async_wait<0>();
        __syncwarp();
        const float *col = ring + lane;
        // n = -4..-1: right taps are rows 0..3, left taps are padding, nothing emitted
#pragma unroll
        for (int n = -4; n < 0; ++n) (void)iir_step(k, st, 0.0f, col[(n + 4) * kIirVCols]);

#pragma unroll 1
        for (int b = 0; b < nbatch; ++b) {
            const int n0 = b * B;
#pragma unroll
            for (int j = 0; j < B; j += 4) issue_rows4(n0 + 4 + D + j);
            cp_async_commit();
            cp_async_wait<D / B>();                    // rows up to n0 + B + 3 have landed (this lane's copies)
            __syncwarp();                              // ... and every other lane's
            // n0 is a multiple of B (8 or 16) and so is RCAP: the left taps (rows n0-6+j) can only wrap at
            // j = 6, the right taps (rows n0+4+j) only at j = B-4 -> two bases each, static offsets otherwise
            float sum[B];
            const float *l0 = col + ((n0 - 6) & (RCAP - 1)) * kIirVCols, *l1 = col + (n0 & (RCAP - 1)) * kIirVCols;
            const float *r0 = col + ((n0 + 4) & (RCAP - 1)) * kIirVCols, *r1 = col + ((n0 + B) & (RCAP - 1)) * kIirVCols;
#pragma unroll
            for (int j = 0; j < B; ++j)
                sum[j] = (j < 6 ? l0[j * kIirVCols] : l1[(j - 6) * kIirVCols]) +
                         (j < B - 4 ? r0[j * kIirVCols] : r1[(j - (B - 4)) * kIirVCols]);
            float *ex = &sm.ex[b & 1][q][0][lane];
            IirPipe P;
            pipe_begin(k, P, st, sum[0]);
#pragma unroll
            for (int j = 0; j < B; ++j)
                ex[j * kIirVCols] = (j + 1 < B) ? pipe_step(k, P, sum[j + 1]) : pipe_end(k, P, st);
            __syncthreads();  // batch b published; the consumers are done with the other buffer
        }
        __syncthreads();      // consumers' last batch
        __syncthreads();      // final reduction
    } else {
        // ---------------- consumers: maps + pooling for CR rows of each batch ----------------
        const int cw = warp - 5;                       // 0..3
        // the XYB samples arrive in 4-row groups: with B = 16 each consumer stages its own four rows,
        // with B = 8 the even warp of a pair stages the four rows the pair shares
        const int grp = (CR == 4) ? cw : (cw >> 1);
        const bool loader = (CR == 4) || ((cw & 1) == 0);
        const float *pa = a.src + poff + ccol;
        const float *pb = a.dist + (long long)cand * a.dist_stride + poff + ccol;
        double dacc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        auto issue_ab4 = [&](int r0) {   // rows r0..r0+3 of both planes
            const int rr = r0 + crow;
            const unsigned o = (unsigned)(min(rr, h - 1) * pitch);
            const int nb = rr < h ? 16 : 0;
            cp_async_16(&sm.ab[0][rr & 31][ccol], pa + o, nb);
            cp_async_16(&sm.ab[1][rr & 31][ccol], pb + o, nb);
        };
        if (loader)
            for (int r0 = 4 * grp; r0 < DA; r0 += B) issue_ab4(r0);
        cp_async_commit();
        cp_async_wait<0>();   // the in-loop wait only covers groups committed inside the loop
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int b = 0; b < nbatch; ++b) {
            if (loader) issue_ab4(b * B + 4 * grp + DA);
            cp_async_commit();
            cp_async_wait<DA / B>();                   // the rows of batch b staged by this warp have landed
            __syncthreads();                           // batch b is in ex[b & 1]; staged samples are visible
            const int j0 = CR * cw, n0 = b * B + j0;   // this warp's rows
            const float *ex = &sm.ex[b & 1][0][j0][lane];
#pragma unroll
            for (int j = 0; j < CR; ++j) {
                const int n = n0 + j;
                if (col_ok && n < h)
                    error_maps(sm.ab[0][n & 31][lane], sm.ab[1][n & 31][lane], ex[(0 * B + j) * kIirVCols],
                               ex[(1 * B + j) * kIirVCols], ex[(2 * B + j) * kIirVCols],
                               ex[(3 * B + j) * kIirVCols], ex[(4 * B + j) * kIirVCols], acc);
            }
            if ((b & 3) == 3) {   // binary32 over at most 16 pixels, binary64 from there on
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    dacc[j] += (double)acc[j];
                    acc[j] = 0.f;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) dacc[j] += (double)acc[j];
        __syncthreads();
        // fixed shuffle tree over the 32 columns, then the four consumers in fixed order
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            double x = dacc[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            if (lane == 0) sm.red[cw][j] = x;
        }
        __syncthreads();
        if (cw == 0 && lane < 6)
            a.partials[(long long)cand * a.partials_stride + (long long)blockIdx.x * 6 + lane] =
                ((sm.red[0][lane] + sm.red[1][lane]) + sm.red[2][lane]) + sm.red[3][lane];
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
inline long long iir_hplane_floats(long long pyr_floats) { return 3 * pyr_floats; }  // per candidate: b, b*b, a*b

inline cudaError_t iir_configure()
{
    cudaError_t e = cudaFuncSetAttribute(k_iir_cols<64, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(IirColsSmem<64, 16>));
    return e;
}

struct IirStreams {
    cudaStream_t side;       // source-side rows tasks (a, a*a), behind set_source
    cudaStream_t side2;      // the a*b rows tasks of a scoring call, next to {b, b*b} on the main stream
    cudaEvent_t fork, join;  // main -> side / side2, side2 -> main
    cudaEvent_t src_done;    // the source's cached row-filtered planes are complete (recorded on `side`)
};

struct IirBuffers {
    float *src_hplanes;      // [2][pyramid]: rows pass of a and a*a, cached per source
    float *cand_hplanes;     // [candidate][3][pyramid]: rows pass of b, b*b, a*b
    long long pyr_stride;    // floats per pyramid (capacity)
};

inline void iir_fill_common(IirArgs &a, const Geom &g, const IirCoef &k, const float *src, const float *dist,
                            long long dist_stride, const IirBuffers &B)
{
    a.g = g;
    a.k = k;
    a.src = src;
    a.dist = dist;
    a.dist_stride = dist_stride;
    const long long P = B.pyr_stride;
    a.hq[0] = B.src_hplanes;               a.hq_cand_stride[0] = 0;          // a
    a.hq[2] = B.src_hplanes + P;           a.hq_cand_stride[2] = 0;          // a*a
    a.hq[1] = B.cand_hplanes;              a.hq_cand_stride[1] = 3 * P;      // b
    a.hq[3] = B.cand_hplanes + P;          a.hq_cand_stride[3] = 3 * P;      // b*b
    a.hq[4] = B.cand_hplanes + 2 * P;      a.hq_cand_stride[4] = 3 * P;      // a*b
}

inline int iir_rows_grid(IirArgs &a, const Geom &g, int per_rowblock)
{
    int n = 0;
    for (int s = 0; s < g.n_scales; ++s) {
        const int nrb = (g.h[s] + kIirRows - 1) / kIirRows;
        a.blocks[s] = per_rowblock * nrb;
        a.first_cta[s] = n;
        n += 3 * per_rowblock * nrb;
    }
    for (int s = g.n_scales; s <= kMaxScales; ++s) a.first_cta[s] = n;
    return n;
}

template <int NPLANES>
inline void launch_rows_kernel(const IirArgs &a, int n_cta, int ncand, cudaStream_t st, int variant)
{
    const dim3 grid(n_cta, ncand);
    const size_t sm = sizeof(IirRowsSmem<NPLANES>);
    switch (variant) {
    case 0: k_iir_rows<NPLANES, 0><<<grid, 32, sm, st>>>(a); break;
    case 1: k_iir_rows<NPLANES, 1><<<grid, 32, sm, st>>>(a); break;
    case 2: k_iir_rows<NPLANES, 2><<<grid, 32, sm, st>>>(a); break;
    default: k_iir_rows<NPLANES, 3><<<grid, 32, sm, st>>>(a); break;
    }
}

// Source side, once per set_source: rows pass of a and a*a into the per-source cache.  Runs on the side
// stream behind the source pyramid, so it overlaps the candidate's upload and pyramid.
inline cudaError_t launch_iir_source_rows(const Geom &g, const IirCoef &k, const float *src_pyr, const IirBuffers &B,
                                          cudaStream_t st, const IirStreams &ss, int *launches)
{
    IirArgs a{};
    iir_fill_common(a, g, k, src_pyr, src_pyr, 0, B);
    a.nq = 2;
    a.qlist[0] = 0;
    a.qlist[1] = 2;
    const int n = iir_rows_grid(a, g, 2);
    cudaEventRecord(ss.fork, st);                 // after the source pyramid
    cudaStreamWaitEvent(ss.side, ss.fork, 0);
    launch_rows_kernel<1>(a, n, 1, ss.side, 0);
    const cudaError_t e = cudaGetLastError();
    cudaEventRecord(ss.src_done, ss.side);
    *launches = 1;
    return e;
}

// Candidate side: rows pass of b, b*b (main stream) and a*b (side stream), then the columns pass with
// the maps and the pooling.  `between` is recorded between the two passes.
inline cudaError_t launch_iir_blur(const Geom &g, const IirCoef &k, const float *src, const float *dist,
                                   long long pyr_stride, const IirBuffers &B, double *partials,
                                   long long partials_stride, const int *first_cta_cols, const int *col_blocks, int n,
                                   cudaStream_t st, const IirStreams &ss, cudaEvent_t between, int *launches,
                                   int variant = 0, bool rows_only = false)
{
    IirArgs a{};
    iir_fill_common(a, g, k, src, dist, pyr_stride, B);
    a.partials = partials;
    a.partials_stride = partials_stride;
    IirArgs a1 = a, a2 = a;   // {b, b*b} | a*b
    a1.nq = 2;
    a1.qlist[0] = 1;
    a1.qlist[1] = 3;
    const int n1 = iir_rows_grid(a1, g, 2), n2 = iir_rows_grid(a2, g, 1);
    cudaEventRecord(ss.fork, st);                 // after the candidate pyramid
    cudaStreamWaitEvent(ss.side2, ss.fork, 0);    // a*b must not queue behind the source rows: own stream
    launch_rows_kernel<1>(a1, n1, n, st, variant);
    launch_rows_kernel<2>(a2, n2, n, ss.side2, variant);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    cudaEventRecord(ss.join, ss.side2);
    cudaStreamWaitEvent(st, ss.join, 0);          // a*b done
    cudaStreamWaitEvent(st, ss.src_done, 0);      // the source's cached rows done
    if (between) cudaEventRecord(between, st);
    *launches = 2;
    if (rows_only) return cudaSuccess;
    for (int s = 0; s <= kMaxScales; ++s) a.first_cta[s] = first_cta_cols[s];
    for (int s = 0; s < kMaxScales; ++s) a.blocks[s] = col_blocks[s];
    // ring depth: the deep ring when every scale-0 task can still be resident, else the shallow one
    const int scale0_tasks = 3 * col_blocks[0] * n;
    if (scale0_tasks <= 148 * 3)
        k_iir_cols<64, 16><<<dim3(first_cta_cols[kMaxScales], n), kIirVThreads, sizeof(IirColsSmem<64, 16>), st>>>(a);
    else
        k_iir_cols<32, 8><<<dim3(first_cta_cols[kMaxScales], n), kIirVThreads, sizeof(IirColsSmem<32, 8>), st>>>(a);
    *launches = 3;
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
