// ssimu2_iir.cuh — K4+K5, RECURSIVE form: the sigma = 1.5 recursive Gaussian of SSIMULACRA2 v2.1
// (Charalampidis 2016, three undamped oscillators k in {1,3,5}, radius N = 5; SURVEY.md Appendix A
// §4) evaluated exactly as published — every row and every column is ONE serial binary32 chain
// from n = -N+1, because the recursion's round-off is not negligible at the metric's C2 = 9e-4
// scale and is therefore part of the published result.
//
// Both passes run the recursion on PACKED pairs (FFMA2 / FADD2 / FMUL2: two IEEE binary32 operations per issued
// instruction, same bits as scalar) and move their tiles with TMA (cp.async.bulk.tensor + mbarrier, ssimu2_tma.cuh):
//
//   k_iir_rows_tma : rows (default).  One CTA = 32 rows of one channel; lane = row.  The pair is two QUANTITIES of
//                 the same pixel: (a, a*a) on the source side, (b, b*b) on the candidate side, with a*b in a scalar
//                 recursion warp.  One elected lane requests 44 x 32 boxes into a four-slot ring (out-of-bounds =
//                 the filter's zero padding); each recursion warp waits on the slot's mbarrier, runs its chain, stages
//                 the finished 32 x 32 chunk in swizzled shared memory and sends it off with a TMA store of its own.
//                 The pairs are written to HBM interleaved (one float2 per pixel), exactly as they are computed.
//   k_iir_rows  : the round-1 form (loader and storer warps, 16-byte cp.async ring, one block barrier per chunk):
//                 OAVIF_SSIMU2_TILES_CP_ASYNC.
//   k_iir_cols  : columns + error maps + pooling.  One CTA = 32 columns of one channel; lane = column.
//                 Two producer warps run the packed recursions of (a, a*a) and (b, b*b) straight from the
//                 interleaved planes, one the scalar recursion of a*b; one lane of the loader warp feeds their
//                 shared-memory rings and the consumers' XYB rows by TMA (TMA = false: the whole warp, by cp.async);
//                 four consumer warps evaluate the SSIM / edge-diff maps on packed pairs of ROWS, one batch
//                 behind — blurred planes are never written to HBM.
//
// HBM traffic per scale pixel and channel: rows pass reads 8 B, writes 20 B; columns pass reads
// 20 B + 8 B.  No tensor cores (nothing here is a contraction).
#pragma once

#include <type_traits>

#include "ssimu2_common.cuh"
#include "ssimu2_tma.cuh"

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

// The same step on a packed pair of independent chains.
struct IirCoef2 {
    f32x2 n2[3], nd1[3];   // (n2, n2) and (-d1, -d1)
    Unit2 u;               // +1 / -1 the compiler cannot see (see ssimu2_common.cuh)
};

struct IirState2 {
    f32x2 p[3], p2[3];
};

__device__ __forceinline__ IirCoef2 iir_coef2(const IirCoef &k, float one, float neg_one)
{
    IirCoef2 r;
    r.u = unit2(one, neg_one);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        r.n2[i] = splat2(k.n2[i]);
        r.nd1[i] = splat2(-k.d1[i]);
    }
    return r;
}

__device__ __forceinline__ f32x2 iir_step2(const IirCoef2 &k, IirState2 &s, f32x2 l, f32x2 r)
{
    const f32x2 sum = add2(l, r);
    f32x2 o[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        f32x2 ok = mul2(sum, k.n2[i]);
        ok = msub2(k.u, ok, s.p2[i]);
        ok = fma2(k.nd1[i], s.p[i], ok);
        s.p2[i] = s.p[i];
        s.p[i] = ok;
        o[i] = ok;
    }
    return add2(add2(o[0], o[1]), o[2]);
}

// The other reading of the published VERTICAL pass (lib/jxl/gauss_blur.cc VerticalBlock as recalled; the oracle's
// ORACLE_VARIANT_VERTICAL_ORDER): out_k = fma(n2, sum, fma(-d1, y[n-1], -y[n-2])) — the product n2 * sum is not
// rounded on its own and -d1 * y[n-1] - y[n-2] is formed first.  OAVIF_SSIMU2_OPT_VERTICAL_ORDER; columns pass only.
__device__ __forceinline__ float iir_step_vorder(const IirCoef &k, IirState &s, float sum)
{
    float o[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float ok = fmaf(k.n2[i], sum, fmaf(-k.d1[i], s.p[i], -s.p2[i]));
        s.p2[i] = s.p[i];
        s.p[i] = ok;
        o[i] = ok;
    }
    return (o[0] + o[1]) + o[2];
}
__device__ __forceinline__ f32x2 iir_step2_vorder(const IirCoef2 &k, IirState2 &s, f32x2 sum)
{
    f32x2 o[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        // -y[n-2] as y[n-2] * -1 (exact); a product in the ADDEND position cannot be contracted into anything
        const f32x2 ok = fma2(k.n2[i], sum, fma2(k.nd1[i], s.p[i], mul2(s.p2[i], k.u.m1)));
        s.p2[i] = s.p[i];
        s.p[i] = ok;
        o[i] = ok;
    }
    return add2(add2(o[0], o[1]), o[2]);
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }

constexpr int kIirRows = 32;     // rows per rows-pass task
constexpr int kIirChunk = 32;    // columns per staged tile
constexpr int kIirPitch = 36;    // smem tile pitch in floats: 16-byte rows, conflict-free 128-bit access
constexpr int kIirPairPitch = 68;  // staging pitch of an interleaved (x, x*x) row: 64 floats + 16 bytes
constexpr int kIirVCols = 32;    // columns per columns-pass task

struct IirArgs {
    Geom g;
    IirCoef k;
    const float *src;
    const float *dist;
    long long dist_stride;      // floats between candidates' pyramids
    // Row-filtered planes.  The pairs (a, a*a) and (b, b*b) are stored INTERLEAVED per pixel (one float2 per
    // pixel: the rows pass produces them that way and the columns pass consumes them that way), pyramid
    // layout with every offset and pitch doubled; a*b is a plain pyramid.  (a, a*a) only depends on the
    // source and lives in a per-source cache; the other two are per candidate.
    float *hpair_src;           // [2 * pyramid]
    float *hpair_cand;          // [candidate][2 * pyramid], stride hcand_stride
    float *hab;                 // [candidate][pyramid],     stride hcand_stride
    long long hcand_stride;
    double *partials;
    long long partials_stride;
    float one, neg_one;         // 1.0f and -1.0f as run-time values (Unit2)
    // oavif_ssimu2_debug_get_cols: when set, the columns pass of (dbg_scale, dbg_channel, dbg_cand) also copies
    // the five blurred values it hands to the maps — mu1, mu2, s11, s22, s12 — to dbg_cols[q][h][w] (tight)
    float *dbg_cols;
    int dbg_scale, dbg_channel, dbg_cand;
    int dbg_strip_major;   // timing experiment only (debug_time_rows | 16384): the columns pass addresses its descriptors
                           // as {strip row, image row, strip, channel} — see oavif_ssimu2_debug_time_rows
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

// ... and on packed pairs
struct IirPipe2 {
    f32x2 p1[3], u[3];
};

__device__ __forceinline__ void pipe2_begin(const IirCoef2 &k, IirPipe2 &P, const IirState2 &st, f32x2 sum0)
{
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        P.p1[i] = st.p[i];
        P.u[i] = msub2(k.u, mul2(sum0, k.n2[i]), st.p2[i]);
    }
}

__device__ __forceinline__ f32x2 pipe2_step(const IirCoef2 &k, IirPipe2 &P, f32x2 sum_next)
{
    f32x2 nw[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) nw[i] = fma2(k.nd1[i], P.p1[i], P.u[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        P.u[i] = msub2(k.u, mul2(sum_next, k.n2[i]), P.p1[i]);
        P.p1[i] = nw[i];
    }
    return add2(add2(nw[0], nw[1]), nw[2]);
}

__device__ __forceinline__ f32x2 pipe2_end(const IirCoef2 &k, IirPipe2 &P, IirState2 &st)
{
    f32x2 nw[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        nw[i] = fma2(k.nd1[i], P.p1[i], P.u[i]);
        st.p2[i] = P.p1[i];
        st.p[i] = nw[i];
    }
    return add2(add2(nw[0], nw[1]), nw[2]);
}

// ------------------------------------------------------------------------------------------------
// rows pass.  grid = (sum over scales of 3 * ceil(h/32), n_candidates); one CTA = 32 rows of one channel.
//   MODE 1 (source rows cached):  the pair (b, b*b) and a*b                       block = 128
//   MODE 2 (first call after set_source): the same, and candidate 0's CTAs also run the pair (a, a*a)
//           of the source, which every later candidate and call reads back from its cache   block = 192
//
// Warp roles: one warp per packed pair recursion, one for a*b (lane = row), a loader and one or two
// storers.  Chunk t emits the 128-byte-aligned outputs n = 32t .. 32t+31: right taps (n+4) come from
// tiles t and t+1, left taps (n-6) from tiles t-1 and t.  While the recursion warps work on chunk t,
// the storers write chunk t-1 out of the staging tiles (whole 128-byte lines, streaming stores) and
// the loader requests tile t+4 (16-byte cp.async, zero-filled beyond the image) and waits for tile
// t+2; one block barrier per chunk.  Tile ring: t-1, t, t+1 in use, t+2 landed or landing, t+3 and
// t+4 in flight (five slots with a shorter look-ahead would let three CTAs share an SM, but measured
// 0.158 ms against 0.137 ms on a 4K frame).
constexpr int kRowSlots = 6, kRowAhead = kRowSlots - 2;

template <int MODE>
struct IirRowsSmem {
    static constexpr int NPAIR = MODE == 2 ? 2 : 1;
    float tile[2][kRowSlots][kIirRows][kIirPitch];     // [plane][ring slot][row][column]; plane 0 = b, plane 1 = a
    float pair[NPAIR][2][kIirRows][kIirPairPitch];     // filtered (x, x*x), interleaved per pixel, double-buffered
    float single[2][kIirRows][kIirPitch];              // filtered a*b, double-buffered
};

typedef float RowTile[kIirRows][kIirPitch];
typedef float RowPairStage[kIirRows][kIirPairPitch];

// Every role below executes the same barriers: one before the first chunk, then one per t = 0 .. nch.

// the packed recursion of (x, x*x) over the staged plane `tile`; inactive warps only keep the barriers
__device__ __forceinline__ void rows_pair_warp(const RowTile *tile, RowPairStage *stage, const IirCoef &k, float one,
                                               float neg_one, int lane, int nch, bool active)
{
    __syncthreads();
    if (!active) {
        for (int t = 0; t <= nch; ++t) __syncthreads();
        return;
    }
    int prev = 0, cur = 1, next = 2;           // slots of tiles t-1, t, t+1 (tile t lives in slot (t + 1) mod 6)
    const IirCoef2 k2 = iir_coef2(k, one, neg_one);
    IirState2 st;
#pragma unroll
    for (int i = 0; i < 3; ++i) st.p[i] = st.p2[i] = splat2(0.0f);
    {   // n = -4 .. -1: right taps are columns 0..3 of tile 0, left taps are padding, nothing emitted
        const float4 x = *reinterpret_cast<const float4 *>(&tile[cur][lane][0]);
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) (void)iir_step2(k2, st, splat2(0.0f), pk2(xs[i], xs[i] * xs[i]));
    }
#pragma unroll 1
    for (int t = 0; t <= nch; ++t) {
        if (t < nch) {
            // samples v[i] = plane value at column 32t - 8 + i, i = 0..43; q[i] its square
            float v[44], q[44];
            auto load4 = [&](int slot_, int j4, int at) {
                const float4 x = *reinterpret_cast<const float4 *>(&tile[slot_][lane][4 * j4]);
                v[at] = x.x; v[at + 1] = x.y; v[at + 2] = x.z; v[at + 3] = x.w;
                unpk2(mul2(pk2(x.x, x.y), pk2(x.x, x.y)), q[at], q[at + 1]);
                unpk2(mul2(pk2(x.z, x.w), pk2(x.z, x.w)), q[at + 2], q[at + 3]);
            };
            load4(prev, 6, 0);
            load4(prev, 7, 4);
#pragma unroll
            for (int j4 = 0; j4 < kIirChunk / 4; ++j4) load4(cur, j4, 8 + 4 * j4);
            load4(next, 0, 40);
            // step j (output column 32t + j): left tap column 32t + j - 6 = [j+2], right tap 32t + j + 4 = [j+12]
            float4 *o4 = reinterpret_cast<float4 *>(&stage[t & 1][lane][0]);
            IirPipe2 P;
            pipe2_begin(k2, P, st, pk2(v[2] + v[12], q[2] + q[12]));
#pragma unroll
            for (int j2 = 0; j2 < kIirChunk / 2; ++j2) {
                f32x2 o[2];
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const int j = 2 * j2 + jj;
                    o[jj] = (j + 1 < kIirChunk) ? pipe2_step(k2, P, pk2(v[j + 3] + v[j + 13], q[j + 3] + q[j + 13]))
                                                : pipe2_end(k2, P, st);
                }
                float4 ov;
                unpk2(o[0], ov.x, ov.y);
                unpk2(o[1], ov.z, ov.w);
                o4[j2] = ov;
            }
        }
        prev = cur;
        cur = next;
        next = next + 1 == kRowSlots ? 0 : next + 1;
        __syncthreads();
    }
}

// the scalar recursion of a*b over the two staged planes
__device__ __forceinline__ void rows_ab_warp(const RowTile *tb, const RowTile *ta, RowTile *stage, const IirCoef &k,
                                             int lane, int nch)
{
    __syncthreads();
    int prev = 0, cur = 1, next = 2;
    IirState st;
#pragma unroll
    for (int i = 0; i < 3; ++i) st.p[i] = st.p2[i] = 0.0f;
    {
        const float4 x = *reinterpret_cast<const float4 *>(&tb[cur][lane][0]);
        const float4 y = *reinterpret_cast<const float4 *>(&ta[cur][lane][0]);
        const float xs[4] = {x.x * y.x, x.y * y.y, x.z * y.z, x.w * y.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) (void)iir_step(k, st, 0.0f, xs[i]);
    }
#pragma unroll 1
    for (int t = 0; t <= nch; ++t) {
        if (t < nch) {
            float v[44];
            auto load4 = [&](int slot_, int j4, int at) {
                const float4 x = *reinterpret_cast<const float4 *>(&tb[slot_][lane][4 * j4]);
                const float4 y = *reinterpret_cast<const float4 *>(&ta[slot_][lane][4 * j4]);
                unpk2(mul2(pk2(x.x, x.y), pk2(y.x, y.y)), v[at], v[at + 1]);
                unpk2(mul2(pk2(x.z, x.w), pk2(y.z, y.w)), v[at + 2], v[at + 3]);
            };
            load4(prev, 6, 0);
            load4(prev, 7, 4);
#pragma unroll
            for (int j4 = 0; j4 < kIirChunk / 4; ++j4) load4(cur, j4, 8 + 4 * j4);
            load4(next, 0, 40);
            float sum[kIirChunk];
#pragma unroll
            for (int j = 0; j < kIirChunk; ++j) sum[j] = v[j + 2] + v[j + 12];   // scalar: both are products
            float4 *o4 = reinterpret_cast<float4 *>(&stage[t & 1][lane][0]);
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
        }
        prev = cur;
        cur = next;
        next = next + 1 == kRowSlots ? 0 : next + 1;
        __syncthreads();
    }
}

// Helpers move 16 bytes per lane: columns 4*(l&7)..+3 of rows (l>>3) + 4i, i = 0..7.  Rows beyond the image
// are clamped to a valid address and zero-filled (loads) or skipped (stores).
struct RowsLanes {
    int sub_row, sub_col;
    unsigned row_off[kIirRows / 4];
    bool row_ok[kIirRows / 4];
};

__device__ __forceinline__ RowsLanes rows_lanes(int lane, int rows_here, int pitch)
{
    RowsLanes L;
    L.sub_row = lane >> 3;
    L.sub_col = (lane & 7) * 4;
#pragma unroll
    for (int i = 0; i < kIirRows / 4; ++i) {
        const int row = L.sub_row + 4 * i;
        L.row_ok[i] = row < rows_here;
        L.row_off[i] = (unsigned)(min(row, rows_here - 1) * pitch) + L.sub_col;
    }
    return L;
}

__device__ __forceinline__ void rows_loader_warp(RowTile *tb, RowTile *ta, const float *gb, const float *ga, int w,
                                                 const RowsLanes &L, int nch)
{
    auto issue_tile = [&](int t, int slot) {
        const int gx = t * kIirChunk + L.sub_col;
        int bytes = 0;
        if (t >= 0 && gx < w) bytes = min(4, w - gx) * 4;
        const unsigned col_off = bytes ? (unsigned)(t * kIirChunk) : 0u;
        const float *c0 = gb + col_off, *c1 = ga + col_off;
#pragma unroll
        for (int i = 0; i < kIirRows / 4; ++i) {
            const int nb = L.row_ok[i] ? bytes : 0;
            cp_async_16(&tb[slot][L.sub_row + 4 * i][L.sub_col], c0 + L.row_off[i], nb);
            cp_async_16(&ta[slot][L.sub_row + 4 * i][L.sub_col], c1 + L.row_off[i], nb);
        }
        cp_async_commit();
    };
#pragma unroll 1
    for (int t = -1; t < kRowAhead; ++t) issue_tile(t, t + 1);
    cp_async_wait<kRowAhead - 2>();       // tiles -1, 0, 1 have landed
    __syncthreads();
    int slot = kRowAhead + 1;             // slot of tile t + kRowAhead
#pragma unroll 1
    for (int t = 0; t <= nch; ++t) {
        issue_tile(t + kRowAhead, slot);  // into the slot of tile t-2: nobody reads it any more
        slot = slot + 1 == kRowSlots ? 0 : slot + 1;
        cp_async_wait<kRowAhead - 2>();   // tile t+2 has landed
        __syncthreads();
    }
}

// chunk t-1 leaves staging buffer (t-1) & 1 while chunk t is computed.  A pair row of a chunk is 256
// contiguous bytes (32 pixels x float2): 16 lanes per row, two rows per instruction; an a*b row is 128.
template <bool WITH_SINGLE>
__device__ __forceinline__ void rows_storer_warp(const RowPairStage *pstage, const RowTile *sstage, float *opair,
                                                 float *oab, const RowsLanes &L, int lane, int rows_here, int pitch,
                                                 int nch, bool active)
{
    const int prow = lane >> 4, ppiece = (lane & 15) * 4;
    __syncthreads();
#pragma unroll 1
    for (int t = 0; t <= nch; ++t) {
        if (t > 0 && active) {
            const int buf = (t - 1) & 1;
            float *cp = opair + 2 * (t - 1) * kIirChunk + ppiece;
#pragma unroll
            for (int i = 0; i < kIirRows / 2; ++i) {
                const int row = prow + 2 * i;
                const float4 v = *reinterpret_cast<const float4 *>(&pstage[buf][row][ppiece]);
                if (row < rows_here) __stcs(reinterpret_cast<float4 *>(cp + (unsigned)(row * 2 * pitch)), v);
            }
            if (WITH_SINGLE) {
                float *c2 = oab + (unsigned)((t - 1) * kIirChunk);
#pragma unroll
                for (int i = 0; i < kIirRows / 4; ++i) {
                    const int row = L.sub_row + 4 * i;
                    const float4 l2 = *reinterpret_cast<const float4 *>(&sstage[buf][row][L.sub_col]);
                    if (L.row_ok[i]) __stcs(reinterpret_cast<float4 *>(c2 + L.row_off[i]), l2);
                }
            }
        }
        __syncthreads();
    }
}

template <int MODE>
__global__ void __launch_bounds__(MODE == 2 ? 192 : 128) k_iir_rows(const __grid_constant__ IirArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IirRowsSmem<MODE> &sm = *reinterpret_cast<IirRowsSmem<MODE> *>(smem_raw);
    constexpr int NW = MODE == 2 ? 6 : 4;

    int s, c, rb;
    decode_cta(a, blockIdx.x, s, c, rb);
    const int cand = blockIdx.y;
    const int w = a.g.w[s], h = a.g.h[s], pitch = a.g.pitch[s];
    const int y0 = rb * kIirRows;
    const int rows_here = min(kIirRows, h - y0);
    const long long poff = a.g.off[s] + (long long)c * a.g.plane[s] + (long long)y0 * pitch;
    // Roles rotate with the CTA index: hardware warp w of a CTA runs on sub-partition w % 4, and the
    // recursion warps are the long poles — co-resident CTAs should not stack them on one sub-partition.
    const int lane = threadIdx.x & 31;
    const int role = (int)(((threadIdx.x >> 5) + blockIdx.x) % NW);
    const int nch = (w + kIirChunk - 1) / kIirChunk;
    const bool with_src = MODE == 2 && cand == 0;
    const float *gb = a.dist + (long long)cand * a.dist_stride + poff, *ga = a.src + poff;

    switch (role) {
    case 0:   // (b, b*b)
        rows_pair_warp(sm.tile[0], sm.pair[0], a.k, a.one, a.neg_one, lane, nch, true);
        break;
    case 1:   // a*b
        rows_ab_warp(sm.tile[0], sm.tile[1], sm.single, a.k, lane, nch);
        break;
    case 2:
        rows_loader_warp(sm.tile[0], sm.tile[1], gb, ga, w, rows_lanes(lane, rows_here, pitch), nch);
        break;
    case 3:
        rows_storer_warp<true>(sm.pair[0], sm.single, a.hpair_cand + (long long)cand * a.hcand_stride + 2 * poff,
                               a.hab + (long long)cand * a.hcand_stride + poff, rows_lanes(lane, rows_here, pitch), lane,
                               rows_here, pitch, nch, true);
        break;
    case 4:   // MODE 2: (a, a*a) of the source, by candidate 0's CTAs
        rows_pair_warp(sm.tile[1], sm.pair[IirRowsSmem<MODE>::NPAIR - 1], a.k, a.one, a.neg_one, lane, nch, with_src);
        break;
    default:
        rows_storer_warp<false>(sm.pair[IirRowsSmem<MODE>::NPAIR - 1], sm.single, a.hpair_src + 2 * poff, nullptr,
                                rows_lanes(lane, rows_here, pitch), lane, rows_here, pitch, nch, with_src);
        break;
    }
}

// ------------------------------------------------------------------------------------------------
// rows pass, TMA form (the default).  Same decomposition and the same arithmetic as k_iir_rows above — one CTA =
// 32 rows of one channel, lane = row, a packed-pair recursion warp for (b, b*b), a scalar one for a*b and, on
// the first call after set_source, a pair warp for the source's (a, a*a) — but no loader warp, no storer warps and
// no block barrier:
//   * one elected lane of a producer warp issues, per 32-column chunk, two cp.async.bulk.tensor loads (planes b
//     and a) of a 44-column x 32-row box starting 8 columns left of the chunk: the chunk's own samples, the
//     left taps (n-6) and the right taps (n+4) in one tile.  Columns left of 0, right of w and rows below h
//     arrive as zeros — the filter's zero padding.  Completion is counted on the ring slot's `full` mbarrier.
//     A 44-float row pitch keeps the lane = row 128-bit reads conflict-free without padding tricks.
//   * each recursion warp waits on `full[slot]`, pulls its 44 samples into registers, and releases the slot on
//     `empty[slot]` at once; the warps never wait for each other.
//   * a finished 32 x 32 chunk is staged in 128-byte-swizzled shared memory (lane = row 128-bit stores land on
//     32 distinct banks) and leaves through cp.async.bulk.tensor stores issued by the warp's own lane 0; rows
//     below h and columns right of w are clipped by the hardware.  Two staging buffers per warp: the buffer of
//     chunk t is reused once the store of chunk t-2 has read it (cp.async.bulk.wait_group.read 1).
constexpr int kRtTileW = 44;                              // staged columns per chunk: 8 of history + 32 + 4 ahead
constexpr int kRtTileBytes = kIirRows * kRtTileW * 4;     // 5632 bytes per plane and chunk

struct IirRowsTmaMaps {   // one descriptor per scale; see iir_rows_tma_maps() for the shapes
    CUtensorMap in_src[kMaxScales], in_dist[kMaxScales];
    CUtensorMap out_psrc[kMaxScales], out_pcand[kMaxScales], out_ab[kMaxScales];
};

// STAGES: tile ring depth (STAGES - 1 chunks of look-ahead).  NBUF: staging buffers per recursion warp (the
// buffer of chunk t is reused once the store of chunk t - NBUF has read it).
// MODE 1: the candidate's half — (b, b*b) and a*b.  MODE 3: the source's half alone — (a, a*a), run once per source
// at set_source time on the context's source stream.  MODE 2: both halves in one launch (A/B timing only).
template <int MODE, int STAGES, int NBUF>
struct IirRowsTmaSmem {
    static constexpr int NPAIR = MODE == 2 ? 2 : 1;
    static constexpr int NPLANES = MODE == 3 ? 1 : 2;
    static constexpr int NSINGLE = MODE == 3 ? 1 : NBUF;
    float pstage[NPAIR][NBUF][2][kIirRows][32];      // [pair][buffer][half of a 64-float pair row][row][32]: 4 KB blocks, 128B-swizzled
    float sstage[NSINGLE][kIirRows][32];             // a*b: [buffer][row][32], 128B-swizzled
    float tile[NPLANES][STAGES][kIirRows][kRtTileW]; // [plane: 0 = b, 1 = a; MODE 3: 0 = a][ring slot][row][column]
    uint64_t full[STAGES], empty[STAGES];
};

typedef float RtTile[kIirRows][kRtTileW];
typedef float RtBlock[kIirRows][32];

template <int STAGES, int NBUF, bool NOCOMP>
__device__ __forceinline__ void rows_tma_pair_warp(const RtTile *tile, RtBlock (*stage)[2], uint64_t *full, uint64_t *empty,
                                                   const CUtensorMap *omap, const IirCoef &k, float one, float neg_one,
                                                   int lane, int nch, int y0, int c, int cand)
{
    const IirCoef2 k2 = iir_coef2(k, one, neg_one);
    IirState2 st;
#pragma unroll
    for (int i = 0; i < 3; ++i) st.p[i] = st.p2[i] = splat2(0.0f);
    const int sw = lane & 7;
    int slot = 0, buf = 0;
    uint32_t phase = 0;
#pragma unroll 1
    for (int t = 0; t < nch; ++t) {
        mbar_wait(&full[slot], phase);
        // samples v[i] = plane value at column 32t - 8 + i, i = 0..43; q[i] its square
        float v[kRtTileW], q[kRtTileW];
        const float *row = &tile[slot][lane][0];
#pragma unroll
        for (int j4 = 0; j4 < kRtTileW / 4; ++j4) {
            const float4 x = *reinterpret_cast<const float4 *>(row + 4 * j4);
            v[4 * j4] = x.x; v[4 * j4 + 1] = x.y; v[4 * j4 + 2] = x.z; v[4 * j4 + 3] = x.w;
            unpk2(mul2(pk2(x.x, x.y), pk2(x.x, x.y)), q[4 * j4], q[4 * j4 + 1]);
            unpk2(mul2(pk2(x.z, x.w), pk2(x.z, x.w)), q[4 * j4 + 2], q[4 * j4 + 3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);     // the tile is in registers: the producer may refill the slot
        if (++slot == STAGES) { slot = 0; phase ^= 1u; }
        if (t == 0) {   // n = -4 .. -1: right taps are columns 0..3, left taps are padding, nothing emitted
#pragma unroll
            for (int i = 0; i < 4; ++i) (void)iir_step2(k2, st, splat2(0.0f), pk2(v[8 + i], q[8 + i]));
        }
        if (t >= NBUF) {   // this staging buffer went out with chunk t - NBUF
            if (lane == 0) tma_store_wait_read<NBUF - 1>();
            __syncwarp();
        }
        // step j (output column 32t + j): left tap column 32t + j - 6 = [j+2], right tap 32t + j + 4 = [j+12]
        RtBlock *blk = stage[buf];
        IirPipe2 P;
        pipe2_begin(k2, P, st, pk2(v[2] + v[12], q[2] + q[12]));
#pragma unroll
        for (int j2 = 0; j2 < kIirChunk / 2; ++j2) {
            f32x2 o[2];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int j = 2 * j2 + jj;
                if (NOCOMP)   // timing experiment only: the data movement without the recursion
                    o[jj] = pk2(v[j + 8], q[j + 8]);
                else
                    o[jj] = (j + 1 < kIirChunk) ? pipe2_step(k2, P, pk2(v[j + 3] + v[j + 13], q[j + 3] + q[j + 13]))
                                                : pipe2_end(k2, P, st);
            }
            float4 ov;
            unpk2(o[0], ov.x, ov.y);
            unpk2(o[1], ov.z, ov.w);
            // 16-byte piece j2 of the 256-byte pair row: block j2 / 8, piece j2 % 8, swizzled with the row
            *reinterpret_cast<float4 *>(&blk[j2 >> 3][lane][((j2 & 7) ^ sw) << 2]) = ov;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            tma_store_4d(omap, 64 * t, y0, c, cand, &blk[0][0][0]);
            tma_store_4d(omap, 64 * t + 32, y0, c, cand, &blk[1][0][0]);
            tma_store_commit();
        }
        if (++buf == NBUF) buf = 0;
    }
    if (lane == 0) tma_store_wait_all<0>();
}

template <int STAGES, int NBUF, bool NOCOMP>
__device__ __forceinline__ void rows_tma_ab_warp(const RtTile *tb, const RtTile *ta, RtBlock *stage, uint64_t *full,
                                                 uint64_t *empty, const CUtensorMap *omap, const IirCoef &k, int lane,
                                                 int nch, int y0, int c, int cand)
{
    IirState st;
#pragma unroll
    for (int i = 0; i < 3; ++i) st.p[i] = st.p2[i] = 0.0f;
    const int sw = lane & 7;
    int slot = 0, buf = 0;
    uint32_t phase = 0;
#pragma unroll 1
    for (int t = 0; t < nch; ++t) {
        mbar_wait(&full[slot], phase);
        float v[kRtTileW];   // a*b at column 32t - 8 + i
        const float *rb = &tb[slot][lane][0], *ra = &ta[slot][lane][0];
#pragma unroll
        for (int j4 = 0; j4 < kRtTileW / 4; ++j4) {
            const float4 x = *reinterpret_cast<const float4 *>(rb + 4 * j4);
            const float4 y = *reinterpret_cast<const float4 *>(ra + 4 * j4);
            unpk2(mul2(pk2(x.x, x.y), pk2(y.x, y.y)), v[4 * j4], v[4 * j4 + 1]);
            unpk2(mul2(pk2(x.z, x.w), pk2(y.z, y.w)), v[4 * j4 + 2], v[4 * j4 + 3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
        if (++slot == STAGES) { slot = 0; phase ^= 1u; }
        if (t == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) (void)iir_step(k, st, 0.0f, v[8 + i]);
        }
        if (t >= NBUF) {
            if (lane == 0) tma_store_wait_read<NBUF - 1>();
            __syncwarp();
        }
        float sum[kIirChunk];
#pragma unroll
        for (int j = 0; j < kIirChunk; ++j) sum[j] = v[j + 2] + v[j + 12];   // scalar: both are products
        float(*blk)[32] = stage[buf];
        IirPipe P;
        pipe_begin(k, P, st, sum[0]);
#pragma unroll
        for (int j4 = 0; j4 < kIirChunk / 4; ++j4) {
            float o[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int j = 4 * j4 + jj;
                o[jj] = NOCOMP ? v[j + 8] : ((j + 1 < kIirChunk) ? pipe_step(k, P, sum[j + 1]) : pipe_end(k, P, st));
            }
            *reinterpret_cast<float4 *>(&blk[lane][(j4 ^ sw) << 2]) = make_float4(o[0], o[1], o[2], o[3]);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            tma_store_4d(omap, 32 * t, y0, c, cand, &blk[0][0]);
            tma_store_commit();
        }
        if (++buf == NBUF) buf = 0;
    }
    if (lane == 0) tma_store_wait_all<0>();
}

template <int MODE, int STAGES, int NBUF, bool NOCOMP = false>
__global__ void __launch_bounds__(MODE == 2 ? 128 : MODE == 3 ? 64 : 96)
    k_iir_rows_tma(const __grid_constant__ IirArgs a, const __grid_constant__ IirRowsTmaMaps tm)
{
    // 128-byte-swizzled staging blocks need 1024-byte alignment.  The array is declared with it (no pointer
    // arithmetic on the base: that would turn every shared-memory access into a generic one) and checked once.
    extern __shared__ __align__(1024) unsigned char smem_swz[];
    typedef IirRowsTmaSmem<MODE, STAGES, NBUF> Smem;
    Smem &sm = *reinterpret_cast<Smem *>(smem_swz);
    constexpr int NW = MODE == 2 ? 4 : MODE == 3 ? 2 : 3;

    int s, c, rb;
    decode_cta(a, blockIdx.x, s, c, rb);
    const int cand = blockIdx.y;
    const int w = a.g.w[s];
    const int y0 = rb * kIirRows;
    const int lane = threadIdx.x & 31;
    // roles rotate with the CTA index so that co-resident CTAs do not stack their recursion warps on one
    // sub-partition (hardware warp w runs on sub-partition w % 4)
    const int role = (int)(((threadIdx.x >> 5) + blockIdx.x) % NW);
    const int nch = (w + kIirChunk - 1) / kIirChunk;
    const bool with_src = MODE == 2 && cand == 0;

    if (threadIdx.x == 0) {
        if (smem_u32(smem_swz) & 1023u) __trap();
#pragma unroll
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&sm.full[i], 1);
            mbar_init(&sm.empty[i], MODE == 3 ? 1 : (with_src ? 3 : 2));   // one arrival per recursion warp that reads the slot
        }
        mbar_init_fence();
    }
    __syncthreads();

    if (MODE == 3) {
        if (role == 0) {          // (a, a*a)
            rows_tma_pair_warp<STAGES, NBUF, NOCOMP>(sm.tile[0], sm.pstage[0], sm.full, sm.empty, &tm.out_psrc[s], a.k, a.one,
                                                     a.neg_one, lane, nch, y0, c, 0);
        } else if (lane == 0) {   // producer
            tma_prefetch_desc(&tm.in_src[s]);
            int slot = 0;
            uint32_t phase = 1;
#pragma unroll 1
            for (int t = 0; t < nch; ++t) {
                mbar_wait(&sm.empty[slot], phase);
                mbar_arrive_expect_tx(&sm.full[slot], kRtTileBytes);
                tma_load_4d(&sm.tile[0][slot][0][0], &tm.in_src[s], kIirChunk * t - 8, y0, c, 0, &sm.full[slot]);
                if (++slot == STAGES) { slot = 0; phase ^= 1u; }
            }
        }
        return;
    }
    constexpr int PA = Smem::NPLANES - 1;   // plane a's tiles
    switch (role) {
    case 0:   // (b, b*b)
        rows_tma_pair_warp<STAGES, NBUF, NOCOMP>(sm.tile[0], sm.pstage[0], sm.full, sm.empty, &tm.out_pcand[s], a.k, a.one, a.neg_one,
                                         lane, nch, y0, c, cand);
        break;
    case 1:   // a*b
        rows_tma_ab_warp<STAGES, NBUF, NOCOMP>(sm.tile[0], sm.tile[PA], sm.sstage, sm.full, sm.empty, &tm.out_ab[s], a.k, lane, nch, y0,
                                       c, cand);
        break;
    case 2:   // producer: one lane feeds the ring
        if (lane == 0) {
            tma_prefetch_desc(&tm.in_dist[s]);
            tma_prefetch_desc(&tm.in_src[s]);
            int slot = 0;
            uint32_t phase = 1;   // waiting on the phase BEFORE the first passes at once: the ring starts empty
#pragma unroll 1
            for (int t = 0; t < nch; ++t) {
                mbar_wait(&sm.empty[slot], phase);
                mbar_arrive_expect_tx(&sm.full[slot], 2 * kRtTileBytes);
                tma_load_4d(&sm.tile[0][slot][0][0], &tm.in_dist[s], kIirChunk * t - 8, y0, c, cand, &sm.full[slot]);
                tma_load_4d(&sm.tile[PA][slot][0][0], &tm.in_src[s], kIirChunk * t - 8, y0, c, 0, &sm.full[slot]);
                if (++slot == STAGES) { slot = 0; phase ^= 1u; }
            }
        }
        break;
    default:  // MODE 2: (a, a*a) of the source, by candidate 0's CTAs
        if (with_src)
            rows_tma_pair_warp<STAGES, NBUF, NOCOMP>(sm.tile[PA], sm.pstage[Smem::NPAIR - 1], sm.full, sm.empty, &tm.out_psrc[s], a.k,
                                             a.one, a.neg_one, lane, nch, y0, c, 0);
        break;
    }
}

// the shipped instance, and what oavif_ssimu2_debug_time_rows can time against it (variant bits 4..6)
constexpr int kRtStages = 4, kRtBufs = 3;

template <int MODE, int STAGES, int NBUF, bool NOCOMP = false>
inline cudaError_t iir_rows_tma_launch(const IirArgs &ar, const IirRowsTmaMaps &maps, int nr, int n, cudaStream_t st)
{
    typedef IirRowsTmaSmem<MODE, STAGES, NBUF> Smem;
    // the attribute belongs to (kernel instance, DEVICE): the corpus driver runs contexts on several devices in one
    // process.  Benign race between workers of one device: they store the same value.
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        const cudaError_t e = cudaFuncSetAttribute(k_iir_rows_tma<MODE, STAGES, NBUF, NOCOMP>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    k_iir_rows_tma<MODE, STAGES, NBUF, NOCOMP><<<dim3(nr, n), MODE == 2 ? 128 : MODE == 3 ? 64 : 96, sizeof(Smem), st>>>(ar, maps);
    return cudaGetLastError();
}

template <int MODE>
inline cudaError_t iir_rows_tma_dispatch(int shape, const IirArgs &ar, const IirRowsTmaMaps &maps, int nr, int n, cudaStream_t st)
{
    switch (shape) {
    case 1: return iir_rows_tma_launch<MODE, 4, 2>(ar, maps, nr, n, st);
    case 2: return iir_rows_tma_launch<MODE, 6, 3>(ar, maps, nr, n, st);
    case 4: return iir_rows_tma_launch<MODE, 4, 3, true>(ar, maps, nr, n, st);   // data movement only (timing experiment)
    default: return iir_rows_tma_launch<MODE, kRtStages, kRtBufs>(ar, maps, nr, n, st);
    }
}

// ------------------------------------------------------------------------------------------------
// columns pass + maps + pooling.  grid = (sum over scales of 3 * ceil(w/32), n_candidates), block = 256.
//
// One CTA owns 32 columns of one channel; lane = column in the arithmetic, so the recursions of a warp
// are 32 independent columns.  Warps 0..2 (producers) run the column recursions out of shared-memory
// rings: warp 0 the packed pair (a, a*a), warp 1 the packed pair (b, b*b) — each reads one float2 per
// tap from its interleaved plane —, warp 2 a*b alone.  Warp 7 (loader) feeds the three rings 32 rows ahead
// of use — TMA form: one lane, 4-row cp.async.bulk.tensor boxes counted on an mbarrier per request group, the
// hardware zero-filling rows and columns beyond the image; cp.async form: the whole warp, 16 bytes per lane —;
// both taps of the recursion are read back from the ring (no register delay line).  Each producer drops its
// filtered values into a double-buffered 16-row batch.  Warps 3..6 (consumers), one batch behind, evaluate the
// SSIM / edge-diff maps and the six pooled sums, two rows of a column as one packed pair, from the batch and the
// pixel's own XYB samples (TMA form: brought in by the loader with everything else; cp.async form: staged by each
// consumer).  One block barrier per 16 rows (DECOUPLED instance: per-hand-over mbarriers instead).
constexpr int kIirVThreads = 256;   // 3 producer warps + 4 consumer warps + 1 loader warp

constexpr int kAbRows = 48;
// TMA form of the columns pass, DECOUPLED instance (OAVIF_SSIMU2_TILES_TMA_DECOUPLED): the warps of a CTA meet at no
// CTA-wide barrier inside the batch loop.  Each hand-over is an mbarrier of its own — group landed (loader ->
// everyone, passed on once through ready[]), batch produced (3 producer warps -> consumers, loader), batch consumed
// (4 consumer warps -> producers, loader) — so a warp waits only for what it reads.  Four barriers per kind although
// the data is double-buffered: a waiter tests a phase PARITY, and with four the next phase of the same parity cannot
// complete before every waiter of this one has passed (argued at each wait below).
// Measured (profiles/r2_cols_sync_forms.txt): the same duration as the one-__syncthreads-per-batch instance to 1 % —
// the barrier samples of the rendezvous form were warps waiting for work that is limited elsewhere, not time lost at
// the barrier.  Both are bit-identical and run under the same tests; the default tile path uses the rendezvous
// instance (fewer executed instructions, one barrier to reason about).

template <int RCAP, int B>
struct IirColsSmem {
    float pring[2][RCAP][2 * kIirVCols];       // producer input rows of the pairs (a, a*a), (b, b*b), row r at [r & (RCAP-1)]
    float sring[RCAP][kIirVCols];              // ... and of a*b
    float ab[2][kAbRows][kIirVCols];           // consumer rows of the two XYB planes: cp.async form rows r at [r & 31];
                                               // TMA form three 16-row batches, batch b at [16 * (b % 3)]
    float ex[2][5][B][kIirVCols];              // filtered values, double-buffered
    double red[4][6];
    uint64_t land[4];                          // TMA form: "request group g has landed", g & 3
    uint64_t full[4], empty[4];                // TMA form: batch b is in ex[b & 1] (3 producers) / has been consumed (4 consumers), b & 3
    uint64_t ready[4];                         // TMA form: the loader's "group g has landed", told to the other warps once
};

// TMA descriptors of the columns pass: 4-row boxes of the interleaved pair planes (64 floats wide) and of a*b (32),
// 16-row boxes of the two XYB pyramids (xa: source, xb: candidates) for the consumers
struct IirColsTmaMaps {
    CUtensorMap psrc[kMaxScales], pcand[kMaxScales], ab[kMaxScales], xa[kMaxScales], xb[kMaxScales];
};

// B rows per exchange batch: the per-batch overhead (barrier, copy issue, address set-up) is paid once
// per B rows by each of the eight warps.  RCAP = rows of a producer ring.
// TAP: the test hook of oavif_ssimu2_debug_get_cols, compiled into a second instance so that the scored path's
// kernel carries none of it (with the hook inline the kernel went from 70 to 128 registers).
// TMA: the loader warp's 16-byte cp.async traffic (about 200 instructions per batch, a tenth of the CTA's) becomes
// twelve cp.async.bulk.tensor issues by one lane; rows below the image and columns right of it arrive as zeros.
template <int RCAP, int B, bool TAP, bool TMA, bool DECOUPLED = false, bool VORDER = false>
__global__ void __launch_bounds__(kIirVThreads) k_iir_cols(const __grid_constant__ IirArgs a, const __grid_constant__ IirColsTmaMaps tm)
{
    extern __shared__ __align__(128) unsigned char smem_cols[];
    IirColsSmem<RCAP, B> &sm = *reinterpret_cast<IirColsSmem<RCAP, B> *>(smem_cols);
    constexpr int D = ((RCAP - B - 10) / B) * B;   // rows of look-ahead: D + B + 10 <= RCAP, D % B == 0
    constexpr int DA = 16;                         // consumer look-ahead (ring of 32 rows)
    static_assert(D >= B && DA % B == 0 && DA + B <= 32 && B % 8 == 0 && RCAP % B == 0, "ring geometry");

    int s, c, cb;
    decode_cta(a, blockIdx.x, s, c, cb);
    const int cand = blockIdx.y;
    const int w = a.g.w[s], h = a.g.h[s], pitch = a.g.pitch[s];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // copy role of a lane: row (lane >> 3) of a 4-row group, columns 4*(lane & 7)..+3 of the 32
    const int crow = lane >> 3, ccol = (lane & 7) * 4;
    const long long poff = a.g.off[s] + (long long)c * a.g.plane[s] + cb * kIirVCols;
    const int nbatch = (h + B - 1) / B;
    // bytes of this lane's 16-byte column group that lie inside the image: the rest is zero-filled
    const int cbytes = max(0, min(16, (w - (cb * kIirVCols + ccol)) * 4));
    constexpr bool DEC = TMA && DECOUPLED;
    auto wait_land = [&](int g) { mbar_wait(&sm.land[g & 3], (unsigned)(g >> 2) & 1u); };
    // Every completed copy of a group wakes the warps parked on land[]; only the loader's lane parks there and passes
    // the completed group on through ready[] (one arrival), so that seven warps wake once per group, not thirteen times.
    auto wait_ready = [&](int g) { mbar_wait(&sm.ready[g & 3], (unsigned)(g >> 2) & 1u); };
    auto wait_full = [&](int b) { mbar_wait(&sm.full[b & 3], (unsigned)(b >> 2) & 1u); };
    auto wait_empty = [&](int b) { mbar_wait(&sm.empty[b & 3], (unsigned)(b >> 2) & 1u); };
    auto warp_arrive = [&](uint64_t *bar) {   // the whole warp's shared-memory accesses precede the arrival
        __syncwarp();
        if (lane == 0) mbar_arrive(bar);
    };

    if (TMA && warp == 7) {
        // ---------------- loader, TMA form: one lane, 4-row boxes (a box never wraps around the 64-row ring) ----------------
        // Requests of a batch complete on land[parity of the batch]; the loader arrives at the barrier that ends batch
        // b only when the rows requested during batch b - 1 (everything batch b + 1 reads) have landed.
        constexpr unsigned kGroupBytes = 4 * (2 * 2 * kIirVCols + kIirVCols) * 4;   // 4 rows of two pair planes and a*b
        constexpr unsigned kXybBytes = 2 * B * kIirVCols * 4;                       // a batch of both XYB planes
        const int x0 = cb * kIirVCols;
        const bool sm_maj = a.dbg_strip_major != 0;   // (timing experiment: same byte counts, contiguous strips)
        auto issue_rows4 = [&](int r0, uint64_t *bar) {
            tma_load_4d(&sm.pring[0][r0 & (RCAP - 1)][0], &tm.psrc[s], sm_maj ? 0 : 2 * x0, r0, sm_maj ? cb : c, sm_maj ? c : 0, bar);
            tma_load_4d(&sm.pring[1][r0 & (RCAP - 1)][0], &tm.pcand[s], sm_maj ? 0 : 2 * x0, r0, sm_maj ? cb : c, sm_maj ? c : cand, bar);
            tma_load_4d(&sm.sring[r0 & (RCAP - 1)][0], &tm.ab[s], sm_maj ? 0 : x0, r0, sm_maj ? cb : c, sm_maj ? c : cand, bar);
        };
        // The consumers' XYB rows of batch nb go to third nb % 3 of their ring.  They are requested while batch
        // nb - 1 is being produced: the consumers are then reading batch nb - 2's third and will read batch nb - 1's
        // next, so the third being refilled is the one of batch nb - 3, which nobody touches any more.
        auto issue_xyb = [&](int nb, int third, uint64_t *bar) {
            tma_load_4d(&sm.ab[0][third * B][0], &tm.xa[s], sm_maj ? 0 : x0, nb * B, sm_maj ? cb : c, sm_maj ? c : 0, bar);
            tma_load_4d(&sm.ab[1][third * B][0], &tm.xb[s], sm_maj ? 0 : x0, nb * B, sm_maj ? cb : c, sm_maj ? c : cand, bar);
        };
#pragma unroll
        for (int j = 1; j <= 6; ++j) {   // rows -6..-1 are padding
            sm.pring[0][RCAP - j][lane] = sm.pring[0][RCAP - j][lane + 32] = 0.0f;
            sm.pring[1][RCAP - j][lane] = sm.pring[1][RCAP - j][lane + 32] = 0.0f;
            sm.sring[RCAP - j][lane] = 0.0f;
        }
        fence_async_smem();   // these slots are rewritten by TMA once real rows wrap around to them
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                mbar_init(&sm.land[i], 1);
                mbar_init(&sm.full[i], 3);
                mbar_init(&sm.empty[i], 4);
                mbar_init(&sm.ready[i], 1);
            }
            mbar_init_fence();
            // group 0 = everything before the first batch's request (rows 0 .. 3 + D, XYB rows of batch 0), on land[0]
            mbar_arrive_expect_tx(&sm.land[0], (4 + D) / 4 * kGroupBytes + kXybBytes);
            for (int r0 = 0; r0 < 4 + D; r0 += 4) issue_rows4(r0, &sm.land[0]);
            issue_xyb(0, 0, &sm.land[0]);
            mbar_wait(&sm.land[0], 0);
            if (DEC) mbar_arrive(&sm.ready[0]);   // nobody waits for group 0, but ready[0]'s phases must count from it
        }
        __syncthreads();      // (S) barriers initialised, rows 0 .. 3+D (all that batches 0 and 1 read) in the rings
        int third = 1;        // (b + 1) % 3
#pragma unroll 1
        for (int b = 0; b < nbatch; ++b) {
            if (lane == 0) {
                const int g = b + 1;                 // group of this batch's requests
                if (DEC) {
                    // The rows of group g replace rows (b-2)B+4 .. (b-1)B+3, last read by the producers' batch b - 1;
                    // the XYB third replaces batch b - 2's.  Parity: the next same-parity phase of full[(b-1)&3] is
                    // batch b + 3, whose rows THIS lane requests two iterations from now; that of empty[(b-2)&3] is
                    // batch b + 2, produced only from the rows requested right below.
                    if (b >= 1) wait_full(b - 1);
                    if (b >= 2) wait_empty(b - 2);
                }
                mbar_arrive_expect_tx(&sm.land[g & 3], B / 4 * kGroupBytes + kXybBytes);
#pragma unroll
                for (int j = 0; j < B; j += 4) issue_rows4(b * B + 4 + D + j, &sm.land[g & 3]);
                issue_xyb(b + 1, third, &sm.land[g & 3]);
                if (b >= 1) {
                    wait_land(b);                    // group b: requested during batch b - 1
                    // ready[b&3]'s next same-parity phase is group b + 4, passed on only after full(b + 3): by then
                    // every producer has started batch b + 1 and (empty(b + 1)) every consumer has finished batch b
                    if (DEC) mbar_arrive(&sm.ready[b & 3]);
                }
            }
            third = third == 2 ? 0 : third + 1;
            if (DEC) __syncwarp();
            else __syncthreads();         // (b)
        }
        // the last group's rows lie below the image (zeros), but they still land in this CTA's shared memory: they
        // must have done so before the CTA gives it up
        if (lane == 0) wait_land(nbatch);
        __syncthreads();      // consumers' last batch
        __syncthreads();      // final reduction
    } else if (warp == 7) {
        // ---------------- loader: feeds the three producer rings ----------------
        // Batch b (rows n0 = 16b ..) reads ring rows n0-6 .. n0+B+3.  The loader requests rows n0+4+D .. n0+3+D+B
        // while batch b runs (their slots held rows n0-28 .. n0-13, dead by then) and arrives at the barrier
        // that ends batch b only when everything batch b+1 reads has landed.
        // pair rows are 256 bytes: 16 lanes per row, two rows per instruction; a*b rows 128 bytes: four rows
        const int prow = lane >> 4, pcol = (lane & 15) * 4;
        const int pbytes = max(0, min(2, w - (cb * kIirVCols + (lane & 15) * 2))) * 8;
        const float *gp0 = a.hpair_src + 2 * poff + pcol;
        const float *gp1 = a.hpair_cand + (long long)cand * a.hcand_stride + 2 * poff + pcol;
        const float *gs = a.hab + (long long)cand * a.hcand_stride + poff + ccol;
        auto issue_rows4 = [&](int r0) {   // rows r0..r0+3 (zeros beyond h) of all planes
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int rr = r0 + 2 * half + prow;
                const unsigned go = (unsigned)(min(rr, h - 1) * 2 * pitch);
                const int nb = rr < h ? pbytes : 0;
                cp_async_16(&sm.pring[0][rr & (RCAP - 1)][pcol], gp0 + go, nb);
                cp_async_16(&sm.pring[1][rr & (RCAP - 1)][pcol], gp1 + go, nb);
            }
            const int rr = r0 + crow;
            cp_async_16(&sm.sring[rr & (RCAP - 1)][ccol], gs + (unsigned)(min(rr, h - 1) * pitch), rr < h ? cbytes : 0);
        };
        // rows -6..-1 are padding: their ring slots hold zeros until real rows wrap around to them
#pragma unroll
        for (int j = 1; j <= 6; ++j) {
            sm.pring[0][RCAP - j][lane] = sm.pring[0][RCAP - j][lane + 32] = 0.0f;
            sm.pring[1][RCAP - j][lane] = sm.pring[1][RCAP - j][lane + 32] = 0.0f;
            sm.sring[RCAP - j][lane] = 0.0f;
        }
        for (int r0 = 0; r0 < 4 + D; r0 += 4) issue_rows4(r0);   // everything before the first batch's request
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();      // (S) rows 0 .. 3+D are in the rings
#pragma unroll 1
        for (int b = 0; b < nbatch; ++b) {
#pragma unroll
            for (int j = 0; j < B; j += 4) issue_rows4(b * B + 4 + D + j);
            cp_async_commit();
            cp_async_wait<D / B - 1>();   // rows up to (b+1)*B + B + 3 have landed
            __syncthreads();              // (b)
        }
        __syncthreads();      // consumers' last batch
        __syncthreads();      // final reduction
    } else if (warp < 2) {
        // ---------------- pair producers: the packed column recursions of (a, a*a) and (b, b*b) ----------------
        // outputs: warp 0 -> mu1 (ex 0), s11 (ex 2); warp 1 -> mu2 (ex 1), s22 (ex 3)
        const IirCoef2 k = iir_coef2(a.k, a.one, a.neg_one);
        IirState2 st;
#pragma unroll
        for (int i = 0; i < 3; ++i) st.p[i] = st.p2[i] = splat2(0.0f);
        const float *col = &sm.pring[warp][0][2 * lane];
        constexpr int kRow = 2 * kIirVCols;
        __syncthreads();      // (S)
        // n = -4..-1: right taps are rows 0..3, left taps are padding, nothing emitted
#pragma unroll
        for (int n = -4; n < 0; ++n) {
            if (VORDER) (void)iir_step2_vorder(k, st, add2(splat2(0.0f), lds2(col + (n + 4) * kRow)));
            else (void)iir_step2(k, st, splat2(0.0f), lds2(col + (n + 4) * kRow));
        }

#pragma unroll 1
        for (int b = 0; b < nbatch; ++b) {
            if (DEC) {
                // batch b reads rows up to bB + 19: group b - 1 (group 0, rows 0 .. 35, landed before (S))
                if (b >= 2) wait_ready(b - 1);
                // ex[b & 1] held batch b - 2.  empty[(b-2)&3]'s next same-parity phase is batch b + 2, which this
                // warp itself must produce first.
                if (b >= 2) wait_empty(b - 2);
            }
            const int n0 = b * B;
            // n0 is a multiple of B and so is RCAP: the left taps (rows n0-6+j) can only wrap at j = 6, the
            // right taps (rows n0+4+j) only at j = B-4 -> two bases each, static offsets otherwise
            const float *l0 = col + ((n0 - 6) & (RCAP - 1)) * kRow, *l1 = col + (n0 & (RCAP - 1)) * kRow;
            const float *r0 = col + ((n0 + 4) & (RCAP - 1)) * kRow, *r1 = col + ((n0 + B) & (RCAP - 1)) * kRow;
            f32x2 sum[B];
#pragma unroll
            for (int j = 0; j < B; ++j)
                sum[j] = add2(lds2(j < 6 ? l0 + j * kRow : l1 + (j - 6) * kRow),
                              lds2(j < B - 4 ? r0 + j * kRow : r1 + (j - (B - 4)) * kRow));
            float *ex0 = &sm.ex[b & 1][warp][0][lane], *ex1 = &sm.ex[b & 1][warp + 2][0][lane];
            if (VORDER) {
#pragma unroll
                for (int j = 0; j < B; ++j) unpk2(iir_step2_vorder(k, st, sum[j]), ex0[j * kIirVCols], ex1[j * kIirVCols]);
            } else {
                IirPipe2 P;
                pipe2_begin(k, P, st, sum[0]);
#pragma unroll
                for (int j = 0; j < B; ++j) {
                    const f32x2 o = (j + 1 < B) ? pipe2_step(k, P, sum[j + 1]) : pipe2_end(k, P, st);
                    unpk2(o, ex0[j * kIirVCols], ex1[j * kIirVCols]);
                }
            }
            if (DEC) warp_arrive(&sm.full[b & 3]);
            else __syncthreads();  // (b) batch b published; the consumers are done with the other buffer
        }
        __syncthreads();      // consumers' last batch
        __syncthreads();      // final reduction
    } else if (warp == 2) {
        // ---------------- single producer: the column recursion of a*b ----------------
        const int q = 4;
        const IirCoef k = a.k;
        IirState st;
#pragma unroll
        for (int i = 0; i < 3; ++i) st.p[i] = st.p2[i] = 0.0f;
        const float *col = &sm.sring[0][lane];
        __syncthreads();      // (S)
#pragma unroll
        for (int n = -4; n < 0; ++n) {
            if (VORDER) (void)iir_step_vorder(k, st, 0.0f + col[(n + 4) * kIirVCols]);
            else (void)iir_step(k, st, 0.0f, col[(n + 4) * kIirVCols]);
        }

#pragma unroll 1
        for (int b = 0; b < nbatch; ++b) {
            if (DEC) {            // as in the pair producers
                if (b >= 2) {
                    wait_ready(b - 1);
                    wait_empty(b - 2);
                }
            }
            const int n0 = b * B;
            float sum[B];
            const float *l0 = col + ((n0 - 6) & (RCAP - 1)) * kIirVCols, *l1 = col + (n0 & (RCAP - 1)) * kIirVCols;
            const float *r0 = col + ((n0 + 4) & (RCAP - 1)) * kIirVCols, *r1 = col + ((n0 + B) & (RCAP - 1)) * kIirVCols;
#pragma unroll
            for (int j = 0; j < B; ++j)
                sum[j] = (j < 6 ? l0[j * kIirVCols] : l1[(j - 6) * kIirVCols]) +
                         (j < B - 4 ? r0[j * kIirVCols] : r1[(j - (B - 4)) * kIirVCols]);
            float *ex = &sm.ex[b & 1][q][0][lane];
            if (VORDER) {
#pragma unroll
                for (int j = 0; j < B; ++j) ex[j * kIirVCols] = iir_step_vorder(k, st, sum[j]);
            } else {
                IirPipe P;
                pipe_begin(k, P, st, sum[0]);
#pragma unroll
                for (int j = 0; j < B; ++j)
                    ex[j * kIirVCols] = (j + 1 < B) ? pipe_step(k, P, sum[j + 1]) : pipe_end(k, P, st);
            }
            if (DEC) warp_arrive(&sm.full[b & 3]);
            else __syncthreads();  // (b)
        }
        __syncthreads();
        __syncthreads();
    } else {
        // ---------------- consumers: maps + pooling ----------------
        // Four consumer warps (3..6) take two of the eight row pairs (2p, 2p+1), p = 0..7, of a 16-row batch
        // each: rows 4*cw .. 4*cw + 3.  With the loader as warp 7 the four sub-partitions of the SM (warp w
        // runs on sub-partition w % 4) carry about equal work: a pair producer costs ~260 instructions per
        // batch, the a*b producer ~260, the loader ~200, two row pairs of maps ~275.
        // A consumer evaluates the two rows of a pair as one packed pair per column, and stages the XYB rows
        // it needs itself.  Columns beyond the image need no test: every ring is zero-filled there by the
        // copies, and all-zero inputs pool to exactly zero.
        const int cw = warp - 3;                                   // 0..3
        const int first_pair = 2 * cw;
        constexpr int npairs = 2;
        const bool stager = crow < 2 * npairs;                     // lanes that copy: 8 per row
        const int srow = 2 * first_pair + crow;
        const float *pa = a.src + poff + ccol;
        const float *pb = a.dist + (long long)cand * a.dist_stride + poff + ccol;
        double dacc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        auto issue_ab = [&](int r0) {   // this lane's row of the batch starting at r0, both planes
            const int rr = r0 + srow;
            const unsigned o = (unsigned)(min(rr, h - 1) * pitch);
            const int nb = rr < h ? cbytes : 0;
            if (stager) {
                cp_async_16(&sm.ab[0][rr & 31][ccol], pa + o, nb);
                cp_async_16(&sm.ab[1][rr & 31][ccol], pb + o, nb);
            }
        };
        static_assert(B == 16 && DA == 16, "the pair assignment is written for 16-row batches");
        if (!TMA) {           // TMA form: the loader warp brings the XYB rows in with everything else
            issue_ab(0);
            cp_async_commit();
            cp_async_wait<0>();   // the in-loop wait only covers groups committed inside the loop
        }
        __syncthreads();      // (S)
        const Unit2 u = unit2(a.one, a.neg_one);
        const f32x2 zero = splat2(0.0f);
        const bool dbg = TAP && a.dbg_cols != nullptr && s == a.dbg_scale && c == a.dbg_channel && cand == a.dbg_cand;
        f32x2 acc[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[j] = zero;
        const float *abw = &sm.ab[0][2 * first_pair][lane];   // this warp's first row pair in ring rows 0..15
        constexpr int kAbPlane = kAbRows * kIirVCols;          // floats between the two staged planes
        int third = 0;                                         // b % 3
#pragma unroll 1
        for (int b = 0; b < nbatch; ++b) {
            if (!TMA) {
                issue_ab(b * B + DA);
                cp_async_commit();
                cp_async_wait<DA / B>();               // the rows of batch b staged by this warp have landed
            }
            if (DEC) {
                // full[b&3]'s next same-parity phase is batch b + 4, which needs this warp's empty(b + 2)
                wait_full(b);
                if (b >= 1) wait_ready(b);             // this batch's XYB rows ride in group b
            } else {
                __syncthreads();                       // batch b is in ex[b & 1]; staged samples are visible
            }
            const float *ex = &sm.ex[b & 1][0][2 * first_pair][lane];
            // B divides 32 (and the thirds are B rows): a batch never wraps inside the ring
            const float *ab = abw + (TMA ? third * B : (b * B) & 31) * kIirVCols;
            third = third == 2 ? 0 : third + 1;
            const bool whole = (b + 1) * B <= h;       // every row of the batch is inside the image
            // one packed evaluation of rows n, n + 1 of this warp's columns.  ODD: the pair's second row lies below an
            // odd-height image and pools as zeros — a separate instance, so that the common path carries none of the
            // fourteen predicated moves that zeroing costs
            auto eval_pair = [&](int g, int n, auto odd) {
                f32x2 in[7];
                in[0] = pk2(ab[(2 * g) * kIirVCols], ab[(2 * g + 1) * kIirVCols]);
                in[1] = pk2(ab[kAbPlane + (2 * g) * kIirVCols], ab[kAbPlane + (2 * g + 1) * kIirVCols]);
#pragma unroll
                for (int q = 0; q < 5; ++q)
                    in[2 + q] = pk2(ex[(q * B + 2 * g) * kIirVCols], ex[(q * B + 2 * g + 1) * kIirVCols]);
                if (TAP && dbg && cb * kIirVCols + lane < w) {   // test hook: what the maps are about to consume
#pragma unroll
                    for (int q = 0; q < 5; ++q) {
                        float lo, hi;
                        unpk2(in[2 + q], lo, hi);
                        float *o = a.dbg_cols + ((long long)q * h + n) * w + cb * kIirVCols + lane;
                        o[0] = lo;
                        if (n + 1 < h) o[w] = hi;
                    }
                }
                if (decltype(odd)::value) {
#pragma unroll
                    for (int i = 0; i < 7; ++i) {
                        float lo, hi;
                        unpk2(in[i], lo, hi);
                        in[i] = pk2(lo, 0.0f);
                    }
                }
                error_maps2(u, in[0], in[1], in[2], in[3], in[4], in[5], in[6], acc);
            };
            if (whole) {
#pragma unroll
                for (int g = 0; g < npairs; ++g) eval_pair(g, b * B + 2 * (first_pair + g), std::false_type{});
            } else {
#pragma unroll
                for (int g = 0; g < npairs; ++g) {
                    const int n = b * B + 2 * (first_pair + g);   // rows n, n + 1
                    if (n + 1 < h) eval_pair(g, n, std::false_type{});
                    else if (n < h) eval_pair(g, n, std::true_type{});
                }
            }
            if (DEC) warp_arrive(&sm.empty[b & 3]);   // ex[b & 1] and this batch's XYB third may be refilled
            if (b & 1) {   // binary32 over at most 4 pixels per accumulator (32 image rows), binary64 from there on
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    float lo, hi;
                    unpk2(acc[j], lo, hi);
                    dacc[j] += (double)lo;
                    dacc[j] += (double)hi;
                    acc[j] = zero;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            float lo, hi;
            unpk2(acc[j], lo, hi);
            dacc[j] += (double)lo;
            dacc[j] += (double)hi;
        }
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

typedef IirColsSmem<64, 16> IirColsDeep;

// Descriptors of the rows pass for one geometry.  Loads: XYB planes as {w, h, 3 channels, images}, box 44 x 32
// (out-of-bounds -> zeros).  Stores: the interleaved pair planes as {2w, h, 3, images} and the a*b planes as
// {w, h, 3, images}, box 32 floats x 32 rows, 128-byte swizzle (out-of-bounds -> clipped).
// The source's descriptors (one set per source buffer set) and the candidates' are built separately.
inline bool iir_rows_tma_maps_src(CUtensorMap in_src[kMaxScales], CUtensorMap out_psrc[kMaxScales], const Geom &g,
                                  const float *src, float *hpair_src)
{
    bool ok = true;
    for (int s = 0; s < g.n_scales && ok; ++s) {
        const uint64_t w = (uint64_t)g.w[s], h = (uint64_t)g.h[s], rowb = (uint64_t)g.pitch[s] * 4, planeb = (uint64_t)g.plane[s] * 4;
        ok = ok && tma_make_4d(&in_src[s], src + g.off[s], w, h, 3, 1, rowb, planeb, 0, kRtTileW, kIirRows, false);
        ok = ok && tma_make_4d(&out_psrc[s], hpair_src + 2 * g.off[s], 2 * w, h, 3, 1, 2 * rowb, 2 * planeb, 0, 32,
                               kIirRows, true);
    }
    return ok;
}

inline bool iir_cols_tma_maps_src(CUtensorMap psrc[kMaxScales], CUtensorMap xa[kMaxScales], const Geom &g, const float *hpair_src,
                                  const float *src)
{
    bool ok = true;
    for (int s = 0; s < g.n_scales && ok; ++s) {
        ok = tma_make_4d(&psrc[s], hpair_src + 2 * g.off[s], 2ull * g.w[s], (uint64_t)g.h[s], 3, 1, (uint64_t)g.pitch[s] * 8,
                         (uint64_t)g.plane[s] * 8, 0, 2 * kIirVCols, 4, false);
        ok = ok && tma_make_4d(&xa[s], src + g.off[s], (uint64_t)g.w[s], (uint64_t)g.h[s], 3, 1, (uint64_t)g.pitch[s] * 4,
                               (uint64_t)g.plane[s] * 4, 0, kIirVCols, 16, false);
    }
    return ok;
}

// Timing experiment (debug_time_rows | 16384, | 32768): descriptors that walk the SAME buffers strip by strip — a
// strip's rows adjacent in memory, as a strip-major layout of the row-filtered planes would have them — or that ask
// for 256-byte L2 promotion on the pair planes.  Strip-major descriptors read the right number of bytes from the
// wrong places: durations only, the sums are garbage.
inline bool iir_cols_tma_maps_experiment(IirColsTmaMaps *m, const Geom &g, const float *hpair_src, const float *src,
                                         const float *hpair_cand, const float *hab, const float *dist, bool strip_major,
                                         bool promote256)
{
    bool ok = true;
    for (int s = 0; s < g.n_scales && ok; ++s) {
        const uint64_t w = (uint64_t)g.w[s], h = (uint64_t)g.h[s], rowb = (uint64_t)g.pitch[s] * 4, planeb = (uint64_t)g.plane[s] * 4;
        const uint64_t strips = (w + kIirVCols - 1) / kIirVCols;
        if (strip_major) {
            ok = ok && tma_make_4d(&m->psrc[s], hpair_src + 2 * g.off[s], 2 * kIirVCols, h, strips, 3, 8 * kIirVCols, h * 8 * kIirVCols, 2 * planeb, 2 * kIirVCols, 4, false, promote256);
            ok = ok && tma_make_4d(&m->pcand[s], hpair_cand + 2 * g.off[s], 2 * kIirVCols, h, strips, 3, 8 * kIirVCols, h * 8 * kIirVCols, 2 * planeb, 2 * kIirVCols, 4, false, promote256);
            ok = ok && tma_make_4d(&m->ab[s], hab + g.off[s], kIirVCols, h, strips, 3, 4 * kIirVCols, h * 4 * kIirVCols, planeb, kIirVCols, 4, false);
            ok = ok && tma_make_4d(&m->xa[s], src + g.off[s], kIirVCols, h, strips, 3, 4 * kIirVCols, h * 4 * kIirVCols, planeb, kIirVCols, 16, false);
            ok = ok && tma_make_4d(&m->xb[s], dist + g.off[s], kIirVCols, h, strips, 3, 4 * kIirVCols, h * 4 * kIirVCols, planeb, kIirVCols, 16, false);
        } else {
            ok = ok && tma_make_4d(&m->psrc[s], hpair_src + 2 * g.off[s], 2 * w, h, 3, 1, 2 * rowb, 2 * planeb, 0, 2 * kIirVCols, 4, false, promote256);
            ok = ok && tma_make_4d(&m->pcand[s], hpair_cand + 2 * g.off[s], 2 * w, h, 3, 1, 2 * rowb, 2 * planeb, 0, 2 * kIirVCols, 4, false, promote256);
            ok = ok && tma_make_4d(&m->ab[s], hab + g.off[s], w, h, 3, 1, rowb, planeb, 0, kIirVCols, 4, false);
            ok = ok && tma_make_4d(&m->xa[s], src + g.off[s], w, h, 3, 1, rowb, planeb, 0, kIirVCols, 16, false);
            ok = ok && tma_make_4d(&m->xb[s], dist + g.off[s], w, h, 3, 1, rowb, planeb, 0, kIirVCols, 16, false);
        }
    }
    return ok;
}

inline bool iir_cols_tma_maps_cand(IirColsTmaMaps *m, const Geom &g, const float *hpair_cand, const float *hab,
                                   long long hcand_stride, const float *dist, long long dist_stride, int n_images)
{
    bool ok = true;
    for (int s = 0; s < g.n_scales && ok; ++s) {
        const uint64_t w = (uint64_t)g.w[s], h = (uint64_t)g.h[s], rowb = (uint64_t)g.pitch[s] * 4, planeb = (uint64_t)g.plane[s] * 4;
        ok = ok && tma_make_4d(&m->pcand[s], hpair_cand + 2 * g.off[s], 2 * w, h, 3, (uint64_t)n_images, 2 * rowb, 2 * planeb,
                               (uint64_t)hcand_stride * 4, 2 * kIirVCols, 4, false);
        ok = ok && tma_make_4d(&m->ab[s], hab + g.off[s], w, h, 3, (uint64_t)n_images, rowb, planeb,
                               (uint64_t)hcand_stride * 4, kIirVCols, 4, false);
        ok = ok && tma_make_4d(&m->xb[s], dist + g.off[s], w, h, 3, (uint64_t)n_images, rowb, planeb,
                               (uint64_t)dist_stride * 4, kIirVCols, 16, false);
    }
    return ok;
}

inline bool iir_rows_tma_maps_cand(IirRowsTmaMaps *m, const Geom &g, const float *dist, long long pyr_stride,
                                   float *hpair_cand, float *hab, long long hcand_stride, int n_images)
{
    bool ok = true;
    for (int s = 0; s < g.n_scales && ok; ++s) {
        const uint64_t w = (uint64_t)g.w[s], h = (uint64_t)g.h[s], rowb = (uint64_t)g.pitch[s] * 4, planeb = (uint64_t)g.plane[s] * 4;
        ok = ok && tma_make_4d(&m->in_dist[s], dist + g.off[s], w, h, 3, (uint64_t)n_images, rowb, planeb,
                               (uint64_t)pyr_stride * 4, kRtTileW, kIirRows, false);
        ok = ok && tma_make_4d(&m->out_pcand[s], hpair_cand + 2 * g.off[s], 2 * w, h, 3, (uint64_t)n_images, 2 * rowb,
                               2 * planeb, (uint64_t)hcand_stride * 4, 32, kIirRows, true);
        ok = ok && tma_make_4d(&m->out_ab[s], hab + g.off[s], w, h, 3, (uint64_t)n_images, rowb, planeb,
                               (uint64_t)hcand_stride * 4, 32, kIirRows, true);
    }
    return ok;
}

inline cudaError_t iir_configure()
{

    cudaError_t e = cudaFuncSetAttribute(k_iir_cols<64, 16, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(IirColsDeep));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_iir_cols<64, 16, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(IirColsDeep));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_iir_cols<64, 16, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(IirColsDeep) + 40 * 1024);   // head room for the occupancy experiment (pad_smem)
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_iir_cols<64, 16, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(IirColsDeep));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_iir_cols<64, 16, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(IirColsDeep));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_iir_cols<64, 16, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(IirColsDeep));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_iir_cols<64, 16, false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(IirColsDeep));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_iir_cols<64, 16, true, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(IirColsDeep));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_iir_rows<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IirRowsSmem<1>));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_iir_rows<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(IirRowsSmem<2>));
}

struct IirBuffers {
    float *src_hplanes;      // [2 * pyramid]: rows pass of (a, a*a), interleaved, cached per source
    float *cand_hplanes;     // [candidate][3 * pyramid]: rows pass of (b, b*b) interleaved, then a*b
    long long pyr_stride;    // floats per pyramid (capacity)
};

inline void iir_fill_common(IirArgs &a, const Geom &g, const IirCoef &k, const float *src, const float *dist,
                            long long dist_stride, const IirBuffers &B)
{
    a.g = g;
    a.k = k;
    a.one = 1.0f;
    a.neg_one = -1.0f;
    a.src = src;
    a.dist = dist;
    a.dist_stride = dist_stride;
    const long long P = B.pyr_stride;
    a.hpair_src = B.src_hplanes;           // (a, a*a)
    a.hpair_cand = B.cand_hplanes;         // (b, b*b)
    a.hab = B.cand_hplanes + 2 * P;        // a*b
    a.hcand_stride = 3 * P;
}

inline int iir_rows_grid(IirArgs &a, const Geom &g)
{
    int n = 0;
    for (int s = 0; s < g.n_scales; ++s) {
        const int nrb = (g.h[s] + kIirRows - 1) / kIirRows;
        a.blocks[s] = nrb;
        a.first_cta[s] = n;
        n += 3 * nrb;
    }
    for (int s = g.n_scales; s <= kMaxScales; ++s) a.first_cta[s] = n;
    return n;
}

struct IirDebugTap {
    float *out;      // device, 5 * w_s * h_s floats
    int scale, channel, cand;
};

// Rows pass.  which: 1 = the candidates' half (b, b*b, a*b), 3 = the source's half (a, a*a) alone, 2 = both in
// one launch.  maps != nullptr selects the TMA kernels (rows_shape: the instance, 0 = the shipped one), else the
// cp.async ones (which have no source-only form: 3 is not accepted there).
inline cudaError_t launch_iir_rows(const IirArgs &base, const Geom &g, int which, int n, cudaStream_t st,
                                   const IirRowsTmaMaps *maps, int rows_shape = 0)
{
    IirArgs ar = base;
    const int nr = iir_rows_grid(ar, g);
    if (maps) {
        switch (which) {
        case 1: return iir_rows_tma_dispatch<1>(rows_shape, ar, *maps, nr, n, st);
        case 2: return iir_rows_tma_dispatch<2>(rows_shape, ar, *maps, nr, n, st);
        default: return iir_rows_tma_dispatch<3>(rows_shape, ar, *maps, nr, 1, st);
        }
    }
    if (which == 2)
        k_iir_rows<2><<<dim3(nr, n), 192, sizeof(IirRowsSmem<2>), st>>>(ar);
    else if (which == 1)
        k_iir_rows<1><<<dim3(nr, n), 128, sizeof(IirRowsSmem<1>), st>>>(ar);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

// Columns pass with the maps and the pooling.
inline cudaError_t launch_iir_cols(const IirArgs &base, const int *first_cta_cols, const int *col_blocks, int n,
                                   cudaStream_t st, const IirDebugTap *tap = nullptr, const IirColsTmaMaps *maps = nullptr,
                                   bool decoupled = false, size_t pad_smem = 0, bool vorder = false)
{
    static const IirColsTmaMaps no_maps{};
    IirArgs a = base;
    for (int s = 0; s <= kMaxScales; ++s) a.first_cta[s] = first_cta_cols[s];
    for (int s = 0; s < kMaxScales; ++s) a.blocks[s] = col_blocks[s];
    const int ctas = first_cta_cols[kMaxScales];
    if (tap) {
        a.dbg_cols = tap->out;
        a.dbg_scale = tap->scale;
        a.dbg_channel = tap->channel;
        a.dbg_cand = tap->cand;
        if (vorder) {   // only the default tile path carries this instance (the API refuses the others)
            if (!maps || decoupled) return cudaErrorNotSupported;
            k_iir_cols<64, 16, true, true, false, true><<<dim3(ctas, n), kIirVThreads, sizeof(IirColsDeep), st>>>(a, *maps);
        } else if (maps && decoupled) k_iir_cols<64, 16, true, true, true><<<dim3(ctas, n), kIirVThreads, sizeof(IirColsDeep), st>>>(a, *maps);
        else if (maps) k_iir_cols<64, 16, true, true><<<dim3(ctas, n), kIirVThreads, sizeof(IirColsDeep), st>>>(a, *maps);
        else k_iir_cols<64, 16, true, false><<<dim3(ctas, n), kIirVThreads, sizeof(IirColsDeep), st>>>(a, no_maps);
    } else {
        // pad_smem (profiling only): a larger shared-memory request lowers the CTAs resident per SM
        if (vorder) {
            if (!maps || decoupled) return cudaErrorNotSupported;
            k_iir_cols<64, 16, false, true, false, true><<<dim3(ctas, n), kIirVThreads, sizeof(IirColsDeep), st>>>(a, *maps);
        } else if (maps && decoupled) k_iir_cols<64, 16, false, true, true><<<dim3(ctas, n), kIirVThreads, sizeof(IirColsDeep), st>>>(a, *maps);
        else if (maps) k_iir_cols<64, 16, false, true><<<dim3(ctas, n), kIirVThreads, sizeof(IirColsDeep) + pad_smem, st>>>(a, *maps);
        else k_iir_cols<64, 16, false, false><<<dim3(ctas, n), kIirVThreads, sizeof(IirColsDeep), st>>>(a, no_maps);
    }
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
