// ssimu2_wave.cuh — K4+K5 in ONE launch: the recursive Gaussian's rows pass, its columns pass, the SSIM / edge-diff maps
// and the pooling, with the row-filtered planes never leaving the SM.
//
// Why: the two-pass form writes 20 B and reads 28 B of row-filtered intermediates per scale pixel and channel, and
// both of its kernels sit at the ~4.7-5 TB/s their read/write mix gets out of HBM (profiles/r2_rows_ab.txt: the rows
// pass takes 93-97 % of its time with the arithmetic REMOVED).  The only way under that floor is not to move the
// bytes.  The published filter is one serial binary32 chain per row and per column, so a tile cannot be filtered
// on its own — but the state of a chain is six floats, and that is all that has to cross a tile boundary:
//
//   * one CTA owns a 32-column strip of one channel of one scale and walks it top to bottom (the columns pass's
//     decomposition: the column recursions of its 32 lanes live in registers for the whole walk);
//   * its ROW recursion warps (lane = row) filter 32-row x 32-column chunks of the strip; the state of the 32 row
//     chains at the strip's left edge comes from the CTA of the strip to the left, and the state at the right
//     edge goes to the CTA to the right — six floats per chain and quantity through an L2-resident mailbox, each
//     value carrying its own tag in the same 8-byte word (no fence, no separate flag: one L2 write + one L2 read
//     per hop).  The strips of a chain therefore run as a wavefront, each one chunk behind its left neighbour;
//   * the filtered chunk goes into a shared-memory ring transposed for free (lane = row writes with a padded
//     pitch, lane = column reads), where the column recursion warps pick it up, and the map warps pool the result
//     one batch later — exactly the "rows via an in-smem transpose, columns in parallel" shape north_star names;
//   * the a / b samples arrive by TMA (cp.async.bulk.tensor, 44-column boxes whose out-of-bounds part is the
//     filter's zero padding), completion on mbarriers.
//
// Deadlock freedom: a CTA only ever waits for the strip to its LEFT.  Units are handed out by an atomic ticket in
// strip-major order, so whatever a resident CTA waits for holds an earlier ticket and is resident or finished.
// A spin limit turns a protocol bug into an error code instead of a hung device.
//
// Arithmetic: the same operations in the same order as k_iir_rows / k_iir_cols (ssimu2_iir.cuh) — state hand-off
// copies floats, nothing is re-associated — so the tests demand the same bits.
//
// HBM traffic per scale pixel and channel: 8 B read (a, b) + 8 B of cached source blur written once per source
// (MODE 2) or read per candidate (MODE 1), against 48-60 B for the two-pass form.
#pragma once

#include <algorithm>
#include <vector>

#include "ssimu2_iir.cuh"

namespace oavif {

constexpr int kWvB = 16;        // rows per phase: one column batch
constexpr int kWvNT = 4;        // a / b tile ring, in 32-row chunks
constexpr int kWvMS = 4;        // cached (mu1, sigma11) ring, in 16-row blocks
constexpr int kWvRing = 64;     // rows of the row-filtered rings (two chunks)
constexpr unsigned kWvSpinLimit = 1u << 24;

struct WaveMaps {
    CUtensorMap in_src[kMaxScales], in_dist[kMaxScales];   // XYB planes {w, h, 3, images}, box 44 x 32
    CUtensorMap in_musig[kMaxScales];                       // cached (mu1, sigma11) pairs {2w, h, 3, 1}, box 64 x 16
};

struct WaveArgs {
    Geom g;
    IirCoef k;
    float one, neg_one;
    const unsigned *units;          // ticket / n_cand -> unit: scale | channel << 4 | strip << 8, strip-major order
    unsigned n_units, n_cand, cand0;
    unsigned *ticket;               // zero before the launch; the CTA that draws the last ticket resets it
    unsigned epoch;                 // distinguishes this launch's mailbox tags from older ones
    unsigned long long *mailbox;    // {value, tag} words; see wave_mailbox_words()
    long long mb_scale_off[kMaxScales], mb_cand_stride;
    int mb_rows[kMaxScales];        // rows per (chain, parity) region of a scale: h rounded up to 32
    float *musig;                   // interleaved (mu1, sigma11) planes of the source, pyramid layout doubled
    double *partials;
    long long partials_stride;
    int first_cta[kMaxScales + 1], blocks[kMaxScales];   // the columns plan: where a unit's six sums go
    int *error_flag;                // set when a hand-off wait gives up
    unsigned long long *dbg_trace;  // TAP instances: per ticket {unit, t0, t after the prologue, t at phase nphase/2, t end} (ns)
    float *dbg_cols;                // oavif_ssimu2_debug_get_cols (TAP instances only)
    int dbg_scale, dbg_channel, dbg_cand;
};

// MODE 2: all five quantities (first call after set_source; also leaves (mu1, sigma11) in the cache).
// MODE 1: the candidate's three — (b, b*b), a*b — with (mu1, sigma11) read back from the cache.
template <int MODE>
struct WaveCfg {
    static constexpr int NP = MODE == 2 ? 2 : 1;              // packed pair recursions
    static constexpr int Q = MODE == 2 ? 5 : 3;               // scalar chains per row in the mailbox
    static constexpr int W_COL0 = 0;                          // column producers: NP pair warps, then a*b
    static constexpr int W_CONS0 = NP + 1;                    // four map warps
    static constexpr int W_ROW0 = NP + 5;                     // row recursion: NP pair warps, then a*b
    static constexpr int W_LOAD = 2 * NP + 6;
    static constexpr int WARPS = 2 * NP + 7;                  // 11 / 9
    static constexpr int THREADS = 32 * WARPS;
};

template <int MODE>
struct WaveSmem {
    typedef WaveCfg<MODE> C;
    float tile[2][kWvNT][kIirRows][kRtTileW];        // [plane: 0 = b, 1 = a][chunk & 3][row][column - (32 t - 8)]
    float pring[C::NP][kWvRing][kIirPairPitch];      // row-filtered pairs, image row r at slot r & 63, (x, x*x) per pixel
    float sring[kWvRing][kIirPitch];                 // row-filtered a*b
    float ex[2][5][kWvB][kIirVCols];                 // column outputs of a phase: mu1, mu2, s11, s22, s12
    float ms[MODE == 1 ? kWvMS : 1][kWvB][2 * kIirVCols];   // MODE 1: cached (mu1, sigma11) blocks
    double red[4][6];
    uint64_t full_tile[kWvNT], full_ms[kWvMS];
    unsigned ticket;
};

inline long long wave_mailbox_words(const Geom &g, int q, long long *scale_off, int *rows)
{
    long long off = 0;
    for (int s = 0; s < kMaxScales; ++s) {
        scale_off[s] = off;
        rows[s] = s < g.n_scales ? ((g.h[s] + 31) / 32) * 32 : 0;
        off += 3LL * 2 * rows[s] * q * 6;   // channels x strip parity x rows x chains x six state words
    }
    return off;
}

__device__ __forceinline__ void st_volatile_v4(unsigned long long *p, unsigned a, unsigned b, unsigned c, unsigned d)
{
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ld_volatile_v4(const unsigned long long *p, unsigned &a, unsigned &b, unsigned &c, unsigned &d)
{
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
}

// N state floats of one row chain group: publish to / fetch from the mailbox.  Every 8-byte word is {value, tag}.
template <int N>
__device__ __forceinline__ void mailbox_put(unsigned long long *p, const float *f, unsigned tag)
{
#pragma unroll
    for (int i = 0; i < N / 2; ++i) st_volatile_v4(p + 2 * i, __float_as_uint(f[2 * i]), tag, __float_as_uint(f[2 * i + 1]), tag);
}
template <int N>
__device__ __forceinline__ bool mailbox_get(const unsigned long long *p, float *f, unsigned tag, bool active)
{
    unsigned spins = 0;
    bool ok = !active;
    // every lane of the warp keeps polling until ALL lanes have their words (the warp moves on together)
    while (true) {
        if (!ok) {
            unsigned v[N], t[N];
#pragma unroll
            for (int i = 0; i < N / 2; ++i) ld_volatile_v4(p + 2 * i, v[2 * i], t[2 * i], v[2 * i + 1], t[2 * i + 1]);
            bool all = true;
#pragma unroll
            for (int i = 0; i < N; ++i) all = all && (t[i] == tag);
            if (all) {
#pragma unroll
                for (int i = 0; i < N; ++i) f[i] = __uint_as_float(v[i]);
                ok = true;
            }
        }
        if (__all_sync(0xffffffffu, ok)) return true;
        if (++spins > kWvSpinLimit) return false;
    }
}

template <int MODE, bool TAP>
__global__ void __launch_bounds__(WaveCfg<MODE>::THREADS, 2)
    k_blur_wave(const __grid_constant__ WaveArgs a, const __grid_constant__ WaveMaps tm)
{
    typedef WaveCfg<MODE> C;
    extern __shared__ __align__(128) unsigned char smem_wave[];
    WaveSmem<MODE> &sm = *reinterpret_cast<WaveSmem<MODE> *>(smem_wave);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- which unit: tickets are drawn in launch order, units are listed strip-major -------------------------
    if (threadIdx.x == 0) {
        const unsigned tk = atomicAdd(a.ticket, 1u);
        if (tk == a.n_units * a.n_cand - 1) *a.ticket = 0u;   // every ticket is out: ready for the next launch
        sm.ticket = tk;
#pragma unroll
        for (int i = 0; i < kWvNT; ++i) mbar_init(&sm.full_tile[i], 1);
#pragma unroll
        for (int i = 0; i < kWvMS; ++i) mbar_init(&sm.full_ms[i], 1);
        mbar_init_fence();
    }
    __syncthreads();
    unsigned long long t_begin = 0;
    if (TAP && a.dbg_trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
    const unsigned unit = a.units[sm.ticket / a.n_cand];
    const int cand = (int)(a.cand0 + sm.ticket % a.n_cand);
    const int s = (int)(unit & 15u), c = (int)((unit >> 4) & 15u), t = (int)(unit >> 8);
    const int w = a.g.w[s], h = a.g.h[s], pitch = a.g.pitch[s];
    const int nstrips = (w + kIirVCols - 1) / kIirVCols;
    const int x0 = t * kIirVCols;
    const int valid_cols = min(kIirVCols, w - x0);
    const int ncolph = (h + 4 + kWvB - 1) / kWvB;          // column steps n = -4 .. h-1 in batches of 16
    const int nchunk = (ncolph * kWvB + kIirRows - 1) / kIirRows;   // row chunks the column pass reads (the last may lie below h: zeros)
    const int nphase = ncolph + 1;                         // the maps run one phase behind the column recursion
    const long long poff = a.g.off[s] + (long long)c * a.g.plane[s];

    // mailbox of this chain: [parity of the writing strip][row][chain][6 words]
    const long long mb_region = (long long)a.mb_rows[s] * C::Q * 6;
    unsigned long long *mb = a.mailbox + (long long)(cand - (int)a.cand0) * a.mb_cand_stride + a.mb_scale_off[s] +
                             (long long)c * 2 * mb_region;
    const unsigned tag_in = (a.epoch << 12) | (unsigned)t, tag_out = (a.epoch << 12) | (unsigned)(t + 1);   // tags are strip + 1

    if (warp == C::W_LOAD) {
        // ---------------- loader: one lane feeds the tile ring (and, MODE 1, the cached source blur) ----------------
        if (lane == 0) {
            tma_prefetch_desc(&tm.in_dist[s]);
            tma_prefetch_desc(&tm.in_src[s]);
            if (MODE == 1) tma_prefetch_desc(&tm.in_musig[s]);
        }
        auto issue_tile = [&](int m) {
            if (lane != 0 || m >= nchunk) return;
            const int st = m & (kWvNT - 1);
            mbar_arrive_expect_tx(&sm.full_tile[st], 2 * kRtTileBytes);
            tma_load_4d(&sm.tile[0][st][0][0], &tm.in_dist[s], x0 - 8, m * kIirRows, c, cand, &sm.full_tile[st]);
            tma_load_4d(&sm.tile[1][st][0][0], &tm.in_src[s], x0 - 8, m * kIirRows, c, 0, &sm.full_tile[st]);
        };
        auto issue_ms = [&](int j) {   // rows 16 j .. 16 j + 15 of the cached (mu1, sigma11) pairs
            if (MODE != 1 || lane != 0 || j * kWvB >= h) return;
            const int st = j & (kWvMS - 1);
            mbar_arrive_expect_tx(&sm.full_ms[st], kWvB * 2 * kIirVCols * 4);
            tma_load_4d(&sm.ms[MODE == 1 ? st : 0][0][0], &tm.in_musig[s], 2 * x0, j * kWvB, c, 0, &sm.full_ms[st]);
        };
        issue_tile(0);
        issue_tile(1);
        issue_ms(0);
        __syncthreads();   // (P) chunk 0 is row-filtered
        unsigned long long t_pro = 0, t_mid = 0, t_end = 0;
        if (TAP && a.dbg_trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_pro));
#pragma unroll 1
        for (int q = 0; q < nphase; ++q) {
            if (!(q & 1)) issue_tile(q / 2 + 2);   // its slot held chunk q/2 - 2, dead since the barrier that ended phase q - 1
            issue_ms(q + 1);
            __syncthreads();
            if (TAP && a.dbg_trace && q == nphase / 2) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_mid));
        }
        __syncthreads();   // final reduction
        if (TAP && a.dbg_trace && lane == 0) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            unsigned long long *o = a.dbg_trace + 5ull * sm.ticket;
            o[0] = unit; o[1] = t_begin; o[2] = t_pro; o[3] = t_mid; o[4] = t_end;
        }
        return;
    }

    if (warp >= C::W_ROW0) {
        // ---------------- row recursion warps: lane = row, a 32 x 32 chunk per two phases ----------------
        const int rw = warp - C::W_ROW0;                 // 0 .. NP-1: packed pairs, NP: a*b
        const bool is_pair = rw < C::NP;
        // MODE 2: pair 0 = (a, a*a) from plane 1, pair 1 = (b, b*b) from plane 0; MODE 1: pair 0 = (b, b*b)
        const int plane = (MODE == 2 && rw == 0) ? 1 : 0;
        const int chain0 = is_pair ? 2 * rw : 2 * C::NP;  // first mailbox chain of this warp
        const IirCoef2 k2 = iir_coef2(a.k, a.one, a.neg_one);
        IirState2 st2;
        IirState st1;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            st2.p[i] = st2.p2[i] = splat2(0.0f);
            st1.p[i] = st1.p2[i] = 0.0f;
        }
        bool alive = true;

        // one half (16 steps) of chunk m: columns 16 half .. 16 half + 15 of the strip
        auto run_half = [&](int m, int half) {
            const int stg = m & (kWvNT - 1);
            const int row = m * kIirRows + lane;
            if (half == 0) {
                mbar_wait(&sm.full_tile[stg], (unsigned)(m / kWvNT) & 1u);
                if (t > 0) {   // the chains' state at the strip's left edge, from the strip to the left
                    const unsigned long long *src = mb + (long long)((t - 1) & 1) * mb_region + ((long long)row * C::Q + chain0) * 6;
                    const bool active = row < h;
                    if (is_pair) {
                        float f[12] = {};
                        if (!mailbox_get<12>(src, f, tag_in, active)) alive = false;
#pragma unroll
                        for (int i = 0; i < 3; ++i) {   // rows below the image: zero state over zero samples = zeros
                            st2.p[i] = active ? pk2(f[i], f[6 + i]) : splat2(0.0f);
                            st2.p2[i] = active ? pk2(f[3 + i], f[9 + i]) : splat2(0.0f);
                        }
                    } else {
                        float f[6] = {};
                        if (!mailbox_get<6>(src, f, tag_in, active)) alive = false;
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            st1.p[i] = active ? f[i] : 0.0f;
                            st1.p2[i] = active ? f[3 + i] : 0.0f;
                        }
                    }
                } else {       // the chains start here: zero state, then n = -4 .. -1 (right taps are columns 0..3)
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        st2.p[i] = st2.p2[i] = splat2(0.0f);
                        st1.p[i] = st1.p2[i] = 0.0f;
                    }
                }
            }
            const int slot = row & (kWvRing - 1);
            // samples at columns x0 - 8 + i; this half needs i = 16 half + 2 .. 16 half + 27 (left tap j + 2, right tap j + 12)
            const int base = 16 * half;
            if (is_pair) {
                const float *src = &sm.tile[plane][stg][lane][base];
                float v[28], qq[28];
#pragma unroll
                for (int j4 = 0; j4 < 7; ++j4) {
                    const float4 x = *reinterpret_cast<const float4 *>(src + 4 * j4);
                    v[4 * j4] = x.x; v[4 * j4 + 1] = x.y; v[4 * j4 + 2] = x.z; v[4 * j4 + 3] = x.w;
                    unpk2(mul2(pk2(x.x, x.y), pk2(x.x, x.y)), qq[4 * j4], qq[4 * j4 + 1]);
                    unpk2(mul2(pk2(x.z, x.w), pk2(x.z, x.w)), qq[4 * j4 + 2], qq[4 * j4 + 3]);
                }
                if (half == 0 && t == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) (void)iir_step2(k2, st2, splat2(0.0f), pk2(v[8 + i], qq[8 + i]));
                }
                float *dst = &sm.pring[rw][slot][2 * base];
                IirPipe2 P;
                pipe2_begin(k2, P, st2, pk2(v[2] + v[12], qq[2] + qq[12]));
#pragma unroll
                for (int j2 = 0; j2 < 8; ++j2) {
                    f32x2 o[2];
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const int j = 2 * j2 + jj;
                        o[jj] = (j + 1 < 16) ? pipe2_step(k2, P, pk2(v[j + 3] + v[j + 13], qq[j + 3] + qq[j + 13]))
                                             : pipe2_end(k2, P, st2);
                    }
                    float4 ov;
                    unpk2(o[0], ov.x, ov.y);
                    unpk2(o[1], ov.z, ov.w);
                    if (valid_cols < kIirVCols) {   // last strip: columns right of the image feed the column pass zeros
                        const int j = base + 2 * j2;
                        if (j >= valid_cols) ov.x = ov.y = 0.0f;
                        if (j + 1 >= valid_cols) ov.z = ov.w = 0.0f;
                    }
                    *reinterpret_cast<float4 *>(dst + 4 * j2) = ov;
                }
            } else {
                const float *sb = &sm.tile[0][stg][lane][base], *sa = &sm.tile[1][stg][lane][base];
                float v[28];
#pragma unroll
                for (int j4 = 0; j4 < 7; ++j4) {
                    const float4 x = *reinterpret_cast<const float4 *>(sb + 4 * j4);
                    const float4 y = *reinterpret_cast<const float4 *>(sa + 4 * j4);
                    unpk2(mul2(pk2(x.x, x.y), pk2(y.x, y.y)), v[4 * j4], v[4 * j4 + 1]);
                    unpk2(mul2(pk2(x.z, x.w), pk2(y.z, y.w)), v[4 * j4 + 2], v[4 * j4 + 3]);
                }
                if (half == 0 && t == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) (void)iir_step(a.k, st1, 0.0f, v[8 + i]);
                }
                float sum[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) sum[j] = v[j + 2] + v[j + 12];
                float *dst = &sm.sring[slot][base];
                IirPipe P;
                pipe_begin(a.k, P, st1, sum[0]);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    float o[4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int j = 4 * j4 + jj;
                        o[jj] = (j + 1 < 16) ? pipe_step(a.k, P, sum[j + 1]) : pipe_end(a.k, P, st1);
                        if (valid_cols < kIirVCols && base + j >= valid_cols) o[jj] = 0.0f;
                    }
                    *reinterpret_cast<float4 *>(dst + 4 * j4) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
            if (half == 1 && t + 1 < nstrips && row < h) {   // the state at the right edge, for the strip to the right
                unsigned long long *dst = mb + (long long)(t & 1) * mb_region + ((long long)row * C::Q + chain0) * 6;
                if (is_pair) {
                    float f[12];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        unpk2(st2.p[i], f[i], f[6 + i]);
                        unpk2(st2.p2[i], f[3 + i], f[9 + i]);
                    }
                    mailbox_put<12>(dst, f, tag_out);
                } else {
                    float f[6];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        f[i] = st1.p[i];
                        f[3 + i] = st1.p2[i];
                    }
                    mailbox_put<6>(dst, f, tag_out);
                }
            }
        };

        run_half(0, 0);
        run_half(0, 1);
        __syncthreads();   // (P)
#pragma unroll 1
        for (int q = 0; q < nphase; ++q) {
            const int m = q / 2 + 1;
            if (m < nchunk) run_half(m, q & 1);
            __syncthreads();
        }
        if (!alive && lane == 0) atomicExch(a.error_flag, 1);
        __syncthreads();   // final reduction
        return;
    }

    if (warp < C::W_CONS0) {
        // ---------------- column recursion warps: lane = column, 16 steps per phase ----------------
        // Phase q runs steps n = 16 q - 4 .. 16 q + 11: right taps are rows 16 q .. 16 q + 15 of the ring (filtered one
        // phase earlier), left taps (n - 6) come out of a ten-deep register delay line — nothing older than the current
        // chunk is read from shared memory, so the row warps may overwrite the other half of the ring meanwhile.
        const int cw = warp - C::W_COL0;
        const bool is_pair = cw < C::NP;
        __syncthreads();   // (P)
        if (is_pair) {
            const IirCoef2 k2 = iir_coef2(a.k, a.one, a.neg_one);
            IirState2 st;
            f32x2 D[10];
#pragma unroll
            for (int i = 0; i < 3; ++i) st.p[i] = st.p2[i] = splat2(0.0f);
#pragma unroll
            for (int i = 0; i < 10; ++i) D[i] = splat2(0.0f);
            // outputs: MODE 2: pair 0 -> mu1 (ex 0), s11 (ex 2); pair 1 -> mu2 (ex 1), s22 (ex 3).  MODE 1: pair 0 -> mu2, s22
            const int e_mu = (MODE == 2) ? cw : 1, e_sg = e_mu + 2;
            const float *col = &sm.pring[cw][0][2 * lane];
#pragma unroll 1
            for (int q = 0; q < nphase; ++q) {
                if (q < ncolph) {
                    const float *r0 = col + ((q * kWvB) & (kWvRing - 1)) * kIirPairPitch;
                    f32x2 R[kWvB];
#pragma unroll
                    for (int j = 0; j < kWvB; ++j) R[j] = lds2(r0 + j * kIirPairPitch);
                    float *ex0 = &sm.ex[q & 1][e_mu][0][lane], *ex1 = &sm.ex[q & 1][e_sg][0][lane];
                    IirPipe2 P;
                    pipe2_begin(k2, P, st, add2(D[0], R[0]));
#pragma unroll
                    for (int j = 0; j < kWvB; ++j) {
                        const f32x2 o = (j + 1 < kWvB) ? pipe2_step(k2, P, add2(j + 1 < 10 ? D[j + 1] : R[j + 1 - 10], R[j + 1]))
                                                       : pipe2_end(k2, P, st);
                        unpk2(o, ex0[j * kIirVCols], ex1[j * kIirVCols]);
                    }
#pragma unroll
                    for (int i = 0; i < 10; ++i) D[i] = R[6 + i];
                }
                __syncthreads();
            }
        } else {
            IirState st;
            float D[10];
#pragma unroll
            for (int i = 0; i < 3; ++i) st.p[i] = st.p2[i] = 0.0f;
#pragma unroll
            for (int i = 0; i < 10; ++i) D[i] = 0.0f;
            const float *col = &sm.sring[0][lane];
#pragma unroll 1
            for (int q = 0; q < nphase; ++q) {
                if (q < ncolph) {
                    const float *r0 = col + ((q * kWvB) & (kWvRing - 1)) * kIirPitch;
                    float R[kWvB];
#pragma unroll
                    for (int j = 0; j < kWvB; ++j) R[j] = r0[j * kIirPitch];
                    float *ex = &sm.ex[q & 1][4][0][lane];
                    IirPipe P;
                    pipe_begin(a.k, P, st, D[0] + R[0]);
#pragma unroll
                    for (int j = 0; j < kWvB; ++j)
                        ex[j * kIirVCols] = (j + 1 < kWvB) ? pipe_step(a.k, P, (j + 1 < 10 ? D[j + 1] : R[j + 1 - 10]) + R[j + 1])
                                                           : pipe_end(a.k, P, st);
#pragma unroll
                    for (int i = 0; i < 10; ++i) D[i] = R[6 + i];
                }
                __syncthreads();
            }
        }
        __syncthreads();   // final reduction
        return;
    }

    // ---------------- map warps: SSIM / edge-diff maps and pooling, one phase behind ----------------
    // Phase q pools the rows the column warps emitted in phase q - 1: image rows 16 q - 20 .. 16 q - 5.  Each of the
    // four warps takes two row pairs; a pair of rows of a column is one packed evaluation (error_maps2).  A warp keeps
    // the IMAGE rows k_iir_cols gives it — rows 4 cw .. 4 cw + 3 of every aligned 16-row group — and moves its
    // binary32 accumulators to binary64 after the same rows, so the pooled sums carry the same bits as the two-pass form.
    {
        const int cw = warp - C::W_CONS0;                  // 0..3
        double dacc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        const Unit2 u = unit2(a.one, a.neg_one);
        const f32x2 zero = splat2(0.0f);
        f32x2 acc[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[j] = zero;
        const bool dbg = TAP && a.dbg_cols != nullptr && s == a.dbg_scale && c == a.dbg_channel && cand == a.dbg_cand;
        float *mus = (MODE == 2 && cand == 0) ? a.musig + 2 * poff + 2 * (x0 + lane) : nullptr;
        __syncthreads();   // (P)
#pragma unroll 1
        for (int q = 0; q < nphase; ++q) {
            if (q >= 1) {
                const int r0 = kWvB * q - 20;              // first image row of the batch
                if (MODE == 1) {                           // the cached source blur of rows r0 .. r0 + 15: two 16-row blocks
                    const int j1 = q - 1, j0 = q - 2;      // rows 16 q - 16 .. (first 12 used), 16 q - 32 .. (last 4 used)
                    if (j0 >= 0 && j0 * kWvB < h) mbar_wait(&sm.full_ms[j0 & (kWvMS - 1)], (unsigned)(j0 / kWvMS) & 1u);
                    if (j1 * kWvB < h) mbar_wait(&sm.full_ms[j1 & (kWvMS - 1)], (unsigned)(j1 / kWvMS) & 1u);
                }
                const float *ex = &sm.ex[(q - 1) & 1][0][0][lane];
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const int jl = ((4 * cw + 4) & 15) + 2 * g;   // local rows jl, jl + 1: image rows = 4 cw + 2 g (mod 16)
                    const int n = r0 + jl;
                    if (n >= 0 && n < h) {
                        const bool two = n + 1 < h;
                        // the pixel's own samples: tile of chunk n >> 5 (both rows of a pair lie in one chunk: n is even)
                        const float *ta = &sm.tile[1][(n >> 5) & (kWvNT - 1)][n & 31][8 + lane];
                        const float *tb = &sm.tile[0][(n >> 5) & (kWvNT - 1)][n & 31][8 + lane];
                        f32x2 in[7];
                        in[0] = pk2(ta[0], ta[kRtTileW]);
                        in[1] = pk2(tb[0], tb[kRtTileW]);
#pragma unroll
                        for (int qq = 0; qq < 5; ++qq)
                            in[2 + qq] = pk2(ex[(qq * kWvB + jl) * kIirVCols], ex[(qq * kWvB + jl + 1) * kIirVCols]);
                        if (MODE == 1) {                   // mu1 -> in[2], sigma11 -> in[4]
                            const float2 m0 = *reinterpret_cast<const float2 *>(&sm.ms[MODE == 1 ? (n >> 4) & (kWvMS - 1) : 0][n & 15][2 * lane]);
                            const float2 m1 = *reinterpret_cast<const float2 *>(&sm.ms[MODE == 1 ? ((n + 1) >> 4) & (kWvMS - 1) : 0][(n + 1) & 15][2 * lane]);
                            in[2] = pk2(m0.x, m1.x);
                            in[4] = pk2(m0.y, m1.y);
                        }
                        if (MODE == 2 && mus != nullptr) { // leave (mu1, sigma11) for the later candidates of this source
                            float lo, hi, slo, shi;
                            unpk2(in[2], lo, hi);
                            unpk2(in[4], slo, shi);
                            *reinterpret_cast<float2 *>(mus + (long long)n * 2 * pitch) = make_float2(lo, slo);
                            if (two) *reinterpret_cast<float2 *>(mus + (long long)(n + 1) * 2 * pitch) = make_float2(hi, shi);
                        }
                        if (TAP && dbg && lane < valid_cols) {
#pragma unroll
                            for (int qq = 0; qq < 5; ++qq) {
                                float lo, hi;
                                unpk2(in[2 + qq], lo, hi);
                                float *o = a.dbg_cols + ((long long)qq * h + n) * w + x0 + lane;
                                o[0] = lo;
                                if (two) o[w] = hi;
                            }
                        }
                        if (!two) {   // odd height: the pair's second row is outside
#pragma unroll
                            for (int i = 0; i < 7; ++i) {
                                float lo, hi;
                                unpk2(in[i], lo, hi);
                                in[i] = pk2(lo, 0.0f);
                            }
                        }
                        if (valid_cols < kIirVCols && lane >= valid_cols) {   // columns right of the image pool zero
#pragma unroll
                            for (int i = 0; i < 7; ++i) in[i] = zero;
                        }
                        error_maps2(u, in[0], in[1], in[2], in[3], in[4], in[5], in[6], acc);
                    }
                }
                // binary32 over at most 4 pixels per accumulator, binary64 from there on: after the rows of every odd
                // aligned 16-row group (warp 3's rows of a group arrive one phase later than the others')
                if (cw == 3 ? (q & 1) : !(q & 1)) {
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
            __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            float lo, hi;
            unpk2(acc[j], lo, hi);
            dacc[j] += (double)lo;
            dacc[j] += (double)hi;
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            double x = dacc[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            if (lane == 0) sm.red[cw][j] = x;
        }
        __syncthreads();   // final reduction
        if (cw == 0 && lane < 6) {
            const long long cta = (long long)a.first_cta[s] + (long long)c * a.blocks[s] + t;
            a.partials[(long long)cand * a.partials_stride + cta * 6 + lane] =
                ((sm.red[0][lane] + sm.red[1][lane]) + sm.red[2][lane]) + sm.red[3][lane];
        }
    }
}

// ---- host side ----------------------------------------------------------------------------------------------
// units in strip-major order: a strip's left neighbour always holds an earlier ticket
inline std::vector<unsigned> wave_units(const Geom &g)
{
    std::vector<unsigned> u;
    int max_strips = 0;
    for (int s = 0; s < g.n_scales; ++s) max_strips = std::max(max_strips, (g.w[s] + kIirVCols - 1) / kIirVCols);
    for (int t = 0; t < max_strips; ++t)
        for (int s = 0; s < g.n_scales; ++s)
            if (t < (g.w[s] + kIirVCols - 1) / kIirVCols)
                for (int c = 0; c < 3; ++c) u.push_back((unsigned)s | ((unsigned)c << 4) | ((unsigned)t << 8));
    return u;
}

inline bool wave_tma_maps(WaveMaps *m, const Geom &g, const float *src, const float *dist, long long pyr_stride,
                          const float *musig, int n_images)
{
    bool ok = true;
    for (int s = 0; s < g.n_scales && ok; ++s) {
        const uint64_t w = (uint64_t)g.w[s], h = (uint64_t)g.h[s], rowb = (uint64_t)g.pitch[s] * 4, planeb = (uint64_t)g.plane[s] * 4;
        ok = ok && tma_make_4d(&m->in_src[s], src + g.off[s], w, h, 3, 1, rowb, planeb, 0, kRtTileW, kIirRows, false);
        ok = ok && tma_make_4d(&m->in_dist[s], dist + g.off[s], w, h, 3, (uint64_t)n_images, rowb, planeb,
                               (uint64_t)pyr_stride * 4, kRtTileW, kIirRows, false);
        ok = ok && tma_make_4d(&m->in_musig[s], musig + 2 * g.off[s], 2 * w, h, 3, 1, 2 * rowb, 2 * planeb, 0,
                               2 * kIirVCols, kWvB, false);
    }
    return ok;
}

template <int MODE, bool TAP>
inline cudaError_t wave_launch(const WaveArgs &a, const WaveMaps &maps, cudaStream_t st)
{
    static bool configured[64] = {};   // per (instance, device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        const cudaError_t e = cudaFuncSetAttribute(k_blur_wave<MODE, TAP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)sizeof(WaveSmem<MODE>));
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    k_blur_wave<MODE, TAP><<<a.n_units * a.n_cand, WaveCfg<MODE>::THREADS, sizeof(WaveSmem<MODE>), st>>>(a, maps);
    return cudaGetLastError();
}

}  // namespace oavif
