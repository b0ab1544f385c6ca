// ssimu2_api.cu — the C ABI of include/oavif_ssimu2.h: context, HBM layout, launches.
//
// Product path only: no CPU fallback, nothing from oracle/.  Build (see oavif_b200/build.py):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
//
// HBM layout per context (sized once in ctx_create for max_w x max_h x max_batch):
//   in_src / in_dist[k]   staged input pixels (RGB8, or 3 YUV planes of u8/u16), tight rows
//   src_pyr               source XYB pyramid: scales 0..5, 3 f32 planes each (Geom)
//   dist_pyr[k]           one XYB pyramid per candidate
//   src_hplanes           RECURSIVE blur only: row-filtered a and a*a of the source (cached with the pyramid)
//   hplanes[k]            RECURSIVE blur only: row-filtered b, b*b, a*b of candidate k
//   partials[k][cta][6]   per-CTA pooled sums (double), reduced in fixed order by k_finalize
//   sums[k][6][18], scores[k]
#include "../../include/oavif_ssimu2.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "ssimu2_common.cuh"
#include "ssimu2_finalize.cuh"
#include "ssimu2_fir.cuh"
#include "ssimu2_iir.cuh"
#include "ssimu2_pyramid.cuh"
#include "ssimu2_wave.cuh"

using namespace oavif;


namespace {

thread_local std::string t_last_error;

struct HostPlanes {  // one candidate's input, host or device pointers
    const void *p[3];
};

// the stateless entry point's process-wide context (oavif_ssimu2_compute_rgb8)
std::mutex g_cached_mu;
oavif_ssimu2_ctx *g_cached = nullptr;
int g_default_device = 0;

}  // namespace

constexpr int kSlots = 2;   // submissions that may be in flight at once (oavif_ssimu2_submit_* / _wait)

// Everything on the candidate side of an evaluation.  Each of the two submission slots owns one, with a compute
// stream of its own: the kernels of two submissions in flight overlap (the rows pass of one fills the tails of the
// other's columns pass), which is what two callers on two contexts got in round 1.
struct CandSide {
    cudaStream_t own_stream = nullptr, stream = nullptr;   // stream: own, or the caller's (set_stream)
    float *d_dist_pyr = nullptr, *d_hplanes = nullptr;
    double *d_partials = nullptr;
    IirRowsTmaMaps cand_maps{};    // TMA descriptors of the candidates' planes (in_dist, out_pcand, out_ab) for maps_w x maps_h
    IirColsTmaMaps cols_maps{};    // ... and the columns pass's view of the row-filtered planes (pcand, ab)
    int maps_w = -1, maps_h = -1;
    // fused kernel (ssimu2_wave.cuh): unit list of the current geometry, ticket counter, mailbox
    unsigned *d_units = nullptr, *d_ticket = nullptr;
    unsigned n_units = 0;
    int units_w = -1, units_h = -1;
    unsigned long long *d_mailbox = nullptr;
    bool allocated = false;
};

// One submission: its staged candidates, its candidate-side buffers, its result buffers and its events.
struct Slot {
    CandSide cs;
    uint8_t *d_in_dist = nullptr;                       // staged candidate pixels (slot 1: allocated on first use)
    double *h_sums = nullptr, *h_scores = nullptr;      // pinned, mapped: k_finalize writes them directly
    double *dm_sums = nullptr, *dm_scores = nullptr;    // device views of the two
    cudaEvent_t up0 = nullptr, up1 = nullptr;           // copy stream: around the upload
    cudaEvent_t k[5] = {};                              // compute stream: pyramid start / end, between the passes, blur end, finalize end
    uint32_t n = 0;
    bool degenerate = false, host_input = false;        // degenerate: below 8x8, nothing was launched
    int n_scales = 0, w[kMaxScales] = {}, h[kMaxScales] = {};
    uint32_t launches = 0;
};

// Everything that only depends on the source: its XYB pyramid, the rows pass of (a, a*a) (RECURSIVE blur) and the
// TMA descriptors over both.
struct SrcSet {
    float *d_pyr = nullptr, *d_hplanes = nullptr;
    // d_hplanes caches what only depends on the source, in the form the selected kernels want it: the ROWS pass of
    // (a, a*a) for the two-pass kernels, the fully blurred (mu1, sigma11) for the fused kernel
    bool rows_valid = false, musig_valid = false;
    cudaEvent_t pyr_ready = nullptr;       // source stream: the pyramid is complete
    cudaEvent_t ready = nullptr;           // source stream: ... and so is everything else enqueued by set_source
    cudaEvent_t last_use[2] = {nullptr, nullptr};   // per slot: its last submission reading this set has finished
    cudaEvent_t cache_done = nullptr;      // the launch that filled d_hplanes (whichever slot's stream it ran on)
    CUtensorMap in_src[kMaxScales], out_psrc[kMaxScales], in_musig[kMaxScales], cols_psrc[kMaxScales], cols_xa[kMaxScales];
    int maps_w = -1, maps_h = -1;
};

struct oavif_ssimu2_ctx {
    int device = 0;
    uint32_t max_w = 0, max_h = 0, max_batch = 0;
    cudaStream_t user_stream = nullptr;                    // set_stream: every submission's kernels go there instead
    CandSide *cs = nullptr;                                // candidate side of the submission being enqueued / retired last
    cudaStream_t copy_stream = nullptr;                    // host -> device uploads: run under the previous submission's kernels
    int blur_mode = OAVIF_SSIMU2_BLUR_RECURSIVE;
    int weight_layout = OAVIF_SSIMU2_WEIGHTS_SIX_SLOTS;
    int transfer = OAVIF_SSIMU2_TRANSFER_F64;
    int vertical_order = OAVIF_SSIMU2_VERTICAL_AS_HORIZONTAL;
    int tile_path = OAVIF_SSIMU2_TILES_TMA;
    int source_rows = OAVIF_SSIMU2_SOURCE_ROWS_WITH_FIRST_SCORE;
    unsigned cap_units = 0, wave_epoch = 0;               // fused kernel: capacities, launch counter (mailbox tags)
    long long cap_mailbox_words = 0;      // per candidate
    int *h_wave_err = nullptr, *dm_wave_err = nullptr;   // pinned, mapped
    unsigned long long *trace_buf = nullptr;             // oavif_ssimu2_debug_wave_trace

    // capacities (computed from max_w x max_h)
    long long cap_pyr_floats = 0, cap_in_bytes = 0, cap_ctas = 0, cap_hplane_floats = 0;

    // current image
    bool have_source = false;
    Geom g{};
    uint32_t last_n = 0;

    // source staging is double-buffered too: set_source of image i+1 uploads while image i is still being scored
    uint8_t *d_in_src[2] = {nullptr, nullptr};
    int src_buf = 0;
    cudaEvent_t src_up0 = nullptr, src_up1 = nullptr, src_consumed[2] = {nullptr, nullptr};
    Slot slot[kSlots];
    int head = 0, tail = 0, inflight = 0, last_slot = 0;   // ring of submissions; last_slot: what get_detail / get_timing report

    // The source side lives twice: set_source of image i+1 builds set (cur ^ 1) on the SOURCE stream while the
    // submissions of image i still read set cur on the compute stream.
    SrcSet src[2];
    int cur = 0;
    cudaStream_t src_stream = nullptr;
    cudaEvent_t ev_user = nullptr;         // orders the source stream behind a caller-owned compute stream
    float *d_lut = nullptr;
    const void **d_tbl = nullptr, **h_tbl = nullptr;   // [slot][3 * (max_batch + 1)]
    float *d_dbg = nullptr;
    long long dbg_floats = 0;
    uint8_t *d_conv = nullptr;     // oavif_ssimu2_yuv444_to_rgb8's output, grown on demand
    size_t conv_bytes = 0;
    // the pixels the last set_source_* call consumed, for oavif_ssimu2_source_samples: staged copy or caller's device pointer
    const void *src_pix = nullptr;
    size_t src_pix_stride = 0;
    int src_pix_channels = 0, src_pix_bits = 0;
    oavif_ssimu2_timing timing{};
    float taps[9] = {};
    IirCoef iir{};
    std::string err;
    // every large device buffer is followed by a 4 KB guard band filled with kGuardByte
    // (compute-sanitizer is not available on the target pool: oavif_ssimu2_debug_check_guards)
    std::vector<std::pair<unsigned char *, size_t>> guards;  // (start of guard, bytes)
};

namespace {

constexpr size_t kGuardBytes = 4096;
constexpr int kGuardByte = 0xA5;

template <class T>
cudaError_t alloc_guarded(oavif_ssimu2_ctx *ctx, T **out, size_t bytes)
{
    const size_t body = (bytes + 255) & ~(size_t)255;
    unsigned char *p = nullptr;
    cudaError_t e = cudaMalloc(&p, body + kGuardBytes);
    if (e != cudaSuccess) return e;
    e = cudaMemset(p + body, kGuardByte, kGuardBytes);
    ctx->guards.emplace_back(p + body, kGuardBytes);
    *out = reinterpret_cast<T *>(p);
    return e;
}

int fail(oavif_ssimu2_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_last_error = buf;
    if (ctx) ctx->err = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? OAVIF_SSIMU2_E_NOMEM : OAVIF_SSIMU2_E_CUDA, \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int rup(int a, int b) { return cdiv(a, b) * b; }

// Scale schedule of v2.1 §2: the size test looks at the previous scale's dimensions.
void make_geom(int w, int h, Geom *g)
{
    memset(g, 0, sizeof *g);
    const int tiles_x = cdiv(w, kPyrTile), tiles_y = cdiv(h, kPyrTile);
    int cw = w, ch = h;
    long long off = 0;
    for (int s = 0; s < kMaxScales; ++s) {
        if (cw < 8 || ch < 8) break;
        if (s) {
            cw = (cw + 1) / 2;
            ch = (ch + 1) / 2;
        }
        g->w[s] = cw;
        g->h[s] = ch;
        g->pitch[s] = rup(tiles_x * (kPyrTile >> s), 32);
        g->rows[s] = tiles_y * (kPyrTile >> s);
        g->plane[s] = (long long)g->pitch[s] * g->rows[s];
        g->off[s] = off;
        off += 3 * g->plane[s];
        g->n_scales = s + 1;
    }
    g->pyr_floats = off;
}

// Allocation never depends on the scale schedule: size for all six scales of max_w x max_h.
long long pyr_capacity(int w, int h)
{
    const int tiles_x = cdiv(w, kPyrTile), tiles_y = cdiv(h, kPyrTile);
    long long off = 0;
    for (int s = 0; s < kMaxScales; ++s)
        off += 3LL * rup(tiles_x * (kPyrTile >> s), 32) * (tiles_y * (kPyrTile >> s));
    return off;
}

struct BlurPlan {
    int first_cta[kMaxScales + 1];
    int per_channel[kMaxScales];
    int tiles_x[kMaxScales], tiles_y[kMaxScales];
    int total;
};

void plan_fir(const Geom &g, BlurPlan *p)
{
    memset(p, 0, sizeof *p);
    int acc = 0;
    for (int s = 0; s < g.n_scales; ++s) {
        p->tiles_x[s] = cdiv(g.w[s], kFirTW);
        p->tiles_y[s] = cdiv(g.h[s], kFirTH);
        p->per_channel[s] = p->tiles_x[s] * p->tiles_y[s];
        p->first_cta[s] = acc;
        acc += 3 * p->per_channel[s];
    }
    for (int s = g.n_scales; s <= kMaxScales; ++s) p->first_cta[s] = acc;
    p->total = acc;
}

void plan_iir_v(const Geom &g, BlurPlan *p)
{
    memset(p, 0, sizeof *p);
    int acc = 0;
    for (int s = 0; s < g.n_scales; ++s) {
        p->tiles_x[s] = cdiv(g.w[s], kIirVCols);
        p->tiles_y[s] = 1;
        p->per_channel[s] = p->tiles_x[s];
        p->first_cta[s] = acc;
        acc += 3 * p->per_channel[s];
    }
    for (int s = g.n_scales; s <= kMaxScales; ++s) p->first_cta[s] = acc;
    p->total = acc;
}

long long cta_capacity(int w, int h)
{
    Geom g;
    make_geom(w, h, &g);
    BlurPlan a, b;
    plan_fir(g, &a);
    plan_iir_v(g, &b);
    // smaller images inside the same bounding box never need more CTAs than the box itself,
    // but a transposed image can: size for the larger of both orientations
    Geom gt;
    make_geom(h, w, &gt);
    BlurPlan at, bt;
    plan_fir(gt, &at);
    plan_iir_v(gt, &bt);
    int m = a.total;
    if (b.total > m) m = b.total;
    if (at.total > m) m = at.total;
    if (bt.total > m) m = bt.total;
    return m + 64;
}

// Closed form of the 9 taps: h[m] = sum_k beta_k cos(omega_k m), |m| <= N-1 (Charalampidis 2016,
// truncated cosines k in {1,3,5}); also yields the recursion coefficients n2, d1.
void solve_gaussian(double sigma, float taps[9], IirCoef *iir)
{
    const double kPi = 3.141592653589793238;
    const double N = (double)std::lround(3.2795 * sigma + 0.2546);
    const double om[3] = {kPi / (2.0 * N), 3.0 * kPi / (2.0 * N), 5.0 * kPi / (2.0 * N)};
    const double p1 = 1.0 / std::tan(0.5 * om[0]), p3 = -1.0 / std::tan(0.5 * om[1]),
                 p5 = 1.0 / std::tan(0.5 * om[2]);
    const double r1 = p1 * p1 / std::sin(om[0]), r3 = -p3 * p3 / std::sin(om[1]),
                 r5 = p5 * p5 / std::sin(om[2]);
    double rho[3];
    for (int i = 0; i < 3; ++i) rho[i] = std::exp(-0.5 * sigma * sigma * om[i] * om[i]) / N;
    const double D13 = p1 * r3 - r1 * p3, D35 = p3 * r5 - r3 * p5, D51 = p5 * r1 - r5 * p1;
    const double z15 = D35 / D13, z35 = D51 / D13;
    // solve [p1 p3 p5; r1 r3 r5; z15 z35 1] beta = [1, N^2 - sigma^2, z15 rho1 + z35 rho3 + rho5]
    double A[3][4] = {{p1, p3, p5, 1.0},
                      {r1, r3, r5, N * N - sigma * sigma},
                      {z15, z35, 1.0, z15 * rho[0] + z35 * rho[1] + rho[2]}};
    for (int c = 0; c < 3; ++c) {  // Gauss-Jordan with partial pivoting
        int piv = c;
        for (int r = c + 1; r < 3; ++r)
            if (std::fabs(A[r][c]) > std::fabs(A[piv][c])) piv = r;
        for (int k = 0; k < 4; ++k) std::swap(A[c][k], A[piv][k]);
        for (int r = 0; r < 3; ++r) {
            if (r == c) continue;
            const double f = A[r][c] / A[c][c];
            for (int k = c; k < 4; ++k) A[r][k] -= f * A[c][k];
        }
    }
    double beta[3];
    for (int i = 0; i < 3; ++i) beta[i] = A[i][3] / A[i][i];
    const int R = (int)N - 1;  // 4
    for (int m = -R; m <= R; ++m) {
        double hm = 0.0;
        for (int i = 0; i < 3; ++i) hm += beta[i] * std::cos(om[i] * m);
        taps[m + R] = (float)hm;
    }
    for (int i = 0; i < 3; ++i) {
        iir->n2[i] = (float)(-beta[i] * std::cos(om[i] * (N + 1.0)));
        iir->d1[i] = (float)(-2.0 * std::cos(om[i]));
    }
}

int kind_of(int depth, int rgba_path)
{
    if (depth == 8) return IN_YUV8;
    if (depth == 10) return rgba_path ? IN_YUV10_RGBA : IN_YUV10_RGB;
    return -1;
}

bool yuv_consts(int matrix, YuvK *k)
{
    switch (matrix) {
    case 2: case 5: case 6: *k = YuvK{16320, 32, 113, 22, 46, 90}; return true;
    case 1: *k = YuvK{16320, 32, 119, 12, 30, 101}; return true;
    case 9: *k = YuvK{16320, 32, 120, 11, 37, 94}; return true;
    default: return false;
    }
}

template <int KIND>
void launch_pyr_kind(const PyrArgs &a, dim3 grid, cudaStream_t st)
{
    k_pyramid<KIND><<<grid, 256, 0, st>>>(a);
}

void launch_pyramid(int kind, const PyrArgs &a, int n, cudaStream_t st)
{
    const dim3 grid(cdiv(a.g.w[0], kPyrTile), cdiv(a.g.h[0], kPyrTile), n);
    switch (kind) {
    case IN_RGB8: launch_pyr_kind<IN_RGB8>(a, grid, st); break;
    case IN_YUV8: launch_pyr_kind<IN_YUV8>(a, grid, st); break;
    case IN_YUV10_RGB: launch_pyr_kind<IN_YUV10_RGB>(a, grid, st); break;
    case IN_YUV10_RGBA: launch_pyr_kind<IN_YUV10_RGBA>(a, grid, st); break;
    default: launch_pyr_kind<IN_PIXELS>(a, grid, st); break;
    }
}

struct InputDesc {
    int kind;            // InputKind
    size_t stride[3];    // caller strides, bytes
    int matrix;
    bool on_device;      // caller pointers are device pointers
    int channels = 3, hbd = 0;  // IN_PIXELS
};

// Bytes per row actually holding pixels, per plane.
size_t row_bytes(const InputDesc &d, int w)
{
    switch (d.kind) {
    case IN_RGB8: return 3u * (size_t)w;
    case IN_YUV8: return (size_t)w;
    case IN_PIXELS: return (size_t)w * d.channels * (d.hbd ? 2 : 1);
    default: return 2u * (size_t)w;
    }
}
int nplanes(int kind) { return (kind == IN_RGB8 || kind == IN_PIXELS) ? 1 : 3; }

// Stage (if host) `n` images and build their pyramids into `out`.  Host pixels go up on the COPY stream into
// `staging`; the compute stream waits for `up1` and runs the pyramid kernel.  `tbl0` selects the rows of the
// pointer table this call owns.
int build_pyramids(oavif_ssimu2_ctx *ctx, cudaStream_t st, const InputDesc &d, uint32_t n, const HostPlanes *imgs,
                   uint8_t *staging, int tbl0, cudaEvent_t up0, cudaEvent_t up1, cudaEvent_t k0, float *out,
                   long long out_stride)
{
    const Geom &g = ctx->g;
    const int w = g.w[0], h = g.h[0];
    PyrArgs a{};
    a.g = g;
    a.out = out;
    a.out_stride = out_stride;
    a.lut = ctx->d_lut;
    a.one = 1.0f;
    a.neg_one = -1.0f;
    if (d.kind != IN_RGB8 && d.kind != IN_PIXELS && !yuv_consts(d.matrix, &a.k))
        return fail(ctx, OAVIF_SSIMU2_E_UNSUPPORTED, "matrix_coefficients %d not on the scored path", d.matrix);
    const size_t rb = row_bytes(d, w);
    const int np = nplanes(d.kind);
    a.channels = d.channels;
    a.hbd = d.hbd;
    if (d.on_device) {
        for (uint32_t i = 0; i < n; ++i)
            for (int p = 0; p < 3; ++p) ctx->h_tbl[tbl0 + 3 * i + p] = imgs[i].p[p < np ? p : 0];
        for (int p = 0; p < 3; ++p) a.stride[p] = (long long)d.stride[p < np ? p : 0];
    } else {
        // tight rows on the device, each plane start 256-byte aligned
        const size_t plane_bytes = (rb * h + 255) & ~(size_t)255;
        CK(cudaEventRecord(up0, ctx->copy_stream));
        for (uint32_t i = 0; i < n; ++i)
            for (int p = 0; p < np; ++p) {
                uint8_t *dst = staging + ((size_t)i * np + p) * plane_bytes;
                CK(cudaMemcpy2DAsync(dst, rb, imgs[i].p[p], d.stride[p], rb, h, cudaMemcpyHostToDevice,
                                     ctx->copy_stream));
                ctx->h_tbl[tbl0 + 3 * i + p] = dst;
            }
        CK(cudaEventRecord(up1, ctx->copy_stream));
        CK(cudaStreamWaitEvent(st, up1, 0));
        if (np == 1)
            for (uint32_t i = 0; i < n; ++i)
                ctx->h_tbl[tbl0 + 3 * i + 1] = ctx->h_tbl[tbl0 + 3 * i + 2] = ctx->h_tbl[tbl0 + 3 * i];
        for (int p = 0; p < 3; ++p) a.stride[p] = (long long)rb;
    }
    if (n <= (uint32_t)kPyrInline) {  // pointers ride in the launch arguments: no table copy
        for (uint32_t i = 0; i < 3 * n; ++i) a.inl[i] = ctx->h_tbl[tbl0 + i];
        a.planes = nullptr;
    } else {
        CK(cudaMemcpyAsync(ctx->d_tbl + tbl0, ctx->h_tbl + tbl0, sizeof(void *) * 3 * n, cudaMemcpyHostToDevice, st));
        a.planes = ctx->d_tbl + tbl0;
    }
    if (k0) CK(cudaEventRecord(k0, st));
    launch_pyramid(d.kind, a, (int)n, st);
    CK(cudaGetLastError());
    return 0;
}

// sRGB -> linear table (v2.1 section 1), stored once per LANE (entry v, lane l at [32 v + l]): the pyramid kernel copies
// it to shared memory as is, where a warp's 32 gathers then hit 32 different banks.  TRANSFER_F64: binary64
// evaluation, one rounding to binary32; TRANSFER_F32: binary32 throughout (the oracle's ORACLE_VARIANT_F32_TRANSFER).
cudaError_t upload_srgb_table(oavif_ssimu2_ctx *ctx)
{
    std::vector<float> lut(256 * 32);
    for (int i = 0; i < 256; ++i) {
        float e;
        if (ctx->transfer == OAVIF_SSIMU2_TRANSFER_F32) {
            const float v = (float)i / 255.0f;
            e = (v <= 0.04045f) ? v / 12.92f : powf((v + 0.055f) / 1.055f, 2.4f);
        } else {
            const double v = (double)i / 255.0;
            e = (float)((v <= 0.04045) ? v / 12.92 : std::pow((v + 0.055) / 1.055, 2.4));
        }
        for (int l = 0; l < 32; ++l) lut[32 * i + l] = e;
    }
    return cudaMemcpy(ctx->d_lut, lut.data(), sizeof(float) * lut.size(), cudaMemcpyHostToDevice);
}

// Everything a w x h image needs against what the context allocated: pyramid floats, staged input bytes and
// — per-CTA partial sums are indexed by CTA — the CTA plans of both blur modes for THIS image (an image inside
// the pyramid/input capacity can still need more CTAs than the max_w x max_h box or its transpose).
long long ctas_needed(int w, int h)
{
    Geom g;
    make_geom(w, h, &g);
    BlurPlan a, b;
    plan_fir(g, &a);
    plan_iir_v(g, &b);
    return a.total > b.total ? a.total : b.total;
}

bool fits_capacity(const oavif_ssimu2_ctx *ctx, uint32_t w, uint32_t h)
{
    if (!(pyr_capacity((int)w, (int)h) <= ctx->cap_pyr_floats && (long long)w * h * 8 + 3 * 256 <= ctx->cap_in_bytes &&
          ctas_needed((int)w, (int)h) <= ctx->cap_ctas))
        return false;
    Geom g;   // the fused kernel's unit list and mailbox
    long long so[kMaxScales];
    int rows[kMaxScales];
    make_geom((int)w, (int)h, &g);
    return wave_units(g).size() <= ctx->cap_units && wave_mailbox_words(g, 5, so, rows) <= ctx->cap_mailbox_words;
}

int check_size(oavif_ssimu2_ctx *ctx, uint32_t w, uint32_t h)
{
    if (w == 0 || h == 0) return fail(ctx, OAVIF_SSIMU2_E_ARG, "zero image dimension");
    if (w > (1u << 16) || h > (1u << 16)) return fail(ctx, OAVIF_SSIMU2_E_ARG, "dimension above 65536");
    // kernels address a plane with 32-bit element offsets, and an interleaved pair plane is two planes wide
    if ((long long)rup(cdiv((int)w, kPyrTile) * kPyrTile, 32) * (cdiv((int)h, kPyrTile) * kPyrTile) >= (1LL << 30))
        return fail(ctx, OAVIF_SSIMU2_E_UNSUPPORTED, "image %ux%u: a plane exceeds 2^30 samples", w, h);
    if (!fits_capacity(ctx, w, h))
        return fail(ctx, OAVIF_SSIMU2_E_STATE, "image %ux%u exceeds context capacity %ux%u", w, h, ctx->max_w,
                    ctx->max_h);
    return 0;
}

// The candidate side of a slot: buffers sized like the context, a compute stream of its own.  Slot 0 at ctx_create;
// slot 1 the first time a submission arrives while another is in flight.
cudaError_t alloc_cand_side(oavif_ssimu2_ctx *ctx, CandSide &c)
{
    if (c.allocated) return cudaSuccess;
    cudaError_t e;
#define TRY(x) if ((e = (x)) != cudaSuccess) return e
    TRY(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
    c.stream = ctx->user_stream ? ctx->user_stream : c.own_stream;
    TRY(alloc_guarded(ctx, &c.d_dist_pyr, sizeof(float) * ctx->cap_pyr_floats * ctx->max_batch));
    TRY(alloc_guarded(ctx, &c.d_hplanes, sizeof(float) * ctx->cap_hplane_floats * ctx->max_batch));
    TRY(alloc_guarded(ctx, &c.d_partials, sizeof(double) * 6 * ctx->cap_ctas * ctx->max_batch));
    TRY(cudaMalloc(&c.d_units, sizeof(unsigned) * ctx->cap_units));
    TRY(cudaMalloc(&c.d_ticket, sizeof(unsigned)));
    TRY(cudaMemset(c.d_ticket, 0, sizeof(unsigned)));
    TRY(alloc_guarded(ctx, &c.d_mailbox, sizeof(unsigned long long) * ctx->cap_mailbox_words * ctx->max_batch));
    TRY(cudaMemset(c.d_mailbox, 0, sizeof(unsigned long long) * ctx->cap_mailbox_words * ctx->max_batch));
#undef TRY
    c.allocated = true;
    return cudaSuccess;
}

// TMA descriptors follow the geometry.  Fills `out` with the candidates' descriptors and those of source set `S`;
// false selects the cp.async kernels (tile path switched off, or no descriptor encoder in this driver).
bool rows_maps_for(oavif_ssimu2_ctx *ctx, SrcSet &S, IirRowsTmaMaps *out)
{
    if (ctx->tile_path == OAVIF_SSIMU2_TILES_CP_ASYNC) return false;
    const Geom &g = ctx->g;
    const long long P = ctx->cap_pyr_floats;
    bool ok = true;
    if (ctx->cs->maps_w != g.w[0] || ctx->cs->maps_h != g.h[0]) {
        ok = iir_rows_tma_maps_cand(&ctx->cs->cand_maps, g, ctx->cs->d_dist_pyr, P, ctx->cs->d_hplanes, ctx->cs->d_hplanes + 2 * P, 3 * P,
                                    (int)ctx->max_batch) &&
             iir_cols_tma_maps_cand(&ctx->cs->cols_maps, g, ctx->cs->d_hplanes, ctx->cs->d_hplanes + 2 * P, 3 * P, ctx->cs->d_dist_pyr, P,
                                    (int)ctx->max_batch);
        if (ok) {
            ctx->cs->maps_w = g.w[0];
            ctx->cs->maps_h = g.h[0];
        }
    }
    if (ok && (S.maps_w != g.w[0] || S.maps_h != g.h[0])) {
        ok = iir_rows_tma_maps_src(S.in_src, S.out_psrc, g, S.d_pyr, S.d_hplanes) && iir_cols_tma_maps_src(S.cols_psrc, S.cols_xa, g, S.d_hplanes, S.d_pyr);
        for (int s = 0; s < g.n_scales && ok; ++s)   // the fused kernel's view of the same buffer: (mu1, sigma11) pairs
            ok = tma_make_4d(&S.in_musig[s], S.d_hplanes + 2 * g.off[s], 2ull * g.w[s], (uint64_t)g.h[s], 3, 1,
                             (uint64_t)g.pitch[s] * 8, (uint64_t)g.plane[s] * 8, 0, 2 * kIirVCols, kWvB, false);
        if (ok) {
            S.maps_w = g.w[0];
            S.maps_h = g.h[0];
        }
    }
    if (!ok) {
        ctx->tile_path = OAVIF_SSIMU2_TILES_CP_ASYNC;
        return false;
    }
    *out = ctx->cs->cand_maps;
    memcpy(out->in_src, S.in_src, sizeof S.in_src);
    memcpy(out->out_psrc, S.out_psrc, sizeof S.out_psrc);
    return true;
}

// the columns pass's descriptors (nullptr: the cp.async loader)
const IirColsTmaMaps *cols_maps_for(oavif_ssimu2_ctx *ctx, SrcSet &S, IirColsTmaMaps *tmp)
{
    IirRowsTmaMaps rows;
    if (!rows_maps_for(ctx, S, &rows)) return nullptr;   // also brings both sets of descriptors up to date
    *tmp = ctx->cs->cols_maps;
    memcpy(tmp->psrc, S.cols_psrc, sizeof tmp->psrc);
    memcpy(tmp->xa, S.cols_xa, sizeof tmp->xa);
    return tmp;
}

IirArgs iir_args_for(oavif_ssimu2_ctx *ctx, const SrcSet &S)
{
    IirArgs a{};
    const IirBuffers B{S.d_hplanes, ctx->cs->d_hplanes, ctx->cap_pyr_floats};
    iir_fill_common(a, ctx->g, ctx->iir, S.d_pyr, ctx->cs->d_dist_pyr, ctx->cap_pyr_floats, B);
    a.partials = ctx->cs->d_partials;
    a.partials_stride = ctx->cap_ctas * 6;
    return a;
}

// The fused kernel over candidates [cand0, cand0 + n): MODE 2 computes the source's quantities too and leaves
// (mu1, sigma11) in the set's cache, MODE 1 reads them back.
int enqueue_wave(oavif_ssimu2_ctx *ctx, SrcSet &Src, const BlurPlan &plan, int mode, uint32_t cand0, uint32_t n,
                 const IirDebugTap *tap = nullptr)
{
    const Geom &g = ctx->g;
    IirRowsTmaMaps rm;
    if (!rows_maps_for(ctx, Src, &rm)) return fail(ctx, OAVIF_SSIMU2_E_CUDA, "no tensor-map encoder: the fused kernel needs TMA");
    if (ctx->cs->units_w != g.w[0] || ctx->cs->units_h != g.h[0]) {
        const std::vector<unsigned> u = wave_units(g);
        if (u.size() > ctx->cap_units) return fail(ctx, OAVIF_SSIMU2_E_STATE, "unit list exceeds capacity");
        // pageable source: the copy has been staged when the call returns, and the stream orders it before the launch
        CK(cudaMemcpyAsync(ctx->cs->d_units, u.data(), u.size() * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->cs->stream));
        ctx->cs->n_units = (unsigned)u.size();
        ctx->cs->units_w = g.w[0];
        ctx->cs->units_h = g.h[0];
    }
    WaveMaps wm;
    memcpy(wm.in_src, rm.in_src, sizeof wm.in_src);
    memcpy(wm.in_dist, rm.in_dist, sizeof wm.in_dist);
    memcpy(wm.in_musig, Src.in_musig, sizeof wm.in_musig);
    WaveArgs a{};
    a.g = g;
    a.k = ctx->iir;
    a.one = 1.0f;
    a.neg_one = -1.0f;
    a.units = ctx->cs->d_units;
    a.n_units = ctx->cs->n_units;
    a.n_cand = n;
    a.cand0 = cand0;
    a.ticket = ctx->cs->d_ticket;
    a.epoch = ++ctx->wave_epoch & 0xfffffu;
    if (a.epoch == 0) a.epoch = ctx->wave_epoch = 1;
    a.mailbox = ctx->cs->d_mailbox;
    const long long words = wave_mailbox_words(g, mode == 2 ? 5 : 3, a.mb_scale_off, a.mb_rows);
    if (words > ctx->cap_mailbox_words) return fail(ctx, OAVIF_SSIMU2_E_STATE, "mailbox exceeds capacity");
    a.mb_cand_stride = ctx->cap_mailbox_words;
    a.musig = Src.d_hplanes;
    a.partials = ctx->cs->d_partials;
    a.partials_stride = ctx->cap_ctas * 6;
    for (int s = 0; s <= kMaxScales; ++s) a.first_cta[s] = plan.first_cta[s];
    for (int s = 0; s < kMaxScales; ++s) a.blocks[s] = plan.tiles_x[s];
    a.error_flag = ctx->dm_wave_err;
    cudaError_t e;
    if (tap) {
        a.dbg_cols = tap->out;
        a.dbg_scale = tap->scale;
        a.dbg_channel = tap->channel;
        a.dbg_cand = tap->cand;
        a.dbg_trace = ctx->trace_buf;
        e = mode == 2 ? wave_launch<2, true>(a, wm, ctx->cs->stream) : wave_launch<1, true>(a, wm, ctx->cs->stream);
    } else {
        e = mode == 2 ? wave_launch<2, false>(a, wm, ctx->cs->stream) : wave_launch<1, false>(a, wm, ctx->cs->stream);
    }
    if (e != cudaSuccess) return fail(ctx, OAVIF_SSIMU2_E_CUDA, "fused blur launch: %s", cudaGetErrorString(e));
    return 0;
}

// blur + maps + pooling + final score of the candidates whose pyramids were just enqueued, into slot S.
// The compute stream has waited for the source's pyramid; `Src.ready` covers the rest of the source side.
int enqueue_blur_and_finalize(oavif_ssimu2_ctx *ctx, Slot &S, SrcSet &Src, uint32_t n)
{
    const Geom &g = ctx->g;
    BlurPlan plan;
    if (ctx->blur_mode == OAVIF_SSIMU2_BLUR_FIR) {
        plan_fir(g, &plan);
        BlurArgs b{};
        b.g = g;
        b.src = Src.d_pyr;
        b.dist = ctx->cs->d_dist_pyr;
        b.dist_stride = ctx->cap_pyr_floats;
        b.partials = ctx->cs->d_partials;
        b.partials_stride = ctx->cap_ctas * 6;
        for (int s = 0; s <= kMaxScales; ++s) b.first_cta[s] = plan.first_cta[s];
        for (int s = 0; s < kMaxScales; ++s) {
            b.tiles_x[s] = plan.tiles_x[s];
            b.tiles_y[s] = plan.tiles_y[s];
        }
        memcpy(b.taps, ctx->taps, sizeof b.taps);
        b.one = 1.0f;
        b.neg_one = -1.0f;
        k_fir_fused<<<dim3(plan.total, n), kFirThreads, kFirSmemBytes, ctx->cs->stream>>>(b);
        CK(cudaGetLastError());
        CK(cudaEventRecord(S.k[2], ctx->cs->stream));
        S.launches += 1;
    } else if (ctx->tile_path == OAVIF_SSIMU2_TILES_FUSED) {
        plan_iir_v(g, &plan);
        CK(cudaStreamWaitEvent(ctx->cs->stream, Src.ready, 0));
        uint32_t first = 0;
        if (!Src.musig_valid) {   // first call after set_source: candidate 0's launch carries the source's quantities
            const int rc = enqueue_wave(ctx, Src, plan, 2, 0, 1);
            if (rc) return rc;
            CK(cudaEventRecord(Src.cache_done, ctx->cs->stream));
            Src.musig_valid = true;
            Src.rows_valid = false;
            S.launches += 1;
            first = 1;
        }
        if (n > first) {
            CK(cudaStreamWaitEvent(ctx->cs->stream, Src.cache_done, 0));   // filled by another slot's stream, perhaps
            const int rc = enqueue_wave(ctx, Src, plan, 1, first, n - first);
            if (rc) return rc;
            S.launches += 1;
        }
        CK(cudaEventRecord(S.k[2], ctx->cs->stream));
    } else {
        plan_iir_v(g, &plan);
        const IirArgs a = iir_args_for(ctx, Src);
        IirRowsTmaMaps maps;
        const bool tma = rows_maps_for(ctx, Src, &maps);
        cudaError_t e;
        bool filled_cache = false;
        if (Src.rows_valid) {
            // the source's half is cached (or was enqueued by set_source on the source stream and may still be
            // running next to this launch): only the columns pass has to wait for it
            e = launch_iir_rows(a, g, 1, (int)n, ctx->cs->stream, tma ? &maps : nullptr);
            S.launches += 1;
        } else {            // the source's half rides in this launch (candidate 0's CTAs carry it along)
            e = launch_iir_rows(a, g, 2, (int)n, ctx->cs->stream, tma ? &maps : nullptr);
            S.launches += 1;
            filled_cache = true;
        }
        if (e != cudaSuccess) return fail(ctx, OAVIF_SSIMU2_E_CUDA, "rows launch: %s", cudaGetErrorString(e));
        if (filled_cache) CK(cudaEventRecord(Src.cache_done, ctx->cs->stream));
        Src.rows_valid = true;
        Src.musig_valid = false;
        CK(cudaEventRecord(S.k[2], ctx->cs->stream));
        CK(cudaStreamWaitEvent(ctx->cs->stream, Src.ready, 0));
        CK(cudaStreamWaitEvent(ctx->cs->stream, Src.cache_done, 0));
        IirColsTmaMaps cmaps;
        e = launch_iir_cols(a, plan.first_cta, plan.tiles_x, (int)n, ctx->cs->stream, nullptr, cols_maps_for(ctx, Src, &cmaps),
                            ctx->tile_path == OAVIF_SSIMU2_TILES_TMA_DECOUPLED, 0,
                            ctx->vertical_order == OAVIF_SSIMU2_VERTICAL_FUSED_OUTER);
        if (e != cudaSuccess) return fail(ctx, OAVIF_SSIMU2_E_CUDA, "columns launch: %s", cudaGetErrorString(e));
        S.launches += 1;
    }
    CK(cudaEventRecord(S.k[3], ctx->cs->stream));

    FinalArgs f{};
    f.n_scales = g.n_scales;
    for (int s = 0; s < kMaxScales; ++s) {
        f.w[s] = g.w[s];
        f.h[s] = g.h[s];
        f.ctas_per_channel[s] = plan.per_channel[s];
    }
    for (int s = 0; s <= kMaxScales; ++s) f.first_cta[s] = plan.first_cta[s];
    f.partials = ctx->cs->d_partials;
    f.partials_stride = ctx->cap_ctas * 6;
    f.sums = S.dm_sums;
    f.scores = S.dm_scores;
    f.contiguous_weights = ctx->weight_layout == OAVIF_SSIMU2_WEIGHTS_CONTIGUOUS;
    k_finalize<<<n, 1024, 0, ctx->cs->stream>>>(f);
    CK(cudaGetLastError());
    S.launches += 1;
    CK(cudaEventRecord(S.k[4], ctx->cs->stream));
    CK(cudaEventRecord(Src.last_use[&S - ctx->slot], ctx->cs->stream));
    return 0;
}

// Enqueue one submission (upload on the copy stream, kernels on the compute stream) and return at once.
int submit_common(oavif_ssimu2_ctx *ctx, const InputDesc &d, uint32_t n, const HostPlanes *imgs)
{
    if (!ctx) return fail(nullptr, OAVIF_SSIMU2_E_ARG, "null context");
    if (!imgs || n == 0) return fail(ctx, OAVIF_SSIMU2_E_ARG, "null argument or empty batch");
    if (!ctx->have_source) return fail(ctx, OAVIF_SSIMU2_E_STATE, "score called before set_source");
    if (n > ctx->max_batch) return fail(ctx, OAVIF_SSIMU2_E_STATE, "batch %u exceeds max_batch %u", n, ctx->max_batch);
    if (ctx->inflight == kSlots)
        return fail(ctx, OAVIF_SSIMU2_E_STATE, "%d submissions already in flight: call oavif_ssimu2_wait first", kSlots);
    if (ctx->vertical_order != OAVIF_SSIMU2_VERTICAL_AS_HORIZONTAL && ctx->blur_mode == OAVIF_SSIMU2_BLUR_RECURSIVE &&
        (ctx->tile_path != OAVIF_SSIMU2_TILES_TMA || !tma_encoder()))
        return fail(ctx, OAVIF_SSIMU2_E_UNSUPPORTED, "the FUSED_OUTER vertical order exists for the default (TMA) tile path only");
    const int w = ctx->g.w[0];
    const int np = nplanes(d.kind);
    for (uint32_t i = 0; i < n; ++i)
        for (int p = 0; p < np; ++p)
            if (!imgs[i].p[p]) return fail(ctx, OAVIF_SSIMU2_E_ARG, "null plane pointer (candidate %u)", i);
    for (int p = 0; p < np; ++p)
        if (d.stride[p] < row_bytes(d, w)) return fail(ctx, OAVIF_SSIMU2_E_ARG, "stride smaller than a row");
    CK(cudaSetDevice(ctx->device));
    if (ctx->inflight == 0) ctx->head = ctx->tail = 0;   // nothing to overlap with: a caller that never overlaps never pays for slot 1
    Slot &S = ctx->slot[ctx->head];
    if (!S.d_in_dist && !d.on_device && ctx->g.n_scales)   // the second slot's buffers exist only once submissions overlap
        CK(alloc_guarded(ctx, &S.d_in_dist, (size_t)ctx->cap_in_bytes * ctx->max_batch));
    CK(alloc_cand_side(ctx, S.cs));
    S.cs.stream = ctx->user_stream ? ctx->user_stream : S.cs.own_stream;
    ctx->cs = &S.cs;
    S.n = n;
    S.launches = 0;
    S.host_input = !d.on_device;
    S.n_scales = ctx->g.n_scales;
    for (int s = 0; s < kMaxScales; ++s) {
        S.w[s] = ctx->g.w[s];
        S.h[s] = ctx->g.h[s];
    }
    S.degenerate = ctx->g.n_scales == 0;   // below 8x8 nothing is evaluated: the published result is 100
    ctx->last_n = n;
    if (!S.degenerate) {
        const int tbl0 = ctx->head * 3 * (int)(ctx->max_batch + 1) + 3;   // row 0 of each table is the source's
        SrcSet &Src = ctx->src[ctx->cur];
        int rc = build_pyramids(ctx, ctx->cs->stream, d, n, imgs, S.d_in_dist, tbl0, S.up0, S.up1, S.k[0], ctx->cs->d_dist_pyr,
                                ctx->cap_pyr_floats);
        if (rc) return rc;
        S.launches += 1;
        CK(cudaEventRecord(S.k[1], ctx->cs->stream));
        CK(cudaStreamWaitEvent(ctx->cs->stream, Src.pyr_ready, 0));   // the candidate's pyramid did not need the source's
        rc = enqueue_blur_and_finalize(ctx, S, Src, n);
        if (rc) return rc;
    }
    ctx->head = (ctx->head + 1) % kSlots;
    ctx->inflight += 1;
    return 0;
}

// Retire the oldest submission: block until its score is there.
int wait_common(oavif_ssimu2_ctx *ctx, double *scores)
{
    if (!ctx) return fail(nullptr, OAVIF_SSIMU2_E_ARG, "null context");
    if (!scores) return fail(ctx, OAVIF_SSIMU2_E_ARG, "null argument");
    if (ctx->inflight == 0) return fail(ctx, OAVIF_SSIMU2_E_STATE, "wait without a submission in flight");
    Slot &S = ctx->slot[ctx->tail];
    ctx->last_slot = ctx->tail;
    ctx->tail = (ctx->tail + 1) % kSlots;
    ctx->inflight -= 1;
    ctx->timing = oavif_ssimu2_timing{};
    if (S.degenerate) {
        memset(S.h_sums, 0, sizeof(double) * S.n * kMaxScales * 18);
        for (uint32_t i = 0; i < S.n; ++i) scores[i] = S.h_scores[i] = 100.0;
        return 0;
    }
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(S.k[4]));
    if (*ctx->h_wave_err) {
        *ctx->h_wave_err = 0;
        return fail(ctx, OAVIF_SSIMU2_E_CUDA, "fused blur: a strip gave up waiting for its left neighbour's row state");
    }
    for (uint32_t i = 0; i < S.n; ++i) scores[i] = S.h_scores[i];
    float ms = 0.f;
    oavif_ssimu2_timing &t = ctx->timing;
    if (S.host_input) { cudaEventElapsedTime(&ms, S.up0, S.up1); t.h2d_ms = ms; }
    cudaEventElapsedTime(&ms, S.k[0], S.k[1]); t.pyramid_ms = ms;
    cudaEventElapsedTime(&ms, S.k[1], S.k[3]); t.blur_ms = ms;
    cudaEventElapsedTime(&ms, S.k[1], S.k[2]); t.blur_a_ms = ms;
    cudaEventElapsedTime(&ms, S.k[2], S.k[3]); t.blur_b_ms = ms;
    cudaEventElapsedTime(&ms, S.k[3], S.k[4]); t.finalize_ms = ms;
    cudaEventElapsedTime(&ms, S.host_input ? S.up0 : S.k[0], S.k[4]); t.total_ms = ms;
    t.launches = S.launches;
    return 0;
}

// the synchronous form: one submission, retired at once (nothing else may be in flight)
int score_common(oavif_ssimu2_ctx *ctx, const InputDesc &d, uint32_t n, const HostPlanes *imgs, double *scores)
{
    if (ctx && !scores) return fail(ctx, OAVIF_SSIMU2_E_ARG, "null argument or empty batch");
    if (ctx && ctx->inflight)
        return fail(ctx, OAVIF_SSIMU2_E_STATE, "a submission is in flight: retire it with oavif_ssimu2_wait first");
    const int rc = submit_common(ctx, d, n, imgs);
    if (rc) return rc;
    return wait_common(ctx, scores);
}

int set_source_common(oavif_ssimu2_ctx *ctx, const void *rgb, uint32_t w, uint32_t h, size_t stride,
                      bool on_device, int channels = 3, int bits = 8)
{
    if (!ctx) return fail(nullptr, OAVIF_SSIMU2_E_ARG, "null context");
    if (!rgb) return fail(ctx, OAVIF_SSIMU2_E_ARG, "null source pointer");
    if (channels < 1 || channels > 4 || (bits != 8 && bits != 16))
        return fail(ctx, OAVIF_SSIMU2_E_UNSUPPORTED, "UnsupportedChannelCount: %d channels, %d bits", channels, bits);
    int rc = check_size(ctx, w, h);
    if (rc) return rc;
    if (stride < (size_t)w * channels * (bits / 8)) return fail(ctx, OAVIF_SSIMU2_E_ARG, "source stride smaller than a row");
    CK(cudaSetDevice(ctx->device));
    ctx->have_source = false;
    ctx->src_pix = nullptr;
    make_geom((int)w, (int)h, &ctx->g);
    if (ctx->g.n_scales == 0) {  // geometry still needed for argument checks in score_*
        ctx->g.w[0] = (int)w;
        ctx->g.h[0] = (int)h;
        ctx->have_source = true;
        return 0;
    }
    InputDesc d{};
    d.kind = (channels == 3 && bits == 8) ? IN_RGB8 : IN_PIXELS;
    d.channels = channels;
    d.hbd = bits == 16;
    d.stride[0] = stride;
    d.on_device = on_device;
    HostPlanes hp{{rgb, nullptr, nullptr}};
    ctx->timing = oavif_ssimu2_timing{};
    // The source side is built on the SOURCE stream into the set no submission reads: submissions of the previous
    // image may still be in flight on the compute stream (pipelined callers), and within one evaluation the
    // source's pyramid and rows pass run next to the candidate's.
    ctx->cur ^= 1;
    SrcSet &Src = ctx->src[ctx->cur];
    Src.rows_valid = false;
    Src.musig_valid = false;
    cudaStream_t ss = ctx->src_stream;
    for (cudaEvent_t e : Src.last_use) CK(cudaStreamWaitEvent(ss, e, 0));   // submissions that read this set two images ago
    const int sb = ctx->src_buf ^= 1;
    if (on_device) {
        if (ctx->user_stream != nullptr) {         // caller-owned compute stream: its earlier work produced the pixels
            CK(cudaEventRecord(ctx->ev_user, ctx->cs->stream));
            CK(cudaStreamWaitEvent(ss, ctx->ev_user, 0));
        }
    } else {
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->src_consumed[sb], 0));   // the pyramid kernel that last read this staging buffer
    }
    rc = build_pyramids(ctx, ss, d, 1, &hp, ctx->d_in_src[sb], ctx->head * 3 * (int)(ctx->max_batch + 1), ctx->src_up0,
                        ctx->src_up1, nullptr, Src.d_pyr, 0);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->src_consumed[sb], ss));
    CK(cudaEventRecord(Src.pyr_ready, ss));
    ctx->src_pix = on_device ? rgb : ctx->d_in_src[sb];
    ctx->src_pix_stride = on_device ? stride : row_bytes(d, (int)w);
    ctx->src_pix_channels = channels;
    ctx->src_pix_bits = bits;
    ctx->timing.launches = 1;
    if (ctx->blur_mode == OAVIF_SSIMU2_BLUR_RECURSIVE &&
        (ctx->tile_path == OAVIF_SSIMU2_TILES_TMA || ctx->tile_path == OAVIF_SSIMU2_TILES_TMA_DECOUPLED) &&
        ctx->source_rows == OAVIF_SSIMU2_SOURCE_ROWS_AT_SET_SOURCE) {
        IirRowsTmaMaps maps;
        if (rows_maps_for(ctx, Src, &maps)) {         // rows pass of (a, a*a), once per source
            const cudaError_t e = launch_iir_rows(iir_args_for(ctx, Src), ctx->g, 3, 1, ss, &maps);
            if (e != cudaSuccess) return fail(ctx, OAVIF_SSIMU2_E_CUDA, "source rows launch: %s", cudaGetErrorString(e));
            Src.rows_valid = true;
            ctx->timing.launches = 2;
        }
    }
    CK(cudaEventRecord(Src.cache_done, ss));
    CK(cudaEventRecord(Src.ready, ss));
    // Return as soon as the caller's pixels have been consumed (host input: after the upload; device
    // input: at once): the kernels keep running behind the calls that follow.
    if (!on_device) {
        CK(cudaEventSynchronize(ctx->src_up1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->src_up0, ctx->src_up1); ctx->timing.h2d_ms = ms;
        ctx->timing.total_ms = ms;
    }
    ctx->have_source = true;
    return 0;
}

}  // namespace

// ================================== C ABI ======================================================

extern "C" {

int oavif_ssimu2_abi_version(void) { return OAVIF_SSIMU2_ABI_VERSION; }

const char *oavif_ssimu2_last_error(const oavif_ssimu2_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : t_last_error.c_str();
}

int oavif_ssimu2_device_pci_bus_id(int device, char *out, size_t cap)
{
    if (!out || cap < 16) return fail(nullptr, OAVIF_SSIMU2_E_ARG, "buffer too small");
    if (cudaDeviceGetPCIBusId(out, (int)cap, device) != cudaSuccess) return fail(nullptr, OAVIF_SSIMU2_E_CUDA, "no such device %d", device);
    for (char *p = out; *p; ++p)
        if (*p >= 'A' && *p <= 'Z') *p = (char)(*p - 'A' + 'a');   // sysfs spells the address in lower case
    return 0;
}

void *oavif_ssimu2_pinned_alloc(size_t bytes)
{
    void *p = nullptr;
    if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        t_last_error = "cudaHostAlloc failed";
        return nullptr;
    }
    return p;
}

void oavif_ssimu2_pinned_free(void *p)
{
    if (p) cudaFreeHost(p);
}

void oavif_ssimu2_ctx_destroy(oavif_ssimu2_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->src_stream) cudaStreamSynchronize(ctx->src_stream);
    for (auto &S : ctx->slot)
        if (S.cs.own_stream) cudaStreamSynchronize(S.cs.own_stream);
    for (auto &p : ctx->d_in_src) cudaFree(p);
    for (auto &S : ctx->slot) {
        cudaFree(S.d_in_dist);
        cudaFreeHost(S.h_sums);
        cudaFreeHost(S.h_scores);
        for (cudaEvent_t e : {S.up0, S.up1, S.k[0], S.k[1], S.k[2], S.k[3], S.k[4]})
            if (e) cudaEventDestroy(e);
    }
    for (cudaEvent_t e : {ctx->src_up0, ctx->src_up1, ctx->src_consumed[0], ctx->src_consumed[1]})
        if (e) cudaEventDestroy(e);
    for (auto &S : ctx->src) {
        cudaFree(S.d_pyr);
        cudaFree(S.d_hplanes);
        for (cudaEvent_t e : {S.pyr_ready, S.ready, S.last_use[0], S.last_use[1], S.cache_done})
            if (e) cudaEventDestroy(e);
    }
    if (ctx->ev_user) cudaEventDestroy(ctx->ev_user);
    if (ctx->src_stream) cudaStreamDestroy(ctx->src_stream);
    cudaFree(ctx->d_lut);
    cudaFreeHost(ctx->h_wave_err);
    cudaFree((void *)ctx->d_tbl);
    cudaFree(ctx->d_dbg);
    cudaFree(ctx->d_conv);
    cudaFreeHost((void *)ctx->h_tbl);
    for (auto &S : ctx->slot) {
        CandSide &c = S.cs;
        cudaFree(c.d_dist_pyr);
        cudaFree(c.d_hplanes);
        cudaFree(c.d_partials);
        cudaFree(c.d_units);
        cudaFree(c.d_ticket);
        cudaFree(c.d_mailbox);
        if (c.own_stream) cudaStreamDestroy(c.own_stream);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

int oavif_ssimu2_ctx_create(int device, uint32_t max_w, uint32_t max_h, uint32_t max_batch,
                            oavif_ssimu2_ctx **out)
{
    if (!out) return fail(nullptr, OAVIF_SSIMU2_E_ARG, "null out pointer");
    *out = nullptr;
    if (max_w == 0 || max_h == 0 || max_batch == 0 || max_w > (1u << 16) || max_h > (1u << 16))
        return fail(nullptr, OAVIF_SSIMU2_E_ARG, "bad context capacity %ux%ux%u", max_w, max_h, max_batch);
    oavif_ssimu2_ctx *ctx = new (std::nothrow) oavif_ssimu2_ctx();
    if (!ctx) return fail(nullptr, OAVIF_SSIMU2_E_NOMEM, "out of host memory");
    ctx->cs = &ctx->slot[0].cs;
    ctx->device = device;
    ctx->max_w = max_w;
    ctx->max_h = max_h;
    ctx->max_batch = max_batch;
#define CKC(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            const int rc_ = fail(nullptr, e_ == cudaErrorMemoryAllocation ? OAVIF_SSIMU2_E_NOMEM : OAVIF_SSIMU2_E_CUDA, \
                                 "%s failed: %s", #call, cudaGetErrorString(e_));                  \
            oavif_ssimu2_ctx_destroy(ctx);                                                         \
            return rc_;                                                                            \
        }                                                                                          \
    } while (0)
    CKC(cudaSetDevice(device));
    CKC(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&ctx->src_stream, cudaStreamNonBlocking));
    CKC(cudaEventCreateWithFlags(&ctx->ev_user, cudaEventDisableTiming));
    for (auto &S : ctx->src)
        for (cudaEvent_t *e : {&S.pyr_ready, &S.ready, &S.last_use[0], &S.last_use[1], &S.cache_done})
            CKC(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    for (auto &S : ctx->slot) {
        CKC(cudaEventCreate(&S.up0));
        CKC(cudaEventCreate(&S.up1));
        for (auto &e : S.k) CKC(cudaEventCreate(&e));
    }
    CKC(cudaEventCreate(&ctx->src_up0));
    CKC(cudaEventCreate(&ctx->src_up1));
    for (auto &e : ctx->src_consumed) CKC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));

    const int mw = (int)max_w, mh = (int)max_h;
    ctx->cap_pyr_floats = std::max(pyr_capacity(mw, mh), pyr_capacity(mh, mw));
    ctx->cap_in_bytes = (long long)mw * mh * 8 + 3 * 256;   // up to RGBA16 (8 B/px)
    ctx->cap_ctas = cta_capacity(mw, mh);
    ctx->cap_hplane_floats = iir_hplane_floats(ctx->cap_pyr_floats);

    for (auto &p : ctx->d_in_src) CKC(alloc_guarded(ctx, &p, (size_t)ctx->cap_in_bytes));
    CKC(alloc_guarded(ctx, &ctx->slot[0].d_in_dist, (size_t)ctx->cap_in_bytes * max_batch));
    for (auto &S : ctx->src) {
        CKC(alloc_guarded(ctx, &S.d_pyr, sizeof(float) * ctx->cap_pyr_floats));
        CKC(alloc_guarded(ctx, &S.d_hplanes, sizeof(float) * 2 * ctx->cap_pyr_floats));
    }
    CKC(cudaMalloc(&ctx->d_lut, sizeof(float) * 256 * 32));
    {   // fused kernel: unit list, ticket, mailbox (sized for the taller orientation of the capacity box), error flag
        Geom gm;
        long long so[kMaxScales];
        int rows[kMaxScales];
        make_geom(mw, mh, &gm);
        ctx->cap_units = (unsigned)wave_units(gm).size();
        long long words = wave_mailbox_words(gm, 5, so, rows);
        make_geom(mh, mw, &gm);
        ctx->cap_units = std::max<unsigned>(ctx->cap_units, (unsigned)wave_units(gm).size()) + 64;
        words = std::max(words, wave_mailbox_words(gm, 5, so, rows));
        ctx->cap_mailbox_words = words + 1024;
        CKC(cudaHostAlloc((void **)&ctx->h_wave_err, sizeof(int), cudaHostAllocMapped));
        *ctx->h_wave_err = 0;
        CKC(cudaHostGetDevicePointer((void **)&ctx->dm_wave_err, ctx->h_wave_err, 0));
    }
    CKC(cudaMalloc((void **)&ctx->d_tbl, sizeof(void *) * 3 * (max_batch + 1) * kSlots));
    CKC(cudaHostAlloc((void **)&ctx->h_tbl, sizeof(void *) * 3 * (max_batch + 1) * kSlots, cudaHostAllocDefault));
    CKC(alloc_cand_side(ctx, ctx->slot[0].cs));   // needs cap_pyr_floats, cap_hplane_floats, cap_ctas, cap_units, cap_mailbox_words
    for (auto &S : ctx->slot) {
        CKC(cudaHostAlloc((void **)&S.h_sums, sizeof(double) * kMaxScales * 18 * max_batch, cudaHostAllocMapped));
        CKC(cudaHostAlloc((void **)&S.h_scores, sizeof(double) * max_batch, cudaHostAllocMapped));
        CKC(cudaHostGetDevicePointer((void **)&S.dm_sums, S.h_sums, 0));
        CKC(cudaHostGetDevicePointer((void **)&S.dm_scores, S.h_scores, 0));
    }

    CKC(upload_srgb_table(ctx));
    solve_gaussian(1.5, ctx->taps, &ctx->iir);

    CKC(cudaFuncSetAttribute(k_fir_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFirSmemBytes));
    CKC(iir_configure());
#undef CKC
    *out = ctx;
    return 0;
}

int oavif_ssimu2_set_option(oavif_ssimu2_ctx *ctx, int option, int value)
{
    if (!ctx) return fail(nullptr, OAVIF_SSIMU2_E_ARG, "null context");
    if (option == OAVIF_SSIMU2_OPT_BLUR &&
        (value == OAVIF_SSIMU2_BLUR_RECURSIVE || value == OAVIF_SSIMU2_BLUR_FIR)) {
        ctx->blur_mode = value;
        return 0;
    }
    if (option == OAVIF_SSIMU2_OPT_TILE_PATH &&
        (value == OAVIF_SSIMU2_TILES_TMA || value == OAVIF_SSIMU2_TILES_CP_ASYNC || value == OAVIF_SSIMU2_TILES_FUSED ||
         value == OAVIF_SSIMU2_TILES_TMA_DECOUPLED)) {
        ctx->tile_path = value;
        return 0;
    }
    if (option == OAVIF_SSIMU2_OPT_SOURCE_ROWS &&
        (value == OAVIF_SSIMU2_SOURCE_ROWS_AT_SET_SOURCE || value == OAVIF_SSIMU2_SOURCE_ROWS_WITH_FIRST_SCORE)) {
        ctx->source_rows = value;
        return 0;
    }
    if (option == OAVIF_SSIMU2_OPT_WEIGHTS &&
        (value == OAVIF_SSIMU2_WEIGHTS_SIX_SLOTS || value == OAVIF_SSIMU2_WEIGHTS_CONTIGUOUS)) {
        ctx->weight_layout = value;
        return 0;
    }
    if (option == OAVIF_SSIMU2_OPT_VERTICAL_ORDER &&
        (value == OAVIF_SSIMU2_VERTICAL_AS_HORIZONTAL || value == OAVIF_SSIMU2_VERTICAL_FUSED_OUTER)) {
        ctx->vertical_order = value;
        return 0;
    }
    if (option == OAVIF_SSIMU2_OPT_TRANSFER && (value == OAVIF_SSIMU2_TRANSFER_F64 || value == OAVIF_SSIMU2_TRANSFER_F32)) {
        if (ctx->inflight) return fail(ctx, OAVIF_SSIMU2_E_STATE, "a submission is in flight: retire it first");
        CK(cudaSetDevice(ctx->device));
        CK(cudaDeviceSynchronize());   // pyramid kernels of earlier calls read the table
        ctx->transfer = value;
        CK(upload_srgb_table(ctx));
        ctx->have_source = false;      // the cached source pyramid was built from the other table
        ctx->src_pix = nullptr;
        return 0;
    }
    return fail(ctx, OAVIF_SSIMU2_E_ARG, "unknown option %d / value %d", option, value);
}

int oavif_ssimu2_get_option(const oavif_ssimu2_ctx *ctx, int option, int *value)
{
    if (!ctx || !value) return fail(nullptr, OAVIF_SSIMU2_E_ARG, "null argument");
    switch (option) {
    case OAVIF_SSIMU2_OPT_BLUR: *value = ctx->blur_mode; return 0;
    case OAVIF_SSIMU2_OPT_WEIGHTS: *value = ctx->weight_layout; return 0;
    case OAVIF_SSIMU2_OPT_TILE_PATH: *value = ctx->tile_path; return 0;
    case OAVIF_SSIMU2_OPT_SOURCE_ROWS: *value = ctx->source_rows; return 0;
    case OAVIF_SSIMU2_OPT_TRANSFER: *value = ctx->transfer; return 0;
    case OAVIF_SSIMU2_OPT_VERTICAL_ORDER: *value = ctx->vertical_order; return 0;
    default: return OAVIF_SSIMU2_E_ARG;
    }
}

int oavif_ssimu2_set_stream(oavif_ssimu2_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return fail(nullptr, OAVIF_SSIMU2_E_ARG, "null context");
    if (ctx->inflight) return fail(ctx, OAVIF_SSIMU2_E_STATE, "a submission is in flight: retire it first");
    ctx->user_stream = (cudaStream_t)cuda_stream;
    for (auto &S : ctx->slot) S.cs.stream = ctx->user_stream ? ctx->user_stream : S.cs.own_stream;
    return 0;
}

int oavif_ssimu2_set_source_rgb8(oavif_ssimu2_ctx *ctx, const uint8_t *rgb, uint32_t w, uint32_t h, size_t stride)
{
    return set_source_common(ctx, rgb, w, h, stride, false);
}

int oavif_ssimu2_set_source_rgb8_dev(oavif_ssimu2_ctx *ctx, const uint8_t *d_rgb, uint32_t w, uint32_t h,
                                     size_t stride)
{
    return set_source_common(ctx, d_rgb, w, h, stride, true);
}

int oavif_ssimu2_set_source_pixels(oavif_ssimu2_ctx *ctx, const void *pixels, uint32_t w, uint32_t h, size_t stride,
                                   int channels, int bits)
{
    return set_source_common(ctx, pixels, w, h, stride, false, channels, bits);
}

int oavif_ssimu2_score_pixels(oavif_ssimu2_ctx *ctx, const void *pixels, size_t stride, int channels, int bits,
                              double *score)
{
    if (channels < 1 || channels > 4 || (bits != 8 && bits != 16))
        return fail(ctx, OAVIF_SSIMU2_E_UNSUPPORTED, "UnsupportedChannelCount: %d channels, %d bits", channels, bits);
    InputDesc d{};
    d.kind = (channels == 3 && bits == 8) ? IN_RGB8 : IN_PIXELS;
    d.channels = channels;
    d.hbd = bits == 16;
    d.stride[0] = stride;
    HostPlanes hp{{pixels, nullptr, nullptr}};
    return score_common(ctx, d, 1, &hp, score);
}

int oavif_ssimu2_score_rgb8(oavif_ssimu2_ctx *ctx, const uint8_t *dist, size_t stride, double *score)
{
    InputDesc d{};
    d.kind = IN_RGB8;
    d.stride[0] = stride;
    HostPlanes hp{{dist, nullptr, nullptr}};
    return score_common(ctx, d, 1, &hp, score);
}

static int batch_rgb8(oavif_ssimu2_ctx *ctx, uint32_t n, const uint8_t *const *dists, size_t stride,
                      double *scores, bool dev)
{
    if (!ctx) return fail(nullptr, OAVIF_SSIMU2_E_ARG, "null context");
    if (!dists || n == 0 || n > ctx->max_batch) return fail(ctx, OAVIF_SSIMU2_E_ARG, "bad batch (n=%u)", n);
    InputDesc d{};
    d.kind = IN_RGB8;
    d.stride[0] = stride;
    d.on_device = dev;
    HostPlanes *hp = new (std::nothrow) HostPlanes[n];
    if (!hp) return fail(ctx, OAVIF_SSIMU2_E_NOMEM, "out of host memory");
    for (uint32_t i = 0; i < n; ++i) hp[i] = HostPlanes{{dists[i], nullptr, nullptr}};
    const int rc = score_common(ctx, d, n, hp, scores);
    delete[] hp;
    return rc;
}

int oavif_ssimu2_score_batch_rgb8(oavif_ssimu2_ctx *ctx, uint32_t n, const uint8_t *const *dists, size_t stride,
                                  double *scores)
{
    return batch_rgb8(ctx, n, dists, stride, scores, false);
}

int oavif_ssimu2_score_batch_rgb8_dev(oavif_ssimu2_ctx *ctx, uint32_t n, const uint8_t *const *d_dists,
                                      size_t stride, double *scores)
{
    return batch_rgb8(ctx, n, d_dists, stride, scores, true);
}

static int batch_yuv(oavif_ssimu2_ctx *ctx, uint32_t n, const void *const *y, const void *const *u,
                     const void *const *v, size_t ys, size_t us, size_t vs, int depth, int matrix, int rgba_path,
                     double *scores, bool dev)
{
    if (!ctx) return fail(nullptr, OAVIF_SSIMU2_E_ARG, "null context");
    if (!y || !u || !v || n == 0 || n > ctx->max_batch) return fail(ctx, OAVIF_SSIMU2_E_ARG, "bad batch (n=%u)", n);
    const int kind = kind_of(depth, rgba_path);
    if (kind < 0) return fail(ctx, OAVIF_SSIMU2_E_UNSUPPORTED, "depth %d not on the scored path (8 or 10)", depth);
    YuvK k;
    if (!yuv_consts(matrix, &k))
        return fail(ctx, OAVIF_SSIMU2_E_UNSUPPORTED, "matrix_coefficients %d not on the scored path", matrix);
    InputDesc d{};
    d.kind = kind;
    d.stride[0] = ys;
    d.stride[1] = us;
    d.stride[2] = vs;
    d.matrix = matrix;
    d.on_device = dev;
    HostPlanes *hp = new (std::nothrow) HostPlanes[n];
    if (!hp) return fail(ctx, OAVIF_SSIMU2_E_NOMEM, "out of host memory");
    for (uint32_t i = 0; i < n; ++i) hp[i] = HostPlanes{{y[i], u[i], v[i]}};
    const int rc = score_common(ctx, d, n, hp, scores);
    delete[] hp;
    return rc;
}

int oavif_ssimu2_score_yuv444(oavif_ssimu2_ctx *ctx, const void *y, const void *u, const void *v, size_t ys,
                              size_t us, size_t vs, int depth, int matrix, int rgba_path, double *score)
{
    return batch_yuv(ctx, 1, &y, &u, &v, ys, us, vs, depth, matrix, rgba_path, score, false);
}

int oavif_ssimu2_score_batch_yuv444(oavif_ssimu2_ctx *ctx, uint32_t n, const void *const *y, const void *const *u,
                                    const void *const *v, size_t ys, size_t us, size_t vs, int depth, int matrix,
                                    int rgba_path, double *scores)
{
    return batch_yuv(ctx, n, y, u, v, ys, us, vs, depth, matrix, rgba_path, scores, false);
}

int oavif_ssimu2_score_batch_yuv444_dev(oavif_ssimu2_ctx *ctx, uint32_t n, const void *const *y,
                                        const void *const *u, const void *const *v, size_t ys, size_t us,
                                        size_t vs, int depth, int matrix, int rgba_path, double *scores)
{
    return batch_yuv(ctx, n, y, u, v, ys, us, vs, depth, matrix, rgba_path, scores, true);
}

// ---- pipelined form: submit (returns at once), wait (retires the oldest) ---------------------------------------
static int submit_rgb8(oavif_ssimu2_ctx *ctx, uint32_t n, const uint8_t *const *dists, size_t stride, bool dev)
{
    if (!ctx) return fail(nullptr, OAVIF_SSIMU2_E_ARG, "null context");
    if (!dists || n == 0 || n > ctx->max_batch) return fail(ctx, OAVIF_SSIMU2_E_ARG, "bad batch (n=%u)", n);
    InputDesc d{};
    d.kind = IN_RGB8;
    d.stride[0] = stride;
    d.on_device = dev;
    std::vector<HostPlanes> hp(n);
    for (uint32_t i = 0; i < n; ++i) hp[i] = HostPlanes{{dists[i], nullptr, nullptr}};
    return submit_common(ctx, d, n, hp.data());
}

static int submit_yuv(oavif_ssimu2_ctx *ctx, uint32_t n, const void *const *y, const void *const *u, const void *const *v,
                      size_t ys, size_t us, size_t vs, int depth, int matrix, int rgba_path, bool dev)
{
    if (!ctx) return fail(nullptr, OAVIF_SSIMU2_E_ARG, "null context");
    if (!y || !u || !v || n == 0 || n > ctx->max_batch) return fail(ctx, OAVIF_SSIMU2_E_ARG, "bad batch (n=%u)", n);
    const int kind = kind_of(depth, rgba_path);
    if (kind < 0) return fail(ctx, OAVIF_SSIMU2_E_UNSUPPORTED, "depth %d not on the scored path (8 or 10)", depth);
    YuvK k;
    if (!yuv_consts(matrix, &k))
        return fail(ctx, OAVIF_SSIMU2_E_UNSUPPORTED, "matrix_coefficients %d not on the scored path", matrix);
    InputDesc d{};
    d.kind = kind;
    d.stride[0] = ys;
    d.stride[1] = us;
    d.stride[2] = vs;
    d.matrix = matrix;
    d.on_device = dev;
    std::vector<HostPlanes> hp(n);
    for (uint32_t i = 0; i < n; ++i) hp[i] = HostPlanes{{y[i], u[i], v[i]}};
    return submit_common(ctx, d, n, hp.data());
}

int oavif_ssimu2_submit_rgb8(oavif_ssimu2_ctx *ctx, uint32_t n, const uint8_t *const *dists, size_t stride)
{
    return submit_rgb8(ctx, n, dists, stride, false);
}

int oavif_ssimu2_submit_rgb8_dev(oavif_ssimu2_ctx *ctx, uint32_t n, const uint8_t *const *d_dists, size_t stride)
{
    return submit_rgb8(ctx, n, d_dists, stride, true);
}

int oavif_ssimu2_submit_yuv444(oavif_ssimu2_ctx *ctx, uint32_t n, const void *const *y, const void *const *u,
                               const void *const *v, size_t ys, size_t us, size_t vs, int depth, int matrix,
                               int rgba_path)
{
    return submit_yuv(ctx, n, y, u, v, ys, us, vs, depth, matrix, rgba_path, false);
}

int oavif_ssimu2_submit_yuv444_dev(oavif_ssimu2_ctx *ctx, uint32_t n, const void *const *y, const void *const *u,
                                   const void *const *v, size_t ys, size_t us, size_t vs, int depth, int matrix,
                                   int rgba_path)
{
    return submit_yuv(ctx, n, y, u, v, ys, us, vs, depth, matrix, rgba_path, true);
}

int oavif_ssimu2_wait(oavif_ssimu2_ctx *ctx, double *scores) { return wait_common(ctx, scores); }

int oavif_ssimu2_in_flight(const oavif_ssimu2_ctx *ctx) { return ctx ? ctx->inflight : 0; }

int oavif_ssimu2_compute_rgb8(const uint8_t *ref, const uint8_t *dist, uint32_t w, uint32_t h, uint32_t channels,
                              double *score)
{
    // The reference's stateless call (tq.zig:37).  One cached context per process on the default device,
    // regrown on demand, released by oavif_ssimu2_release_cached().
    if (!ref || !dist || !score) return fail(nullptr, OAVIF_SSIMU2_E_ARG, "null argument");
    if (channels != 3) return fail(nullptr, OAVIF_SSIMU2_E_UNSUPPORTED, "channels = %u (oavif always passes 3)", channels);
    if (w == 0 || h == 0) return fail(nullptr, OAVIF_SSIMU2_E_ARG, "zero image dimension");
    std::lock_guard<std::mutex> lock(g_cached_mu);
    oavif_ssimu2_ctx *&cached = g_cached;
    if (cached && (cached->device != g_default_device || !fits_capacity(cached, w, h))) {
        oavif_ssimu2_ctx_destroy(cached);
        cached = nullptr;
    }
    if (!cached) {
        const int rc = oavif_ssimu2_ctx_create(g_default_device, w, h, 1, &cached);
        if (rc) return rc;
    }
    int rc = oavif_ssimu2_set_source_rgb8(cached, ref, w, h, (size_t)3 * w);
    if (rc == 0) rc = oavif_ssimu2_score_rgb8(cached, dist, (size_t)3 * w, score);
    if (rc) t_last_error = cached->err;
    return rc;
}

int oavif_ssimu2_set_default_device(int device)
{
    if (device < 0) return fail(nullptr, OAVIF_SSIMU2_E_ARG, "bad device %d", device);
    std::lock_guard<std::mutex> lock(g_cached_mu);
    g_default_device = device;
    return 0;
}

void oavif_ssimu2_release_cached(void)
{
    std::lock_guard<std::mutex> lock(g_cached_mu);
    if (g_cached) oavif_ssimu2_ctx_destroy(g_cached);
    g_cached = nullptr;
}

int oavif_ssimu2_yuv444_to_rgb8(oavif_ssimu2_ctx *ctx, const void *y, const void *u, const void *v, size_t ys,
                                size_t us, size_t vs, uint32_t w, uint32_t h, int depth, int matrix, int rgba_path,
                                uint8_t *rgb_out)
{
    if (!ctx) return fail(nullptr, OAVIF_SSIMU2_E_ARG, "null context");
    if (!y || !u || !v || !rgb_out) return fail(ctx, OAVIF_SSIMU2_E_ARG, "null argument");
    const int kind = kind_of(depth, rgba_path);
    if (kind < 0) return fail(ctx, OAVIF_SSIMU2_E_UNSUPPORTED, "depth %d not on the scored path (8 or 10)", depth);
    Yuv2RgbArgs a{};
    if (!yuv_consts(matrix, &a.k))
        return fail(ctx, OAVIF_SSIMU2_E_UNSUPPORTED, "matrix_coefficients %d not on the scored path", matrix);
    if (w == 0 || h == 0 || (long long)w * h * 8 + 768 > ctx->cap_in_bytes)
        return fail(ctx, OAVIF_SSIMU2_E_STATE, "image %ux%u exceeds context capacity", w, h);
    InputDesc dd{};
    dd.kind = kind;
    const size_t rb = row_bytes(dd, (int)w);
    if (ys < rb || us < rb || vs < rb) return fail(ctx, OAVIF_SSIMU2_E_ARG, "stride smaller than a row");
    CK(cudaSetDevice(ctx->device));
    const size_t plane_bytes = (rb * h + 255) & ~(size_t)255;
    if (ctx->inflight) return fail(ctx, OAVIF_SSIMU2_E_STATE, "a submission is in flight: retire it first");
    uint8_t *base = ctx->slot[0].d_in_dist;
    const void *src[3] = {y, u, v};
    const size_t st[3] = {ys, us, vs};
    for (int p = 0; p < 3; ++p)
        CK(cudaMemcpy2DAsync(base + p * plane_bytes, rb, src[p], st[p], rb, h, cudaMemcpyHostToDevice, ctx->cs->stream));
    a.y = base;
    a.u = base + plane_bytes;
    a.v = base + 2 * plane_bytes;
    a.stride[0] = a.stride[1] = a.stride[2] = (long long)rb;
    a.w = (int)w;
    a.h = (int)h;
    // RGB8 goes to a buffer of its own (grown on demand): the staged source and its cached pyramid stay valid
    const size_t out_bytes = (size_t)w * h * 3;
    if (out_bytes > ctx->conv_bytes) {
        cudaFree(ctx->d_conv);
        ctx->d_conv = nullptr;
        ctx->conv_bytes = 0;
        CK(cudaMalloc(&ctx->d_conv, out_bytes));
        ctx->conv_bytes = out_bytes;
    }
    a.out = ctx->d_conv;
    const dim3 grid(cdiv((int)w, 256), h);
    if (kind == IN_YUV8) k_yuv_to_rgb8<IN_YUV8><<<grid, 256, 0, ctx->cs->stream>>>(a);
    else if (kind == IN_YUV10_RGB) k_yuv_to_rgb8<IN_YUV10_RGB><<<grid, 256, 0, ctx->cs->stream>>>(a);
    else k_yuv_to_rgb8<IN_YUV10_RGBA><<<grid, 256, 0, ctx->cs->stream>>>(a);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(rgb_out, a.out, (size_t)w * h * 3, cudaMemcpyDeviceToHost, ctx->cs->stream));
    CK(cudaStreamSynchronize(ctx->cs->stream));
    return 0;
}

int oavif_ssimu2_source_samples(oavif_ssimu2_ctx *ctx, int out_depth, void *out, size_t out_bytes)
{
    if (!ctx) return fail(nullptr, OAVIF_SSIMU2_E_ARG, "null context");
    if (!out) return fail(ctx, OAVIF_SSIMU2_E_ARG, "null argument");
    if (!ctx->have_source || !ctx->src_pix)
        return fail(ctx, OAVIF_SSIMU2_E_STATE, "no source pixels on the device (set_source_* first; images below 8x8 are not staged)");
    const int bits = ctx->src_pix_bits;
    int mode;
    if (bits == 8 && out_depth == 10) mode = 0;
    else if (bits == 16 && out_depth == 10) mode = 1;
    else if (bits == 16 && out_depth == 8) mode = 2;
    else   // 8 -> 8: the reference hands src.data through untouched (io.zig:611-613)
        return fail(ctx, OAVIF_SSIMU2_E_UNSUPPORTED, "no conversion from %d-bit source samples to depth %d", bits, out_depth);
    const int w = ctx->g.w[0], h = ctx->g.h[0];
    SamplesArgs a{};
    a.row_samples = w * ctx->src_pix_channels;
    a.h = h;
    const size_t need = (size_t)a.row_samples * h * (out_depth > 8 ? 2 : 1);
    if (out_bytes < need) return fail(ctx, OAVIF_SSIMU2_E_ARG, "output holds %zu bytes, %zu needed", out_bytes, need);
    CK(cudaSetDevice(ctx->device));
    if (need > ctx->conv_bytes) {
        cudaFree(ctx->d_conv);
        ctx->d_conv = nullptr;
        ctx->conv_bytes = 0;
        CK(cudaMalloc(&ctx->d_conv, need));
        ctx->conv_bytes = need;
    }
    a.in = ctx->src_pix;
    a.out = ctx->d_conv;
    a.stride = (long long)ctx->src_pix_stride;
    // on the source stream, behind the upload and the pyramid kernel that read the same staged pixels
    cudaStream_t ss = ctx->src_stream;
    const dim3 grid(cdiv(cdiv(a.row_samples, 4), 256), h);
    if (mode == 0) k_source_samples<0><<<grid, 256, 0, ss>>>(a);
    else if (mode == 1) k_source_samples<1><<<grid, 256, 0, ss>>>(a);
    else k_source_samples<2><<<grid, 256, 0, ss>>>(a);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->d_conv, need, cudaMemcpyDeviceToHost, ss));
    CK(cudaStreamSynchronize(ss));
    return 0;
}

int oavif_ssimu2_get_detail(oavif_ssimu2_ctx *ctx, uint32_t candidate, oavif_ssimu2_detail *out)
{
    if (!ctx || !out) return fail(ctx, OAVIF_SSIMU2_E_ARG, "null argument");
    const Slot &S = ctx->slot[ctx->last_slot];   // the submission retired last
    if (candidate >= S.n) return fail(ctx, OAVIF_SSIMU2_E_STATE, "candidate %u not in the last call", candidate);
    memset(out, 0, sizeof *out);
    out->n_scales = S.n_scales;
    for (int s = 0; s < S.n_scales; ++s) {
        out->w[s] = S.w[s];
        out->h[s] = S.h[s];
        for (int i = 0; i < 18; ++i) out->sums[s][i] = S.h_sums[((size_t)candidate * kMaxScales + s) * 18 + i];
    }
    out->score = S.h_scores[candidate];
    return 0;
}

int oavif_ssimu2_get_timing(oavif_ssimu2_ctx *ctx, oavif_ssimu2_timing *out)
{
    if (!ctx || !out) return fail(ctx, OAVIF_SSIMU2_E_ARG, "null argument");
    *out = ctx->timing;
    return 0;
}

int oavif_ssimu2_debug_get_xyb(oavif_ssimu2_ctx *ctx, int which, int scale, int channel, float *out,
                               uint32_t *w_out, uint32_t *h_out)
{
    if (!ctx || !out || !w_out || !h_out) return fail(ctx, OAVIF_SSIMU2_E_ARG, "null argument");
    if (!ctx->have_source || scale < 0 || scale >= ctx->g.n_scales || channel < 0 || channel > 2 || which < 0 ||
        which > (int)ctx->max_batch)
        return fail(ctx, OAVIF_SSIMU2_E_STATE, "no such plane");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->src_stream));  // the source pyramid may still be in flight
    CK(cudaStreamSynchronize(ctx->cs->stream));
    const Geom &g = ctx->g;
    const float *base = which == 0 ? ctx->src[ctx->cur].d_pyr : ctx->cs->d_dist_pyr + (long long)(which - 1) * ctx->cap_pyr_floats;
    const float *p = base + g.off[scale] + (long long)channel * g.plane[scale];
    CK(cudaMemcpy2D(out, sizeof(float) * g.w[scale], p, sizeof(float) * g.pitch[scale], sizeof(float) * g.w[scale],
                    g.h[scale], cudaMemcpyDeviceToHost));
    *w_out = (uint32_t)g.w[scale];
    *h_out = (uint32_t)g.h[scale];
    return 0;
}

int oavif_ssimu2_debug_get_rows(oavif_ssimu2_ctx *ctx, int candidate, int quantity, int scale, int channel, float *out,
                                uint32_t *w_out, uint32_t *h_out)
{
    if (!ctx || !out || !w_out || !h_out) return fail(ctx, OAVIF_SSIMU2_E_ARG, "null argument");
    if (!ctx->have_source || ctx->last_n == 0 || ctx->blur_mode != OAVIF_SSIMU2_BLUR_RECURSIVE || !ctx->src[ctx->cur].rows_valid ||
        scale < 0 || scale >= ctx->g.n_scales || channel < 0 || channel > 2 || quantity < 0 || quantity > 4 ||
        candidate < 0 || candidate >= (int)ctx->last_n)
        return fail(ctx, OAVIF_SSIMU2_E_STATE, "no such row-filtered plane (needs a RECURSIVE score call first)");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->cs->stream));
    const Geom &g = ctx->g;
    const long long P = ctx->cap_pyr_floats;
    const long long poff = g.off[scale] + (long long)channel * g.plane[scale];
    const int w = g.w[scale], h = g.h[scale], pitch = g.pitch[scale];
    // quantities 0..4 = a, b, a*a, b*b, a*b.  (a, a*a) and (b, b*b) are interleaved pair planes, a*b is plain.
    const float *p;
    size_t spitch, elem;
    if (quantity == 4) {
        p = ctx->cs->d_hplanes + (long long)candidate * 3 * P + 2 * P + poff;
        spitch = sizeof(float) * pitch;
        elem = sizeof(float);
    } else {
        const float *base = (quantity & 1) ? ctx->cs->d_hplanes + (long long)candidate * 3 * P : ctx->src[ctx->cur].d_hplanes;
        p = base + 2 * poff + (quantity >> 1);
        spitch = sizeof(float) * 2 * pitch;
        elem = 2 * sizeof(float);
    }
    if (elem == sizeof(float)) {
        CK(cudaMemcpy2D(out, sizeof(float) * w, p, spitch, sizeof(float) * w, h, cudaMemcpyDeviceToHost));
    } else {
        // one float out of every float2: copy the interleaved rows and pick on the host
        std::vector<float> tmp((size_t)2 * w * h);
        CK(cudaMemcpy2D(tmp.data(), sizeof(float) * 2 * w, p - (quantity >> 1), spitch, sizeof(float) * 2 * w, h,
                        cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < (size_t)w * h; ++i) out[i] = tmp[2 * i + (quantity >> 1)];
    }
    *w_out = (uint32_t)w;
    *h_out = (uint32_t)h;
    return 0;
}

int oavif_ssimu2_debug_get_cols(oavif_ssimu2_ctx *ctx, int candidate, int scale, int channel, float *out,
                                uint32_t *w_out, uint32_t *h_out)
{
    if (!ctx || !out || !w_out || !h_out) return fail(ctx, OAVIF_SSIMU2_E_ARG, "null argument");
    if (!ctx->have_source || ctx->last_n == 0 || ctx->blur_mode != OAVIF_SSIMU2_BLUR_RECURSIVE || !(ctx->src[ctx->cur].rows_valid || ctx->src[ctx->cur].musig_valid) ||
        scale < 0 || scale >= ctx->g.n_scales || channel < 0 || channel > 2 || candidate < 0 ||
        candidate >= (int)ctx->last_n)
        return fail(ctx, OAVIF_SSIMU2_E_STATE, "no such blurred plane (needs a RECURSIVE score call first)");
    if (ctx->inflight) return fail(ctx, OAVIF_SSIMU2_E_STATE, "a submission is in flight: retire it first");
    CK(cudaSetDevice(ctx->device));
    const Geom &g = ctx->g;
    const int w = g.w[scale], h = g.h[scale];
    const long long need = 5LL * w * h;
    if (need > ctx->dbg_floats) {
        cudaFree(ctx->d_dbg);
        ctx->d_dbg = nullptr;
        ctx->dbg_floats = 0;
        CK(cudaMalloc(&ctx->d_dbg, sizeof(float) * need));
        ctx->dbg_floats = need;
    }
    // the scored path's own columns kernel once more over the row-filtered planes the last call left, with the
    // tap on; the pooled sums it rewrites are the ones already there (same inputs, fixed order)
    BlurPlan plan;
    plan_iir_v(g, &plan);
    const IirDebugTap tap{ctx->d_dbg, scale, channel, candidate};
    if (ctx->tile_path == OAVIF_SSIMU2_TILES_FUSED) {   // the fused kernel's own tap instance, cached-source form
        const int rc = enqueue_wave(ctx, ctx->src[ctx->cur], plan, 1, 0, ctx->last_n, &tap);
        if (rc) return rc;
    } else {
        IirColsTmaMaps cmaps;
        const cudaError_t e = launch_iir_cols(iir_args_for(ctx, ctx->src[ctx->cur]), plan.first_cta, plan.tiles_x,
                                              (int)ctx->last_n, ctx->cs->stream, &tap, cols_maps_for(ctx, ctx->src[ctx->cur], &cmaps),
                                              ctx->tile_path == OAVIF_SSIMU2_TILES_TMA_DECOUPLED, 0,
                                              ctx->vertical_order == OAVIF_SSIMU2_VERTICAL_FUSED_OUTER);
        if (e != cudaSuccess) return fail(ctx, OAVIF_SSIMU2_E_CUDA, "columns launch: %s", cudaGetErrorString(e));
    }
    CK(cudaMemcpyAsync(out, ctx->d_dbg, sizeof(float) * need, cudaMemcpyDeviceToHost, ctx->cs->stream));
    CK(cudaStreamSynchronize(ctx->cs->stream));
    *w_out = (uint32_t)w;
    *h_out = (uint32_t)h;
    return 0;
}

int oavif_ssimu2_debug_wave_trace(oavif_ssimu2_ctx *ctx, int mode, uint64_t *out, uint32_t cap_units, uint32_t *n_units)
{
    if (!ctx || !out || !n_units) return fail(ctx, OAVIF_SSIMU2_E_ARG, "null argument");
    if (!ctx->have_source || ctx->last_n == 0 || ctx->g.n_scales == 0 || ctx->inflight)
        return fail(ctx, OAVIF_SSIMU2_E_STATE, "needs a previous score call and nothing in flight");
    SrcSet &Src = ctx->src[ctx->cur];
    if (mode == 1 && !Src.musig_valid) return fail(ctx, OAVIF_SSIMU2_E_STATE, "no cached source blur: score in FUSED mode first");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->src_stream));
    BlurPlan plan;
    plan_iir_v(ctx->g, &plan);
    unsigned long long *d_trace = nullptr;
    CK(cudaMalloc(&d_trace, sizeof(unsigned long long) * 5 * (ctx->cap_units + 64)));
    CK(cudaMemsetAsync(d_trace, 0, sizeof(unsigned long long) * 5 * (ctx->cap_units + 64), ctx->cs->stream));
    ctx->trace_buf = d_trace;
    const IirDebugTap tap{nullptr, -1, -1, -1};
    const int rc = enqueue_wave(ctx, Src, plan, mode == 2 ? 2 : 1, 0, 1, &tap);
    ctx->trace_buf = nullptr;
    if (rc) {
        cudaFree(d_trace);
        return rc;
    }
    CK(cudaStreamSynchronize(ctx->cs->stream));
    *n_units = ctx->cs->n_units;
    const uint32_t n = std::min(cap_units, ctx->cs->n_units);
    CK(cudaMemcpy(out, d_trace, sizeof(unsigned long long) * 5 * n, cudaMemcpyDeviceToHost));
    cudaFree(d_trace);
    return 0;
}

int oavif_ssimu2_debug_blur(oavif_ssimu2_ctx *ctx, const float *in, uint32_t w, uint32_t h, float *out)
{
    if (!ctx || !in || !out || w == 0 || h == 0) return fail(ctx, OAVIF_SSIMU2_E_ARG, "null argument");
    CK(cudaSetDevice(ctx->device));
    const int pitch = rup((int)w, 32);
    const long long need = 3LL * pitch * h;
    if (need > ctx->dbg_floats) {
        cudaFree(ctx->d_dbg);
        ctx->d_dbg = nullptr;
        ctx->dbg_floats = 0;
        CK(cudaMalloc(&ctx->d_dbg, sizeof(float) * need));
        ctx->dbg_floats = need;
    }
    float *d_in = ctx->d_dbg, *d_tmp = d_in + (long long)pitch * h, *d_out = d_tmp + (long long)pitch * h;
    CK(cudaMemcpy2DAsync(d_in, sizeof(float) * pitch, in, sizeof(float) * w, sizeof(float) * w, h,
                         cudaMemcpyHostToDevice, ctx->cs->stream));
    const cudaError_t e = launch_debug_blur(ctx->blur_mode == OAVIF_SSIMU2_BLUR_FIR, ctx->taps, ctx->iir, d_in,
                                            d_tmp, d_out, (int)w, (int)h, pitch, ctx->cs->stream);
    if (e != cudaSuccess) return fail(ctx, OAVIF_SSIMU2_E_CUDA, "debug blur launch: %s", cudaGetErrorString(e));
    CK(cudaMemcpy2DAsync(out, sizeof(float) * w, d_out, sizeof(float) * pitch, sizeof(float) * w, h,
                         cudaMemcpyDeviceToHost, ctx->cs->stream));
    CK(cudaStreamSynchronize(ctx->cs->stream));
    return 0;
}

int oavif_ssimu2_debug_check_guards(oavif_ssimu2_ctx *ctx)
{
    if (!ctx) return fail(nullptr, OAVIF_SSIMU2_E_ARG, "null context");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    std::vector<unsigned char> host(kGuardBytes);
    int idx = 0;
    for (const auto &g : ctx->guards) {
        CK(cudaMemcpy(host.data(), g.first, g.second, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < g.second; ++i)
            if (host[i] != kGuardByte)
                return fail(ctx, OAVIF_SSIMU2_E_STATE, "guard band %d overwritten at byte %zu", idx, i);
        ++idx;
    }
    return 0;
}

int oavif_ssimu2_debug_time_rows(oavif_ssimu2_ctx *ctx, int variant, int iters, float *mean_ms)
{
    if (!ctx || !mean_ms || iters <= 0) return fail(ctx, OAVIF_SSIMU2_E_ARG, "bad argument");
    if (!ctx->have_source || ctx->last_n == 0 || ctx->g.n_scales == 0)
        return fail(ctx, OAVIF_SSIMU2_E_STATE, "needs a previous score call");
    CK(cudaSetDevice(ctx->device));
    BlurPlan plan;
    plan_iir_v(ctx->g, &plan);
    if (ctx->inflight) return fail(ctx, OAVIF_SSIMU2_E_STATE, "a submission is in flight: retire it first");
    SrcSet &Src = ctx->src[ctx->cur];
    IirRowsTmaMaps maps;
    const bool tma = !(variant & 8) && rows_maps_for(ctx, Src, &maps);
    const int which = (variant & 128) ? 3 : (variant & 4) ? 1 : 2;
    if (which == 3 && !tma) return fail(ctx, OAVIF_SSIMU2_E_ARG, "the source-only rows kernel exists in the TMA form only");
    if ((variant & 2048) && !Src.musig_valid) return fail(ctx, OAVIF_SSIMU2_E_STATE, "no cached source blur: score in FUSED mode first");
    CK(cudaStreamSynchronize(ctx->src_stream));
    CK(cudaEventRecord(ctx->src_up0, ctx->cs->stream));
    for (int i = 0; i < iters; ++i) {
        cudaError_t e;
        if (variant & (1024 | 2048)) {   // the fused kernel: all five quantities / the candidate's three (cached source blur)
            BlurPlan cp;
            plan_iir_v(ctx->g, &cp);
            const int rc = enqueue_wave(ctx, Src, cp, (variant & 1024) ? 2 : 1, 0, 1);
            if (rc) return rc;
            e = cudaSuccess;
        } else if (variant & 256) { // the two halves as the product issues them: source stream next to compute stream
            if (!tma) return fail(ctx, OAVIF_SSIMU2_E_ARG, "the split rows pass exists in the TMA form only");
            CK(cudaEventRecord(ctx->ev_user, ctx->cs->stream));
            CK(cudaStreamWaitEvent(ctx->src_stream, ctx->ev_user, 0));
            e = launch_iir_rows(iir_args_for(ctx, Src), ctx->g, 3, 1, ctx->src_stream, &maps);
            if (e == cudaSuccess) e = launch_iir_rows(iir_args_for(ctx, Src), ctx->g, 1, 1, ctx->cs->stream, &maps);
            CK(cudaEventRecord(Src.ready, ctx->src_stream));
            CK(cudaStreamWaitEvent(ctx->cs->stream, Src.ready, 0));
        } else if (variant & 512) { // the columns pass alone
            BlurPlan cp;
            plan_iir_v(ctx->g, &cp);
            IirColsTmaMaps cmaps;
            IirArgs ca = iir_args_for(ctx, Src);
            const IirColsTmaMaps *cm = (variant & 8) ? nullptr : cols_maps_for(ctx, Src, &cmaps);
            if (cm && (variant & (16384 | 32768))) {   // layout / promotion experiment: durations only
                const long long P = ctx->cap_pyr_floats;
                if (!iir_cols_tma_maps_experiment(&cmaps, ctx->g, Src.d_hplanes, Src.d_pyr, ctx->cs->d_hplanes, ctx->cs->d_hplanes + 2 * P,
                                                  ctx->cs->d_dist_pyr, (variant & 16384) != 0, (variant & 32768) != 0))
                    return fail(ctx, OAVIF_SSIMU2_E_CUDA, "experiment descriptors");
                ca.dbg_strip_major = (variant & 16384) ? 1 : 0;
            }
            e = launch_iir_cols(ca, cp.first_cta, cp.tiles_x, 1, ctx->cs->stream, nullptr, cm, (variant & 4096) != 0,
                                (variant & 8192) ? 40 * 1024 : 0);
        } else {
            e = launch_iir_rows(iir_args_for(ctx, Src), ctx->g, which, 1, ctx->cs->stream, tma ? &maps : nullptr, (variant >> 4) & 7);
        }
        if (e != cudaSuccess) return fail(ctx, OAVIF_SSIMU2_E_CUDA, "rows launch: %s", cudaGetErrorString(e));
    }
    CK(cudaEventRecord(ctx->src_up1, ctx->cs->stream));
    CK(cudaStreamSynchronize(ctx->cs->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->src_up0, ctx->src_up1));
    *mean_ms = ms / iters;
    return 0;
}

}  // extern "C"
