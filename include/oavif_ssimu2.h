/*
 * include/oavif_ssimu2.h — C ABI of the B200 (sm_100a) SSIMULACRA2 scorer for oavif's
 * target-quality search loop.
 *
 * What it replaces (all paths under /root/reference/):
 *   src/tq.zig:37        fssimu2.computeSsimu2(allocator, e.rgb, decoded_rgb, e.w, e.h, 3, null)
 *                        -> oavif_ssimu2_compute_rgb8()              (stateless, 1:1)
 *                        -> oavif_ssimu2_set_source_rgb8() once per image (main.zig:86 e.rgb)
 *                           + oavif_ssimu2_score_*() once per search pass (tq.zig:150)
 *   src/io.zig:452-482   decodeAvifCommon: avifImageYUVToRGB at forced depth 8
 *   src/io.zig:638-666   decodeAvifToRgb: per-pixel repack to tight RGB8
 *                        -> oavif_ssimu2_score_yuv444(): takes decoder->image->yuvPlanes
 *                           straight after avifDecoderNextImage (io.zig:463) and performs the
 *                           same integer YUV->RGB8 arithmetic on the device
 *                        -> oavif_ssimu2_yuv444_to_rgb8(): that conversion alone
 *   tq.zig:135-181       (new, additive) batched probing: oavif_ssimu2_score_batch_*()
 *
 * Conventions: plain pointers and sizes only; every entry point returns 0 on success or a
 * negative OAVIF_SSIMU2_E_* code (the Zig shim maps them onto an error set, as `try` at
 * tq.zig:37 expects); oavif_ssimu2_last_error() gives the text.  A context is single-owner:
 * one host thread at a time (the corpus driver creates one per GPU worker).  Calls are
 * synchronous at this boundary — the callee has finished reading caller memory on return,
 * which is what tq.zig:26-27 (`defer allocator.free(decoded_rgb)`) requires.
 * There is NO CPU fallback: without a usable CUDA device every call fails with
 * OAVIF_SSIMU2_E_CUDA.
 */
#ifndef OAVIF_SSIMU2_H
#define OAVIF_SSIMU2_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OAVIF_SSIMU2_ABI_VERSION 1
#define OAVIF_SSIMU2_MAX_SCALES 6

enum {
    OAVIF_SSIMU2_OK = 0,
    OAVIF_SSIMU2_E_ARG = -1,      /* null pointer, zero size, stride too small, bad enum       */
    OAVIF_SSIMU2_E_CUDA = -2,     /* CUDA runtime/driver error (no device, launch failure ...) */
    OAVIF_SSIMU2_E_NOMEM = -3,    /* device or pinned allocation failed                        */
    OAVIF_SSIMU2_E_STATE = -4,    /* score_* before set_source_*, or size beyond ctx capacity  */
    OAVIF_SSIMU2_E_UNSUPPORTED = -5 /* matrix / depth / channel count outside the scored path  */
};

/* Gaussian-blur evaluation.  Both compute the sigma = 1.5 filter of SSIMULACRA2 v2.1.
 *   RECURSIVE: the published 3-oscillator recursion in binary32, every row and column as one
 *              serial chain (columns in parallel; rows through a shared-memory transpose).
 *              Its round-off is part of the published result, so this is the default.
 *   FIR:       the exactly equivalent 9-tap kernel in one tile-fused launch per scale set
 *              (no intermediate planes); differs from RECURSIVE only by that round-off.      */
enum { OAVIF_SSIMU2_BLUR_RECURSIVE = 0, OAVIF_SSIMU2_BLUR_FIR = 1 };

/* How the 108 weights of the final sum are laid over images with FEWER than six scales (min side < ~256; every
 * BASELINE config has six, where both readings coincide).  Nothing in /root/reference settles which one
 * fssimu2 0.1.1 uses (its source is un-vendored, build.zig.zon:7-10), so both are offered:
 *   SIX_SLOTS:  weight index ((c*6 + scale)*2 + n)*3 + k, absent scales contribute zero (default; the
 *               fixed six-scale table of the vszip lineage fssimu2 derives from)
 *   CONTIGUOUS: libjxl tools/ssimulacra2.cc Msssim::Score(): `for scale < scales.size()` with a running
 *               i++, i.e. index ((c*n_scales + scale)*2 + n)*3 + k                               */
enum { OAVIF_SSIMU2_WEIGHTS_SIX_SLOTS = 0, OAVIF_SSIMU2_WEIGHTS_CONTIGUOUS = 1 };

/* How the RECURSIVE kernels move their tiles.  TMA (default): cp.async.bulk.tensor loads and stores issued by one
 * elected lane, completion on mbarriers, zero padding and edge clipping done by the hardware.  CP_ASYNC: the
 * round-1 kernels (dedicated loader / storer warps, 16-byte cp.async, one block barrier per chunk), kept for A/B
 * measurements and for drivers without a tensor-map encoder.  FUSED: ONE kernel for both passes, the maps and the
 * pooling (ssimu2_wave.cuh): column strips run as a wavefront, the six-float state of every row chain is handed from
 * a strip to its right neighbour through an L2-resident mailbox, and the row-filtered planes never reach HBM; the
 * per-source cache then holds the fully blurred (mu1, sigma11) instead of the rows pass of (a, a*a).
 * TMA_DECOUPLED: TMA, with the columns kernel's per-batch block barrier replaced by one mbarrier per hand-over
 * (rows landed / batch produced / batch consumed); measured equal to TMA, kept for A/B (profiles/r2_cols_sync_forms.txt).
 * Same arithmetic, same bits in all four. */
enum { OAVIF_SSIMU2_TILES_TMA = 0, OAVIF_SSIMU2_TILES_CP_ASYNC = 1, OAVIF_SSIMU2_TILES_FUSED = 2,
       OAVIF_SSIMU2_TILES_TMA_DECOUPLED = 3 };

/* When the rows pass of the source-only quantities (a, a*a) runs (RECURSIVE blur, TMA kernels).  WITH_FIRST_SCORE
 * (default): carried by the first scoring call's rows kernel (candidate 0's CTAs run a second pair warp), cached
 * for every later candidate of the source.  AT_SET_SOURCE: a launch of its own on the source stream, enqueued by
 * set_source next to the source's pyramid.  Measured equal for pipelined callers (0.518 ms per 4K evaluation either
 * way) and 5 % slower for synchronous ones (the lone source half is bound by its one recursion warp per CTA), hence
 * not the default; useful when set_source happens long before the first score. */
enum { OAVIF_SSIMU2_SOURCE_ROWS_AT_SET_SOURCE = 0, OAVIF_SSIMU2_SOURCE_ROWS_WITH_FIRST_SCORE = 1 };

/* How the 256-entry sRGB -> linear table is made.  F64 (default): the transfer function in binary64, rounded once to
 * binary32.  F32: evaluated in binary32 (powf) — a few entries differ in their last bit, which moves scores by up to
 * 0.05 (profiles/r2_variant_envelope.json); which one fssimu2 0.1.1 uses is not known here, so
 * both exist, each bit-identical to the oracle's matching variant.  Setting it (nothing may be in flight) drops the
 * cached source: call set_source_* again. */
enum { OAVIF_SSIMU2_TRANSFER_F64 = 0, OAVIF_SSIMU2_TRANSFER_F32 = 1 };

/* Operation order of the recursion step in the VERTICAL pass (RECURSIVE blur).  AS_HORIZONTAL (default): the same
 * sequence as the horizontal pass, (n2 * sum - y[n-2]) then fma(-d1, y[n-1], .).  FUSED_OUTER: fma(n2, sum,
 * fma(-d1, y[n-1], -y[n-2])), the vertical block of lib/jxl/gauss_blur.cc as recalled — the oracle's
 * ORACLE_VARIANT_VERTICAL_ORDER, up to 0.058 away in the score (profiles/r2_variant_envelope.json).  Which one fssimu2
 * 0.1.1 follows is not known here; both are bit-identical to the oracle's matching variant.  FUSED_OUTER exists for
 * the default tile path only (score_* return E_UNSUPPORTED under CP_ASYNC, FUSED and TMA_DECOUPLED). */
enum { OAVIF_SSIMU2_VERTICAL_AS_HORIZONTAL = 0, OAVIF_SSIMU2_VERTICAL_FUSED_OUTER = 1 };

enum { OAVIF_SSIMU2_OPT_BLUR = 1, OAVIF_SSIMU2_OPT_WEIGHTS = 2, OAVIF_SSIMU2_OPT_TILE_PATH = 3, OAVIF_SSIMU2_OPT_SOURCE_ROWS = 4,
       OAVIF_SSIMU2_OPT_TRANSFER = 5, OAVIF_SSIMU2_OPT_VERTICAL_ORDER = 6 };

typedef struct oavif_ssimu2_ctx oavif_ssimu2_ctx;

/* Pooled results of the last score call for one candidate (tests, trace tools). */
typedef struct {
    int32_t n_scales;
    int32_t w[OAVIF_SSIMU2_MAX_SCALES];
    int32_t h[OAVIF_SSIMU2_MAX_SCALES];
    /* [scale][c*6 + {sum d, sum d^4, sum artifact, sum artifact^4, sum detail, sum detail^4}] */
    double sums[OAVIF_SSIMU2_MAX_SCALES][18];
    double score;
} oavif_ssimu2_detail;

/* Device-side milliseconds of the last call, from CUDA events on the context's stream. */
typedef struct {
    float h2d_ms;       /* host -> device copies of this call's pixels           */
    float pyramid_ms;   /* YUV->RGB8, sRGB->linear, 2x pyramid, XYB              */
    float blur_ms;      /* blur + error maps + pooling kernels (a + b)           */
    float blur_a_ms;    /* RECURSIVE: rows pass (b, b*b, a*b; + a, a*a on the first call after set_source). FIR: the fused kernel */
    float blur_b_ms;    /* RECURSIVE: columns pass + maps + pooling.  FIR: 0             */
    float finalize_ms;  /* fixed-order reduction, weights, score, D2H of scores  */
    float total_ms;     /* first event to last event                             */
    uint32_t launches;  /* kernels of this library launched by the call          */
} oavif_ssimu2_timing;

/* ---- lifetime ------------------------------------------------------------------------ */

int oavif_ssimu2_abi_version(void);

/* device: CUDA ordinal.  max_w/max_h: largest image the context will see.  max_batch: most
 * candidates scored per call (>= 1).  All device and pinned memory is allocated here; no
 * allocation happens in score calls. */
int oavif_ssimu2_ctx_create(int device, uint32_t max_w, uint32_t max_h, uint32_t max_batch,
                            oavif_ssimu2_ctx **out);
void oavif_ssimu2_ctx_destroy(oavif_ssimu2_ctx *ctx);

int oavif_ssimu2_set_option(oavif_ssimu2_ctx *ctx, int option, int value);
int oavif_ssimu2_get_option(const oavif_ssimu2_ctx *ctx, int option, int *value);

/* Launch on a caller-owned cudaStream_t (passed as void*) instead of the context's own. */
int oavif_ssimu2_set_stream(oavif_ssimu2_ctx *ctx, void *cuda_stream);

const char *oavif_ssimu2_last_error(const oavif_ssimu2_ctx *ctx); /* ctx may be NULL */

/* "0000:c0:00.0" of a CUDA device, lower case as sysfs spells it: /sys/bus/pci/devices/<id>/local_cpulist and
 * numa_node tell a caller where to run (and first-touch its staging) to be next to that GPU's PCIe root. */
int oavif_ssimu2_device_pci_bus_id(int device, char *out, size_t cap);

/* Pinned host staging for src/io.zig's decode buffers (cudaHostAlloc / cudaFreeHost).  The pages are pinned
 * where the CALLING thread runs: bind the thread next to the GPU first (see above). */
void *oavif_ssimu2_pinned_alloc(size_t bytes);
void oavif_ssimu2_pinned_free(void *p);

/* ---- source side: once per image (main.zig:86) ------------------------------------------ */

/* rgb: interleaved 8-bit RGB, `stride` bytes per row (>= 3*w).  Uploads the source and builds its six-scale XYB
 * pyramid on the context's SOURCE stream, next to whatever the compute streams are doing (the candidate's pyramid of
 * the same evaluation, or the previous image's submissions): the source side exists twice, so a new source never
 * waits for submissions that still read the old one.  For the RECURSIVE blur the rows pass of the two source-only
 * quantities (a, a*a) is computed once per source and cached (see OAVIF_SSIMU2_OPT_SOURCE_ROWS for when).
 * Returns once the caller's pixels have been read. */
int oavif_ssimu2_set_source_rgb8(oavif_ssimu2_ctx *ctx, const uint8_t *rgb, uint32_t w,
                                 uint32_t h, size_t stride);

/* The loaders' native layouts straight in: 1..4 interleaved channels of 8- or 16-bit samples (native
 * endian), reduced to RGB8 on the device exactly like Image.toRGB8 (src/io.zig:57-133: 16-bit >> 8,
 * alpha dropped, gray replicated).  channels outside 1..4 -> E_UNSUPPORTED (UnsupportedChannelCount). */
int oavif_ssimu2_set_source_pixels(oavif_ssimu2_ctx *ctx, const void *pixels, uint32_t w, uint32_t h,
                                   size_t stride, int channels, int bits);

/* ---- distorted side: once per search pass (tq.zig:150) ------------------------------------- */

int oavif_ssimu2_score_rgb8(oavif_ssimu2_ctx *ctx, const uint8_t *dist, size_t stride,
                            double *score);

/* A decoded frame in the layout avifImageYUVToRGB leaves behind (RGB or RGBA, io.zig:473): the
 * per-pixel repack of decodeAvifToRgb (io.zig:654-663) happens on the device. */
int oavif_ssimu2_score_pixels(oavif_ssimu2_ctx *ctx, const void *pixels, size_t stride, int channels,
                              int bits, double *score);

/* Decoded planes as libavif hands them over: depth 8 -> uint8_t samples, depth 10 -> uint16_t;
 * strides in BYTES; matrix = AV1 matrix_coefficients (1, 2, 5, 6, 9).
 * FULL RANGE ONLY, and there is no range argument: the caller must check avifImage.yuvRange ==
 * AVIF_RANGE_FULL (what avifImageCreate defaults to, io.zig:546) and otherwise take the score_pixels path
 * with libavif's own conversion.  The integer arithmetic is libavif's libyuv path (a libavif built without
 * libyuv converts in float and can differ by 1 LSB; score_pixels is the bit-faithful route there too).
 * rgba_path != 0 reproduces the conversion libavif runs when the decoded image has an alpha
 * plane (io.zig:473); alpha itself is never scored (io.zig:654-663). */
int oavif_ssimu2_score_yuv444(oavif_ssimu2_ctx *ctx, const void *y, const void *u, const void *v,
                              size_t y_stride, size_t u_stride, size_t v_stride, int depth,
                              int matrix, int rgba_path, double *score);

/* n candidates (n <= max_batch) against the cached source in one pass over the device. */
int oavif_ssimu2_score_batch_rgb8(oavif_ssimu2_ctx *ctx, uint32_t n, const uint8_t *const *dists,
                                  size_t stride, double *scores);
int oavif_ssimu2_score_batch_yuv444(oavif_ssimu2_ctx *ctx, uint32_t n, const void *const *y,
                                    const void *const *u, const void *const *v, size_t y_stride,
                                    size_t u_stride, size_t v_stride, int depth, int matrix,
                                    int rgba_path, double *scores);

/* ---- pipelined form ------------------------------------------------------------------------------
 * submit_* enqueues the upload (on the context's copy stream) and the kernels (on its compute stream) of
 * n candidates and returns at once; wait retires the OLDEST submission and stores its n scores.  Up to
 * two submissions may be in flight, and set_source_* may be called while one is: the upload of image
 * i+1 then runs under the kernels of image i, which is what one caller needs to keep the PCIe link busy,
 * and the two submissions' kernels run on two compute streams with candidate-side buffers of their own,
 * so the rows pass of one fills the tails of the other's columns pass
 * (the synchronous score_* calls are submit + wait).  Caller memory passed to submit_* must stay valid
 * and unmodified until the matching wait returns.  The second slot's buffers (staging, pyramid, row-filtered
 * planes: about as much again as the context itself) are allocated by the first submit that needs them.  get_detail / get_timing describe the
 * submission retired last. */
int oavif_ssimu2_submit_rgb8(oavif_ssimu2_ctx *ctx, uint32_t n, const uint8_t *const *dists,
                             size_t stride);
int oavif_ssimu2_submit_yuv444(oavif_ssimu2_ctx *ctx, uint32_t n, const void *const *y,
                               const void *const *u, const void *const *v, size_t y_stride,
                               size_t u_stride, size_t v_stride, int depth, int matrix, int rgba_path);
int oavif_ssimu2_submit_rgb8_dev(oavif_ssimu2_ctx *ctx, uint32_t n, const uint8_t *const *d_dists,
                                 size_t stride);
int oavif_ssimu2_submit_yuv444_dev(oavif_ssimu2_ctx *ctx, uint32_t n, const void *const *d_y,
                                   const void *const *d_u, const void *const *d_v, size_t y_stride,
                                   size_t u_stride, size_t v_stride, int depth, int matrix, int rgba_path);
int oavif_ssimu2_wait(oavif_ssimu2_ctx *ctx, double *scores);
int oavif_ssimu2_in_flight(const oavif_ssimu2_ctx *ctx);

/* ---- device-resident inputs (pointers are CUDA device pointers on the context's device) --
 * set_source_rgb8_dev only enqueues work: the buffer must stay unmodified until the next score call
 * on this context has returned.  Ordering: with a caller-owned stream (set_stream) the source is read
 * after everything enqueued on that stream so far; with the context's own streams the pixels must already
 * be complete when the call is made. */

int oavif_ssimu2_set_source_rgb8_dev(oavif_ssimu2_ctx *ctx, const uint8_t *d_rgb, uint32_t w,
                                     uint32_t h, size_t stride);
int oavif_ssimu2_score_batch_rgb8_dev(oavif_ssimu2_ctx *ctx, uint32_t n,
                                      const uint8_t *const *d_dists, size_t stride, double *scores);
int oavif_ssimu2_score_batch_yuv444_dev(oavif_ssimu2_ctx *ctx, uint32_t n, const void *const *d_y,
                                        const void *const *d_u, const void *const *d_v,
                                        size_t y_stride, size_t u_stride, size_t v_stride,
                                        int depth, int matrix, int rgba_path, double *scores);

/* ---- stateless forms ------------------------------------------------------------------------ */

/* fssimu2.computeSsimu2(ref, dist, w, h, channels) — tq.zig:37.  channels must be 3 (the only
 * value oavif passes); tight rows.  Runs on ONE process-wide context that is created on first use on the
 * default device (0 unless oavif_ssimu2_set_default_device was called), regrown when an image does not fit,
 * serialised by a mutex, and kept alive (about 1 GB of HBM at 4K) until oavif_ssimu2_release_cached(). */
int oavif_ssimu2_compute_rgb8(const uint8_t *ref, const uint8_t *dist, uint32_t w, uint32_t h,
                              uint32_t channels, double *score);
int oavif_ssimu2_set_default_device(int device);
void oavif_ssimu2_release_cached(void);

/* The encoder-side depth conversions of encodeAvifToBuffer (src/io.zig:562-609), on the device, from the pixels the
 * last oavif_ssimu2_set_source_* call staged: the array avifImageRGBToYUV is handed,
 *     8-bit source,  out_depth 10:  (v * 1023 + 127) / 255      io.zig:566-572
 *     16-bit source, out_depth 10:  v >> 6                        io.zig:581-587
 *     16-bit source, out_depth 8:   v >> 8                        io.zig:596-602
 * with every channel kept (alpha rides along to the encoder, io.zig:564) and tight rows: `out` receives
 * w * h * channels samples, uint16_t for depth 10 and uint8_t for depth 8.  The reference runs these scalar loops in
 * EVERY search pass; a caller runs this once per image and hands the same array to every encode.  An 8-bit source
 * at depth 8 needs no conversion (io.zig:611-613) -> E_UNSUPPORTED.  Call it between set_source_* and the next
 * set_source_* (device sources: while the caller's buffer is alive); images below 8x8 are not staged -> E_STATE. */
int oavif_ssimu2_source_samples(oavif_ssimu2_ctx *ctx, int out_depth, void *out, size_t out_bytes);

/* decodeAvifToRgb's pixel work (io.zig:470-478, 654-663) alone: planes -> tight RGB8.  Uses the candidate
 * staging buffer and an output buffer of its own; the cached source is untouched, so it may be called in the
 * middle of a search. */
int oavif_ssimu2_yuv444_to_rgb8(oavif_ssimu2_ctx *ctx, const void *y, const void *u, const void *v,
                                size_t y_stride, size_t u_stride, size_t v_stride, uint32_t w,
                                uint32_t h, int depth, int matrix, int rgba_path, uint8_t *rgb_out);

/* ---- introspection ----------------------------------------------------------------------------- */

int oavif_ssimu2_get_detail(oavif_ssimu2_ctx *ctx, uint32_t candidate, oavif_ssimu2_detail *out);
int oavif_ssimu2_get_timing(oavif_ssimu2_ctx *ctx, oavif_ssimu2_timing *out);

/* Copy one XYB plane of the cached pyramids to the host (tight w_s*h_s floats).
 * which: 0 = source, 1 + k = candidate k of the last score call.  channel: 0 X, 1 Y, 2 B. */
int oavif_ssimu2_debug_get_xyb(oavif_ssimu2_ctx *ctx, int which, int scale, int channel,
                               float *out, uint32_t *w_out, uint32_t *h_out);

/* The RECURSIVE rows pass as the scored path left it after the last score call: quantity 0..4 = a, b,
 * a*a, b*b, a*b (row-filtered, before the columns pass), for one candidate, scale and channel; w*h floats.
 * Lets a test compare the product kernel's recursion bit for bit with the CPU oracle's horizontal pass. */
int oavif_ssimu2_debug_get_rows(oavif_ssimu2_ctx *ctx, int candidate, int quantity, int scale, int channel,
                                float *out, uint32_t *w_out, uint32_t *h_out);

/* The RECURSIVE columns pass of the scored path: re-runs the product kernel (k_iir_cols) over the row-filtered
 * planes the last score call left and copies out the five fully blurred values it hands to the error maps —
 * mu1, mu2, sigma11, sigma22, sigma12, in that order, each w*h floats — for one candidate, scale and channel.
 * out: 5*w*h floats.  Blurred planes never exist in HBM on the scored path; this tap is how a test compares
 * them bit for bit with the CPU oracle's full blur. */
int oavif_ssimu2_debug_get_cols(oavif_ssimu2_ctx *ctx, int candidate, int scale, int channel, float *out,
                                uint32_t *w_out, uint32_t *h_out);

/* Timeline of one launch of the FUSED kernel over the last scored pair (mode 2: all five quantities, 1: cached source
 * blur): per CTA, in ticket order, five 64-bit words {unit = scale | channel << 4 | strip << 8, start, end of the
 * prologue, middle phase, end} from %globaltimer (ns).  Shows how the wavefront of strips actually advances. */
int oavif_ssimu2_debug_wave_trace(oavif_ssimu2_ctx *ctx, int mode, uint64_t *out, uint32_t cap_units, uint32_t *n_units);

/* Blur one host plane with the selected blur on the device (tests of the filter alone). */
int oavif_ssimu2_debug_blur(oavif_ssimu2_ctx *ctx, const float *in, uint32_t w, uint32_t h,
                            float *out);

/* Every large device buffer of a context is followed by a 4 KB guard band; returns 0 if all bands are
 * intact, E_STATE (with the band index in last_error) if a kernel wrote past a buffer.  Stands in for
 * compute-sanitizer, which the target pool does not allow. */
int oavif_ssimu2_debug_check_guards(oavif_ssimu2_ctx *ctx);

/* Profiling aid: re-run only the RECURSIVE rows pass on the pyramids of the last score call,
 * `iters` times, and report its mean device time.  variant 0 runs both halves; bit 2 (value 4) leaves
 * out the source half (a, a*a), i.e. times what a call with a warm source cache runs; bit 3 (value 8) forces
 * the cp.async kernels whatever OAVIF_SSIMU2_OPT_TILE_PATH says; bits 4..6 select a TMA instance (ring depth /
 * staging buffers, 0 = the shipped one); 128 = the source half alone; 256 = both halves issued the way the scored
 * path issues them (source stream next to compute stream; the time is that of the pair); 512 = the COLUMNS pass alone
 * (| 8: its cp.async loader, | 4096: its instance without the per-batch block barrier, | 8192: 40 KB of unused
 * shared memory on top, i.e. two CTAs per SM instead of three, | 16384: descriptors that read strip-major — durations
 * only, the data is wrong —, | 32768: 256-byte L2 promotion on the pair planes' descriptors);
 * 1024 / 2048 = the FUSED kernel with all five quantities / with the cached source blur. */
int oavif_ssimu2_debug_time_rows(oavif_ssimu2_ctx *ctx, int variant, int iters, float *mean_ms);

#ifdef __cplusplus
}
#endif
#endif
