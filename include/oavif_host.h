/*
 * include/oavif_host.h — C view of the C++ host harness (oavif_b200/host/cpp), for tests and tools.
 * This is NOT the drop-in boundary (that is oavif_ssimu2.h); it exposes the restated callers of the
 * scored path — the search policy of /root/reference/src/tq.zig, the encode/decode glue of
 * src/io.zig and the corpus loop of scripts/measure.py — so that they can be driven from ctypes.
 */
#ifndef OAVIF_HOST_H
#define OAVIF_HOST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {              /* AvifEncOptions, src/parse_args.zig:48-63 (same defaults via oavif_host_default_opts) */
    uint32_t quality_alpha, speed, max_threads, tile_rows_log2, tile_cols_log2, auto_tiling;
    double score_tgt;
    uint32_t tenbit;
    char tune[16];
    double tolerance;
    uint32_t max_pass;
    int32_t quality;          /* -1: target-quality search; >= 0: -q bypass (main.zig:93-100) */
    uint32_t color_primaries, transfer_characteristics, matrix_coefficients;
} oavif_host_opts;

typedef struct {
    uint32_t q;               /* e.q  */
    double score;             /* e.t.score */
    uint32_t num_pass;        /* e.t.num_pass */
    uint32_t early_exit;
    uint32_t n_history;
    uint32_t hist_q[16];
    double hist_score[16];
    uint32_t device_passes, probes, wasted;   /* batched mode accounting */
    uint64_t size;            /* bytes oavif would write */
    uint32_t reencoded;       /* main.zig:113 */
    double encode_ms, decode_ms, score_ms, total_ms;
    char log[512];            /* stderr lines of main.zig:102-116 */
} oavif_host_result;

typedef double (*oavif_host_probe_fn)(void *user, uint32_t q);
typedef void (*oavif_host_probe_batch_fn)(void *user, uint32_t n, const uint32_t *qs, double *scores);

/* Scorer injected by tests (e.g. the CPU oracle) instead of the CUDA library. */
typedef int (*oavif_host_set_source_fn)(void *user, const uint8_t *rgb, uint32_t w, uint32_t h);
typedef int (*oavif_host_score_fn)(void *user, const void *y, const void *u, const void *v, uint32_t y_stride,
                                   uint32_t u_stride, uint32_t v_stride, uint32_t w, uint32_t h, int depth,
                                   int matrix, int rgba_path, double *score);

/* Stateless form for the corpus driver's CPU-scored arm (many workers call it concurrently): source RGB8 and one
 * decoded candidate in, one score out. */
typedef int (*oavif_host_score_pair_fn)(void *user, const uint8_t *src_rgb, const void *y, const void *u, const void *v,
                                        uint32_t y_stride, uint32_t u_stride, uint32_t v_stride, uint32_t w, uint32_t h,
                                        int depth, int matrix, int rgba_path, double *score);

void oavif_host_default_opts(oavif_host_opts *o);
const char *oavif_host_last_error(void);

/* tq.zig:40-43, 73-122 */
uint32_t oavif_host_predict_q(double score_tgt);
uint32_t oavif_host_interpolate_q(uint32_t lo, uint32_t hi, const uint32_t *qs, const double *scores, uint32_t n,
                                  double target);
/* tq.zig:124-210 with the probe supplied by the caller */
int oavif_host_tq_search(double score_tgt, double tolerance, uint32_t max_pass, oavif_host_probe_fn probe,
                         void *user, oavif_host_result *out);
int oavif_host_tq_search_batched(double score_tgt, double tolerance, uint32_t max_pass, uint32_t width,
                                 oavif_host_probe_batch_fn probe, void *user, oavif_host_result *out);

/* main.zig:86-116 for one 8-bit image (channels 3 or 4).  device >= 0: CUDA scorer on that device;
 * device < 0: the injected scorer.  avif_out may be NULL. */
int oavif_host_search_image(const char *libavif_path, const uint8_t *pixels, uint32_t w, uint32_t h,
                            uint32_t channels, const oavif_host_opts *opts, uint32_t batch_width, int device,
                            int blur_mode, oavif_host_set_source_fn set_source, oavif_host_score_fn score,
                            void *user, oavif_host_result *out, uint8_t *avif_out, size_t avif_cap);

/* tq.hpp decisionMargins: for each pass of a finished search, the smallest increase / decrease of that pass's score
 * that would have changed what the search did (next quantizer, stop, final choice); `limit` = none within it. */
int oavif_host_tq_margins(double score_tgt, double tolerance, uint32_t max_pass, const uint32_t *qs, const double *scores,
                          uint32_t n, double limit, double *flip_up, double *flip_down);

typedef struct {
    double wall_s, scorer_device_ms;
    uint32_t n_ok, n_err, workers, host_cpus;
    double mean_encode_ms, mean_decode_ms, mean_score_ms, mean_passes;
    uint64_t final_bytes_total;
    uint32_t margin_hist[7];        /* <=1e-4, <=1e-3, <=0.01, <=0.05, <=0.1, <=0.5, >0.5 */
} oavif_host_corpus_stats;

/* scripts/measure.py over a procedural corpus (seed = index, kind = seed mod 4).  n_gpus x workers_per_gpu host
 * threads pull images from one shared counter; GPU arm: CUDA scorer on devices first_gpu.. (pinned_staging: decode
 * hand-off through pinned memory); CPU arm: pass score_pair (then no GPU is touched).  Writes the measure.py CSV
 * (csv_path) and the per-image trace (csv_path + ".trace.csv"), returns the summary text. */
int oavif_host_corpus_synth(const char *libavif_path, uint32_t count, uint32_t w, uint32_t h, int first_gpu,
                            int n_gpus, uint32_t workers_per_gpu, uint32_t batch_width, int blur_mode,
                            int pinned_staging, oavif_host_score_pair_fn score_pair, void *user,
                            const oavif_host_opts *opts, const char *csv_path, char *summary, size_t summary_cap,
                            oavif_host_corpus_stats *stats);

/* Encode / decode helpers for fixtures: io.zig:544-636 and io.zig:638-666. */
int oavif_host_encode(const char *libavif_path, const uint8_t *pixels, uint32_t w, uint32_t h, uint32_t channels,
                      uint32_t q, const oavif_host_opts *opts, uint8_t *out, size_t cap, size_t *size);
int oavif_host_decode_rgb8(const char *libavif_path, const uint8_t *avif, size_t size, uint8_t *rgb_out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
