// host_capi.cpp — extern "C" face of the harness (include/oavif_host.h).
#include <cstring>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../../include/oavif_host.h"
#include "oavif_host.hpp"

using namespace oavif_host;

namespace {
thread_local std::string t_err;

EncOptions to_cpp(const oavif_host_opts *o)
{
    EncOptions e;
    if (!o) return e;
    e.quality_alpha = o->quality_alpha;
    e.speed = o->speed;
    e.max_threads = o->max_threads;
    e.tile_rows_log2 = o->tile_rows_log2;
    e.tile_cols_log2 = o->tile_cols_log2;
    e.auto_tiling = o->auto_tiling != 0;
    e.score_tgt = o->score_tgt;
    e.tenbit = o->tenbit != 0;
    e.tune = o->tune;
    e.tolerance = o->tolerance;
    e.max_pass = o->max_pass;
    e.quality = o->quality;
    e.color_primaries = o->color_primaries;
    e.transfer_characteristics = o->transfer_characteristics;
    e.matrix_coefficients = o->matrix_coefficients;
    return e;
}

void fill(oavif_host_result *out, const TQResult &r, const BatchedStats *b)
{
    out->q = r.q;
    out->score = r.score;
    out->num_pass = r.num_pass;
    out->early_exit = r.early_exit;
    out->n_history = (uint32_t)std::min<size_t>(16, r.history.size());
    for (uint32_t i = 0; i < out->n_history; ++i) {
        out->hist_q[i] = r.history[i].q;
        out->hist_score[i] = r.history[i].score;
    }
    if (b) {
        out->device_passes = b->device_passes;
        out->probes = b->probes;
        out->wasted = b->wasted;
    }
}

struct CallbackScorer : ScorerIface {
    oavif_host_set_source_fn ss;
    oavif_host_score_fn sc;
    void *user;
    uint32_t w = 0, h = 0;
    void set_source(const uint8_t *rgb, uint32_t w_, uint32_t h_) override
    {
        w = w_;
        h = h_;
        if (ss(user, rgb, w_, h_) != 0) throw std::runtime_error("injected set_source failed");
    }
    std::vector<double> score(const std::vector<const Decoded *> &c) override
    {
        std::vector<double> out(c.size());
        for (size_t i = 0; i < c.size(); ++i) {
            const AvifImageView &im = c[i]->img;
            if (sc(user, im.plane(0), im.plane(1), im.plane(2), im.rowBytes(0), im.rowBytes(1), im.rowBytes(2), w, h,
                   (int)im.depth(), (int)im.matrixCoefficients(), im.alphaPlane() != nullptr, &out[i]) != 0)
                throw std::runtime_error("injected score failed");
        }
        return out;
    }
};
// the corpus driver's CPU arm: keeps the source pointer (alive for the whole search) and calls a stateless scorer
struct PairCallbackScorer : ScorerIface {
    oavif_host_score_pair_fn fn;
    void *user;
    const uint8_t *src = nullptr;
    uint32_t w = 0, h = 0;
    PairCallbackScorer(oavif_host_score_pair_fn f, void *u) : fn(f), user(u) {}
    void set_source(const uint8_t *rgb, uint32_t w_, uint32_t h_) override
    {
        src = rgb;
        w = w_;
        h = h_;
    }
    std::vector<double> score(const std::vector<const Decoded *> &c) override
    {
        std::vector<double> out(c.size());
        for (size_t i = 0; i < c.size(); ++i) {
            const AvifImageView &im = c[i]->img;
            if (fn(user, src, im.plane(0), im.plane(1), im.plane(2), im.rowBytes(0), im.rowBytes(1), im.rowBytes(2), w, h,
                   (int)im.depth(), (int)im.matrixCoefficients(), im.alphaPlane() != nullptr, &out[i]) != 0)
                throw std::runtime_error("injected score failed");
        }
        return out;
    }
};
}  // namespace

extern "C" {

void oavif_host_default_opts(oavif_host_opts *o)
{
    if (!o) return;
    memset(o, 0, sizeof *o);
    const EncOptions e;
    o->quality_alpha = e.quality_alpha;
    o->speed = e.speed;
    o->max_threads = e.max_threads;
    o->auto_tiling = e.auto_tiling;
    o->score_tgt = e.score_tgt;
    o->tenbit = e.tenbit;
    strncpy(o->tune, e.tune.c_str(), sizeof o->tune - 1);
    o->tolerance = e.tolerance;
    o->max_pass = e.max_pass;
    o->quality = e.quality;
    o->color_primaries = e.color_primaries;
    o->transfer_characteristics = e.transfer_characteristics;
    o->matrix_coefficients = e.matrix_coefficients;
}

const char *oavif_host_last_error(void) { return t_err.c_str(); }

uint32_t oavif_host_predict_q(double t) { return predictQFromScore(t); }

uint32_t oavif_host_interpolate_q(uint32_t lo, uint32_t hi, const uint32_t *qs, const double *scores, uint32_t n,
                                  double target)
{
    std::vector<PassResult> h;
    for (uint32_t i = 0; i < n; ++i) h.push_back({qs[i], scores[i]});
    return interpolateQuantizer(lo, hi, h, target);
}

int oavif_host_tq_search(double tgt, double tol, uint32_t max_pass, oavif_host_probe_fn probe, void *user,
                         oavif_host_result *out)
{
    if (!probe || !out) return -1;
    memset(out, 0, sizeof *out);
    TQOptions o{tgt, tol, max_pass};
    fill(out, findTargetQuality(o, [&](uint32_t q) { return probe(user, q); }), nullptr);
    return 0;
}

int oavif_host_tq_search_batched(double tgt, double tol, uint32_t max_pass, uint32_t width,
                                 oavif_host_probe_batch_fn probe, void *user, oavif_host_result *out)
{
    if (!probe || !out) return -1;
    memset(out, 0, sizeof *out);
    TQOptions o{tgt, tol, max_pass};
    BatchedStats st;
    TQResult r = findTargetQualityBatched(o, width, [&](const std::vector<uint32_t> &qs) {
        std::vector<double> sc(qs.size());
        probe(user, (uint32_t)qs.size(), qs.data(), sc.data());
        return sc;
    }, &st);
    fill(out, r, &st);
    return 0;
}

int oavif_host_search_image(const char *libavif_path, const uint8_t *pixels, uint32_t w, uint32_t h,
                            uint32_t channels, const oavif_host_opts *opts, uint32_t batch_width, int device,
                            int blur_mode, oavif_host_set_source_fn set_source, oavif_host_score_fn score,
                            void *user, oavif_host_result *out, uint8_t *avif_out, size_t avif_cap)
{
    if (!libavif_path || !pixels || !out || (channels != 3 && channels != 4) || !w || !h) {
        t_err = "bad argument";
        return -1;
    }
    try {
        memset(out, 0, sizeof *out);
        LibAvif lib(libavif_path);
        Codec codec(lib);
        HostImage img;
        img.w = w;
        img.h = h;
        img.channels = channels;
        img.data.assign(pixels, pixels + (size_t)w * h * channels);
        const EncOptions o = to_cpp(opts);
        SearchResult sr;
        if (device >= 0) {
            GpuScorer gs(device, w, h, std::max(1u, batch_width), blur_mode);
            sr = search_image(codec, gs, img, o, batch_width, batch_width);
        } else {
            if (!set_source || !score) {
                t_err = "no scorer";
                return -1;
            }
            CallbackScorer cs;
            cs.ss = set_source;
            cs.sc = score;
            cs.user = user;
            sr = search_image(codec, cs, img, o, batch_width, batch_width);
        }
        fill(out, sr.tq, &sr.batched);
        out->size = sr.size;
        out->reencoded = sr.reencoded;
        out->encode_ms = sr.encode_ms;
        out->decode_ms = sr.decode_ms;
        out->score_ms = sr.score_ms;
        out->total_ms = sr.total_ms;
        strncpy(out->log, sr.log.c_str(), sizeof out->log - 1);
        if (avif_out) {
            if (sr.avif.size() > avif_cap) {
                t_err = "avif_out too small";
                return -2;
            }
            memcpy(avif_out, sr.avif.data(), sr.avif.size());
        }
        return 0;
    } catch (const std::exception &e) {
        t_err = e.what();
        return -3;
    }
}

int oavif_host_tq_margins(double tgt, double tol, uint32_t max_pass, const uint32_t *qs, const double *scores, uint32_t n,
                          double limit, double *flip_up, double *flip_down)
{
    TQOptions o;
    o.score_tgt = tgt;
    o.tolerance = tol;
    o.max_pass = max_pass;
    std::vector<PassResult> h;
    for (uint32_t i = 0; i < n; ++i) h.push_back({qs[i], scores[i]});
    const auto ms = decisionMargins(o, h, limit);
    for (uint32_t i = 0; i < n; ++i) {
        flip_up[i] = ms[i].flip_up;
        flip_down[i] = ms[i].flip_down;
    }
    return 0;
}

int oavif_host_corpus_synth(const char *libavif_path, uint32_t count, uint32_t w, uint32_t h, int first_gpu,
                            int n_gpus, uint32_t workers_per_gpu, uint32_t batch_width, int blur_mode,
                            int pinned_staging, oavif_host_score_pair_fn score_pair, void *user,
                            const oavif_host_opts *opts, const char *csv_path, char *summary, size_t summary_cap,
                            oavif_host_corpus_stats *stats)
{
    try {
        CorpusSpec spec;
        spec.synth_count = count;
        spec.synth_w = w;
        spec.synth_h = h;
        spec.first_gpu = first_gpu;
        spec.n_gpus = n_gpus;
        spec.workers_per_gpu = workers_per_gpu;
        spec.batch_width = batch_width;
        spec.blur_mode = blur_mode;
        spec.pinned_staging = pinned_staging != 0;
        if (score_pair)
            spec.scorer_factory = [score_pair, user](int) {
                return std::unique_ptr<ScorerIface>(new PairCallbackScorer(score_pair, user));
            };
        CorpusStats st;
        const auto rows = run_corpus(libavif_path, spec, to_cpp(opts), &st);
        if (csv_path) {
            const std::string files[2][2] = {{csv_path, corpus_csv(rows)},
                                             {std::string(csv_path) + ".trace.csv", corpus_trace_csv(rows)}};
            for (const auto &fc : files) {
                FILE *f = fopen(fc[0].c_str(), "wb");
                if (!f) throw std::runtime_error("cannot write " + fc[0]);
                fwrite(fc[1].data(), 1, fc[1].size(), f);
                fclose(f);
            }
        }
        if (summary && summary_cap) {
            const std::string s = corpus_summary(rows, st);
            strncpy(summary, s.c_str(), summary_cap - 1);
            summary[summary_cap - 1] = 0;
        }
        if (stats) {
            memset(stats, 0, sizeof *stats);
            stats->wall_s = st.wall_s;
            stats->scorer_device_ms = st.scorer_device_ms;
            stats->workers = st.workers;
            stats->host_cpus = st.host_cpus;
            const double edges[] = {1e-4, 1e-3, 1e-2, 0.05, 0.1, 0.5, 1e30};
            for (const auto &r : rows) {
                if (r.status != "ok") {
                    ++stats->n_err;
                    continue;
                }
                ++stats->n_ok;
                stats->mean_encode_ms += r.encode_ms;
                stats->mean_decode_ms += r.decode_ms;
                stats->mean_score_ms += r.score_ms;
                stats->mean_passes += r.passes;
                stats->final_bytes_total += r.final_bytes;
                for (int b = 0; b < 7; ++b)
                    if (r.margin <= edges[b]) {
                        ++stats->margin_hist[b];
                        break;
                    }
            }
            if (stats->n_ok) {
                stats->mean_encode_ms /= stats->n_ok;
                stats->mean_decode_ms /= stats->n_ok;
                stats->mean_score_ms /= stats->n_ok;
                stats->mean_passes /= stats->n_ok;
            }
        }
        for (const auto &r : rows)
            if (r.status != "ok") t_err = r.error;
        return 0;
    } catch (const std::exception &e) {
        t_err = e.what();
        return -3;
    }
}

int oavif_host_encode(const char *libavif_path, const uint8_t *pixels, uint32_t w, uint32_t h, uint32_t channels,
                      uint32_t q, const oavif_host_opts *opts, uint8_t *out, size_t cap, size_t *size)
{
    try {
        LibAvif lib(libavif_path);
        Codec codec(lib);
        HostImage img;
        img.w = w;
        img.h = h;
        img.channels = channels;
        img.data.assign(pixels, pixels + (size_t)w * h * channels);
        const auto bytes = codec.encode(img, q, to_cpp(opts));
        if (size) *size = bytes.size();
        if (out) {
            if (bytes.size() > cap) {
                t_err = "output buffer too small";
                return -2;
            }
            memcpy(out, bytes.data(), bytes.size());
        }
        return 0;
    } catch (const std::exception &e) {
        t_err = e.what();
        return -3;
    }
}

int oavif_host_decode_rgb8(const char *libavif_path, const uint8_t *avif, size_t size, uint8_t *rgb_out, size_t cap)
{
    try {
        LibAvif lib(libavif_path);
        Codec codec(lib);
        const auto rgb = codec.decode_to_rgb8(std::vector<uint8_t>(avif, avif + size));
        if (rgb.size() > cap) {
            t_err = "output buffer too small";
            return -2;
        }
        memcpy(rgb_out, rgb.data(), rgb.size());
        return 0;
    } catch (const std::exception &e) {
        t_err = e.what();
        return -3;
    }
}

}  // extern "C"
