// oavif-b200 — command-line face of the host harness: the reference's CLI surface for the parts that
// touch the scored path (flags of src/parse_args.zig:76-122, stderr lines of src/main.zig:78-116) plus
// the two additive modes north_star asks for: --batch K (batched probing) and --corpus (the
// scripts/measure.py sweep, images sharded over --gpus G GPUs, no NCCL).
//
//   oavif-b200 [options] in.{ppm,pam} out.avif
//   oavif-b200 [options] --corpus synth:N:WxH out.csv
// Inputs: PPM/PAM only (the other loaders of src/io.zig are out of scope); set OAVIF_LIBAVIF to the
// libavif shared object.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../../include/oavif_host.h"
#include "oavif_host.hpp"

using namespace oavif_host;

static bool arg_val(int &i, int argc, char **argv, std::string &out, const char *name)
{  // a value starting with '-' counts as missing (parse_args.zig:126)
    if (i + 1 >= argc || argv[i + 1][0] == '-') {
        fprintf(stderr, "Error: Missing value for %s\n", name);
        return false;
    }
    out = argv[++i];
    return true;
}

static bool int_arg(int &i, int argc, char **argv, long lo, long hi, const char *name, long &v)
{
    std::string s;
    if (!arg_val(i, argc, argv, s, name)) return false;
    char *end = nullptr;
    v = strtol(s.c_str(), &end, 10);
    if (*end || v < lo || v > hi) {
        fprintf(stderr, "Error: %s must be between %ld and %ld\n", name, lo, hi);
        return false;
    }
    return true;
}

static bool float_arg(int &i, int argc, char **argv, double lo, double hi, const char *name, double &v)
{
    std::string s;
    if (!arg_val(i, argc, argv, s, name)) return false;
    char *end = nullptr;
    v = strtod(s.c_str(), &end);
    if (*end || v < lo || v > hi) {
        fprintf(stderr, "Error: %s must be between %g and %g\n", name, lo, hi);
        return false;
    }
    return true;
}

int main(int argc, char **argv)
{
    fprintf(stderr, "\x1b[31moavif\x1b[0m | b200 host harness\n");
    EncOptions o;
    std::string in, out, corpus;
    long v = 0, batch = 1, gpus = 1, workers = 1, device = 0, blur = 0, pinned = 1;
    double d = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        std::string sv;
        if (a == "-s" || a == "--speed") { if (!int_arg(i, argc, argv, 0, 10, "--speed", v)) return 2; o.speed = (uint32_t)v; }
        else if (a == "-t" || a == "--score-tgt") { if (!float_arg(i, argc, argv, 30, 100, "--score-tgt", d)) return 2; o.score_tgt = d; }
        // parse_args.zig:88 accepts 0..99 while its help text says 0..100; widened (pure relaxation)
        else if (a == "--quality-alpha") { if (!int_arg(i, argc, argv, 0, 100, "--quality-alpha", v)) return 2; o.quality_alpha = (uint32_t)v; }
        else if (a == "--max-threads") { if (!int_arg(i, argc, argv, 1, 255, "--max-threads", v)) return 2; o.max_threads = (uint32_t)v; }
        else if (a == "--tile-rows-log2") { if (!int_arg(i, argc, argv, 0, 6, "--tile-rows-log2", v)) return 2; o.tile_rows_log2 = (uint32_t)v; }
        else if (a == "--tile-cols-log2") { if (!int_arg(i, argc, argv, 0, 6, "--tile-cols-log2", v)) return 2; o.tile_cols_log2 = (uint32_t)v; }
        else if (a == "--auto-tiling") { if (!int_arg(i, argc, argv, 0, 1, "--auto-tiling", v)) return 2; o.auto_tiling = v != 0; }
        else if (a == "--tune") {
            if (!arg_val(i, argc, argv, sv, "--tune")) return 2;
            if (sv != "ssim" && sv != "iq" && sv != "ssimulacra2") { fprintf(stderr, "Error: InvalidTuneMode\n"); return 2; }
            o.tune = sv;
        }
        else if (a == "--tenbit") { if (!int_arg(i, argc, argv, 0, 1, "--tenbit", v)) return 2; o.tenbit = v != 0; }
        else if (a == "--tolerance") { if (!float_arg(i, argc, argv, 1, 100, "--tolerance", d)) return 2; o.tolerance = d; }
        else if (a == "--max-pass") { if (!int_arg(i, argc, argv, 1, 12, "--max-pass", v)) return 2; o.max_pass = (uint32_t)v; }
        else if (a == "-q" || a == "--quality") { if (!int_arg(i, argc, argv, 0, 100, "--quality", v)) return 2; o.quality = (int)v; }
        else if (a == "--color-primaries") { if (!int_arg(i, argc, argv, 1, 22, "--color-primaries", v)) return 2; o.color_primaries = (uint32_t)v; }
        else if (a == "--transfer-characteristics") { if (!int_arg(i, argc, argv, 1, 18, "--transfer-characteristics", v)) return 2; o.transfer_characteristics = (uint32_t)v; }
        else if (a == "--matrix-coefficients") { if (!int_arg(i, argc, argv, 0, 14, "--matrix-coefficients", v)) return 2; o.matrix_coefficients = (uint32_t)v; }
        // ---- additive, default-off ----
        else if (a == "--batch") { if (!int_arg(i, argc, argv, 1, 16, "--batch", batch)) return 2; }
        else if (a == "--gpus") { if (!int_arg(i, argc, argv, 1, 8, "--gpus", gpus)) return 2; }
        else if (a == "--workers-per-gpu") { if (!int_arg(i, argc, argv, 1, 64, "--workers-per-gpu", workers)) return 2; }
        else if (a == "--device") { if (!int_arg(i, argc, argv, 0, 7, "--device", device)) return 2; }
        else if (a == "--blur") { if (!int_arg(i, argc, argv, 0, 1, "--blur", blur)) return 2; }
        else if (a == "--pinned-staging") { if (!int_arg(i, argc, argv, 0, 1, "--pinned-staging", pinned)) return 2; }
        else if (a == "--corpus") { if (!arg_val(i, argc, argv, corpus, "--corpus")) return 2; }
        else if (in.empty()) in = a;
        else if (out.empty()) out = a;
        else { fprintf(stderr, "Error: Unexpected argument: %s\n", a.c_str()); return 2; }
    }
    const char *lib = getenv("OAVIF_LIBAVIF");
    if (!lib) { fprintf(stderr, "Error: set OAVIF_LIBAVIF to the libavif shared object\n"); return 2; }
    try {
        if (!corpus.empty()) {
            CorpusSpec spec;
            unsigned n = 0, w = 0, h = 0;
            if (sscanf(corpus.c_str(), "synth:%u:%ux%u", &n, &w, &h) != 3) { fprintf(stderr, "Error: --corpus synth:N:WxH\n"); return 2; }
            spec.synth_count = n; spec.synth_w = w; spec.synth_h = h;
            spec.first_gpu = (int)device; spec.n_gpus = (int)gpus; spec.workers_per_gpu = (uint32_t)workers;
            spec.batch_width = (uint32_t)batch; spec.blur_mode = (int)blur;
            spec.pinned_staging = pinned != 0;
            CorpusStats st;
            const auto rows = run_corpus(lib, spec, o, &st);
            const std::string csv_path = in.empty() ? "corpus.csv" : in;
            const std::string files[2][2] = {{csv_path, corpus_csv(rows)}, {csv_path + ".trace.csv", corpus_trace_csv(rows)}};
            for (const auto &fc : files) {
                FILE *f = fopen(fc[0].c_str(), "wb");
                if (!f) throw std::runtime_error("cannot write " + fc[0]);
                fwrite(fc[1].data(), 1, fc[1].size(), f);
                fclose(f);
            }
            fprintf(stderr, "%s\nResults written to %s\n", corpus_summary(rows, st).c_str(), csv_path.c_str());
            return 0;
        }
        if (in.empty() || out.empty()) { fprintf(stderr, "error: MissingInputOrOutput\n"); return 2; }
        LibAvif L(lib);
        Codec codec(L);
        HostImage img = load_pnm(in);
        fprintf(stderr, "Read %ux%u, %s, 8-bit, %zu bytes\n", img.w, img.h, img.channels > 3 ? "RGBA" : "RGB", img.file_bytes);
        SearchResult r;
        if (o.quality >= 0) {  // -q bypass never touches the scorer (main.zig:93-100)
            struct NoScorer : ScorerIface {
                void set_source(const uint8_t *, uint32_t, uint32_t) override {}
                std::vector<double> score(const std::vector<const Decoded *> &) override { return {}; }
            } none;
            r = search_image(codec, none, img, o, 1, 1);
        } else {
            GpuScorer scorer((int)device, img.w, img.h, (uint32_t)batch, (int)blur);
            r = search_image(codec, scorer, img, o, (uint32_t)batch, (uint32_t)batch);
        }
        fputs(r.log.c_str(), stderr);
        if (getenv("OAVIF_MARGINS"))   // trace tooling: how far each pass's score was from changing the search
            for (const auto &m : r.margins)
                fprintf(stderr, "pass %u: q%u score %.6f  flips at +%.6f / -%.6f\n", m.pass, m.q, m.score, m.flip_up, m.flip_down);
        {
            FILE *f = fopen(out.c_str(), "wb");
            if (!f) throw std::runtime_error("cannot write " + out);
            fwrite(r.avif.data(), 1, r.avif.size(), f);
            fclose(f);
        }
        return 0;
    } catch (const std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
