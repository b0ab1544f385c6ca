// oavif_host.hpp — C++ host harness above the C ABI of include/oavif_ssimu2.h.
//
// The reference's host side is Zig and cannot be compiled in this image, so the parts of it that sit
// on either side of the scored path are restated here in C++, function for function:
//   AvifEncOptions / copyToEncoder      src/parse_args.zig:48-74      -> EncOptions, Codec::encode
//   encodeAvifToBuffer                  src/io.zig:544-636            -> Codec::encode
//   decodeAvifCommon (up to NextImage)  src/io.zig:452-466            -> Codec::decode (planes handed over,
//                                                                         no avifImageYUVToRGB, no repack)
//   computeScoreAtQuality               src/tq.zig:21-38              -> ImageJob::probe / probe_batch
//   findTargetQuality                   src/tq.zig:124-210            -> tq.hpp
//   main(): "Found q.." / re-encode     src/main.zig:103-116          -> search_image
//   scripts/measure.py                  CSV schema lines 178-206      -> run_corpus
// The loaders (PNG/JPEG/WebP, src/io.zig:136-445) are out of scope: the harness reads PAM/PPM (the
// one raw format the reference also reads, io.zig:309-406) or generates procedural images.
#pragma once

#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "avif_dl.hpp"
#include "tq.hpp"

struct oavif_ssimu2_ctx;

namespace oavif_host {

struct EncOptions {  // parse_args.zig:48-63, same defaults
    uint32_t quality_alpha = 0, speed = 9, max_threads = 1, tile_rows_log2 = 0, tile_cols_log2 = 0;
    bool auto_tiling = true;
    double score_tgt = 80.0;
    bool tenbit = true;
    std::string tune = "iq";
    double tolerance = 2.0;
    uint32_t max_pass = 6;
    int quality = -1;  // -q bypass when >= 0
    uint32_t color_primaries = 2, transfer_characteristics = 2, matrix_coefficients = 2;
};

struct HostImage {  // the subset of io.zig's Image the harness produces: 8-bit, 3 or 4 channels
    uint32_t w = 0, h = 0, channels = 3;
    std::vector<uint8_t> data;  // interleaved, tight rows
    size_t file_bytes = 0;      // "Original Bytes" of the measure.py CSV
    std::string name;
};

HostImage load_pnm(const std::string &path);                       // P6 / P7 (RGB, RGB_ALPHA), MAXVAL 255
HostImage synth_image(uint32_t w, uint32_t h, uint32_t kind, uint64_t seed, bool alpha = false);
std::vector<uint8_t> to_rgb8(const HostImage &img);                // Image.toRGB8, io.zig:57-133 (8-bit cases)

struct Decoded {  // a decoder kept alive so that its planes can be handed to the scorer in place
    const LibAvif *lib = nullptr;
    void *decoder = nullptr;
    AvifImageView img{nullptr};
    Decoded() = default;
    Decoded(const Decoded &) = delete;
    Decoded &operator=(const Decoded &) = delete;
    Decoded(Decoded &&o) noexcept { *this = std::move(o); }
    Decoded &operator=(Decoded &&o) noexcept
    {
        std::swap(lib, o.lib);
        std::swap(decoder, o.decoder);
        std::swap(img, o.img);
        return *this;
    }
    ~Decoded();
};

class Codec {
  public:
    explicit Codec(const LibAvif &lib) : L(lib) {}
    // io.zig:544-636.  samples10: the source already converted to 10 bits (ScorerIface::source_samples10), or null
    std::vector<uint8_t> encode(const HostImage &img, uint32_t q, const EncOptions &o,
                                const std::vector<uint16_t> *samples10 = nullptr) const;
    Decoded decode(const std::vector<uint8_t> &avif) const;                                       // io.zig:452-466
    std::vector<uint8_t> decode_to_rgb8(const std::vector<uint8_t> &avif) const;                 // io.zig:638-666
    const LibAvif &L;
};

// How one decoded candidate is scored.  The product path is the CUDA library; tests may inject
// another scorer (the CPU oracle) through this interface to compare search traces.
struct ScorerIface {
    virtual ~ScorerIface() = default;
    virtual void set_source(const uint8_t *rgb, uint32_t w, uint32_t h) = 0;
    // The source as loaded (main.zig:86-87).  Default: Image.toRGB8 on the host, kept alive for the search.
    virtual void set_source_image(const HostImage &img)
    {
        rgb_keep_ = to_rgb8(img);
        set_source(rgb_keep_.data(), img.w, img.h);
    }
    // The 10-bit sample array encodeAvifToBuffer rebuilds from the source in EVERY pass (io.zig:566-579), made once
    // per image from the pixels the scorer staged.  false: this scorer keeps no device copy, and Codec::encode runs
    // the reference's loop on the host in each pass.
    virtual bool source_samples10(std::vector<uint16_t> &out)
    {
        (void)out;
        return false;
    }
    virtual std::vector<double> score(const std::vector<const Decoded *> &cands) = 0;
  protected:
    std::vector<uint8_t> rgb_keep_;
};

class GpuScorer : public ScorerIface {  // include/oavif_ssimu2.h
  public:
    // pinned_staging (default): copy the decoder's planes into a per-scorer pinned ring (oavif_ssimu2_pinned_alloc —
    // the "pinned host staging in src/io.zig" of the north star) and upload from there by DMA; off: hand libavif's
    // pageable planes to the library, whose cudaMemcpy2DAsync stages them through the driver's own pinned buffers.
    // Measured four times on the config-5 corpus with 16 workers (profiles/r2_decode_handoff.json), each on a fresh
    // box: scoring wall time per image 10.1 / 4.6 / 4.5 / 3.2 ms pinned against 5.6 / 7.2 / 17.0 / 14.4 ms pageable,
    // 25.0 / 26.2 / 26.0 / 27.0 against 26.2 / 26.7 / 24.9 / 26.3 encodes/s.  Throughput is the same within the
    // run-to-run spread (scoring is ~1 % of a search); the pinned form's scoring time is the lower and steadier one
    // in three runs of four, so it is the default.
    GpuScorer(int device, uint32_t max_w, uint32_t max_h, uint32_t max_batch, int blur_mode, bool pinned_staging = true);
    ~GpuScorer() override;
    void set_source(const uint8_t *rgb, uint32_t w, uint32_t h) override;
    void set_source_image(const HostImage &img) override;      // oavif_ssimu2_set_source_pixels: toRGB8 on the device
    bool source_samples10(std::vector<uint16_t> &out) override; // oavif_ssimu2_source_samples
    std::vector<double> score(const std::vector<const Decoded *> &cands) override;
    double device_ms = 0.0;  // accumulated total_ms of the calls
  private:
    oavif_ssimu2_ctx *ctx_ = nullptr;
    uint32_t max_batch_;
    bool pinned_;
    uint8_t *stage_ = nullptr;   // pinned: max_batch x 3 planes
    size_t stage_bytes_ = 0, plane_cap_ = 0;
    size_t src_samples_ = 0;     // w * h * channels of the staged source
};

struct SearchResult {
    TQResult tq;
    BatchedStats batched;
    std::vector<DecisionMargin> margins;   // tq.hpp: how far each pass's score was from changing the search
    std::vector<uint8_t> avif;   // the bytes oavif would write
    size_t size = 0;             // e.buf.size as printed by main.zig:116
    bool reencoded = false;      // main.zig:113 path (extra encode, not counted in num_pass)
    uint32_t out_depth = 8;
    double encode_ms = 0, decode_ms = 0, score_ms = 0, total_ms = 0;
    std::string log;             // the stderr lines of main.zig:78-116
};

// main.zig:86-116 for one image: toRGB8, (bypass | search), write/re-encode.  batch_width <= 1 is the
// reference's sequential loop; > 1 scores speculative candidates together (same decisions, tq.hpp).
SearchResult search_image(const Codec &codec, ScorerIface &scorer, const HostImage &img, const EncOptions &o,
                          uint32_t batch_width, uint32_t host_threads);

struct CorpusRow {  // measure.py:178-206, plus what the trace file carries
    std::string image, status, error;
    size_t orig_bytes = 0, final_bytes = 0;
    double encoding_time_ms = 0;
    uint32_t passes = 0, q = 0;
    double score = 0;
    int gpu = -1, worker = -1;
    double encode_ms = 0, decode_ms = 0, score_ms = 0;   // per-stage host wall time of this image's search
    double margin = 0;                                    // min over passes of min(flip_up, flip_down) (tq.hpp)
    std::string trace;                                    // "q:score q:score ..." in probe order
};

struct CorpusSpec {
    std::vector<std::string> files;       // PAM/PPM inputs, or
    uint32_t synth_count = 0, synth_w = 1920, synth_h = 1080;  // procedural: seed = index, kind = seed mod 4
    int first_gpu = 0, n_gpus = 1;
    uint32_t workers_per_gpu = 1, batch_width = 1;
    int blur_mode = 0;
    bool pinned_staging = true;    // see GpuScorer
    // The CPU-scored arm: when set, every worker scores through this factory instead of the CUDA library (tests and
    // bench tooling inject the CPU oracle here; the product never does).  n_gpus * workers_per_gpu workers still.
    std::function<std::unique_ptr<ScorerIface>(int worker)> scorer_factory;
};

struct CorpusStats {
    double wall_s = 0;
    double scorer_device_ms = 0;     // sum over workers of the device time of their score calls (GPU arm)
    uint32_t workers = 0, host_cpus = 0;
};

// scripts/measure.py as a library call.  Workers (n_gpus x workers_per_gpu host threads, one scorer context
// each, each bound to its GPU) pull image indices from ONE shared atomic counter — per-image time varies
// 250-830 ms x 2-6 passes, so a static i mod (G*W) shard leaves GPUs idle at the end — and there is no
// collective: rows come back in image order whatever worker produced them.
std::vector<CorpusRow> run_corpus(const std::string &libavif_path, const CorpusSpec &spec, const EncOptions &o,
                                  CorpusStats *stats);
std::string corpus_csv(const std::vector<CorpusRow> &rows);            // the nine columns of measure.py:180-192
std::string corpus_trace_csv(const std::vector<CorpusRow> &rows);      // per image: q, score, passes, stage times, margin, trace
std::string corpus_summary(const std::vector<CorpusRow> &rows, const CorpusStats &stats);   // measure.py:250-269 + extras

}  // namespace oavif_host
