// oavif_host.cpp — see oavif_host.hpp.  Links against liboavif_ssimu2.so (the CUDA scorer) and
// dlopens libavif.  Nothing here computes a score on the CPU.
#include "oavif_host.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <cstdarg>
#include <cstring>
#include <sched.h>
#include <unistd.h>

#include "../../../include/oavif_ssimu2.h"

namespace oavif_host {

namespace {
double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}
std::string fmt(const char *f, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof buf, f, ap);
    va_end(ap);
    return buf;
}
}  // namespace

// ------------------------------------------------------------------------------------------------
// images

// NB: no <iostream>/<fstream>/<sstream> anywhere in the harness: the toolchain links libstdc++
// statically into this shared object, and a second copy of the iostream locale machinery inside a
// host process that already has one (CPython + ctypes) is not initialised reliably.
HostImage load_pnm(const std::string &path)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    struct Closer {
        FILE *f;
        ~Closer() { fclose(f); }
    } closer{f};
    HostImage img;
    img.name = path.substr(path.find_last_of('/') + 1);
    auto token = [&]() -> std::string {  // whitespace-separated token, '#' comments skipped
        std::string t;
        int ch;
        for (;;) {
            ch = fgetc(f);
            if (ch == EOF) return t;
            if (ch == '#' && t.empty()) {
                while (ch != '\n' && ch != EOF) ch = fgetc(f);
                continue;
            }
            if (ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r') {
                if (!t.empty()) return t;
                continue;
            }
            t.push_back((char)ch);
        }
    };
    const std::string magic = token();
    uint32_t maxval = 0;
    if (magic == "P6") {
        img.w = (uint32_t)std::stoul(token());
        img.h = (uint32_t)std::stoul(token());
        maxval = (uint32_t)std::stoul(token());  // token() consumed the single whitespace after MAXVAL
        img.channels = 3;
    } else if (magic == "P7") {  // io.zig:309-406
        uint32_t depth = 0;
        for (;;) {
            const std::string key = token();
            if (key.empty()) throw std::runtime_error("PAM: truncated header");
            if (key == "ENDHDR") break;
            if (key == "WIDTH") img.w = (uint32_t)std::stoul(token());
            else if (key == "HEIGHT") img.h = (uint32_t)std::stoul(token());
            else if (key == "DEPTH") depth = (uint32_t)std::stoul(token());
            else if (key == "MAXVAL") maxval = (uint32_t)std::stoul(token());
            else if (key == "TUPLTYPE") (void)token();
        }
        if (depth != 3 && depth != 4) throw std::runtime_error("PAM: only RGB / RGB_ALPHA supported by the harness");
        img.channels = depth;
    } else {
        throw std::runtime_error("unsupported image format (harness reads P6/P7 only): " + path);
    }
    if (maxval != 255 || img.w == 0 || img.h == 0) throw std::runtime_error("PNM: need MAXVAL 255");
    img.data.resize((size_t)img.w * img.h * img.channels);
    if (fread(img.data.data(), 1, img.data.size(), f) != img.data.size()) throw std::runtime_error("PNM: short pixel data");
    fseek(f, 0, SEEK_END);
    img.file_bytes = (size_t)ftell(f);
    return img;
}

static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// Procedural corpus image (gradients / edges / textured noise / mixture): the harness's own
// deterministic definition — a pure function of (w, h, kind, seed, x, y).
HostImage synth_image(uint32_t w, uint32_t h, uint32_t kind, uint64_t seed, bool alpha)
{
    HostImage img;
    img.w = w;
    img.h = h;
    img.channels = alpha ? 4 : 3;
    img.data.resize((size_t)w * h * img.channels);
    img.name = fmt("synth_%05llu_k%u_%ux%u", (unsigned long long)seed, kind & 3, w, h);
    double p[12];
    for (int i = 0; i < 12; ++i) p[i] = (double)(splitmix64(seed * 1315423911ull + i) >> 11) / 9007199254740992.0;
    const int per[3] = {8 << (int)(p[0] * 4.999), 8 << (int)(p[1] * 4.999), 8 << (int)(p[2] * 4.999)};
    std::vector<float> noise((size_t)w * h * 3);
    for (size_t i = 0; i < noise.size(); ++i) {
        const uint64_t r = splitmix64(seed * 0x100000001B3ull + i);
        float g = 0.f;
        for (int k = 0; k < 4; ++k) g += (float)((r >> (16 * k)) & 0xffff) / 65536.0f;
        noise[i] = (g - 2.0f) * 20.0f;
    }
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x) {
            const double u = (double)x / std::max(1u, w - 1), v = (double)y / std::max(1u, h - 1);
            double px[3];
            for (int c = 0; c < 3; ++c) {
                const double grad = 255.0 * (p[3 + c] * (1 - u) * (1 - v) + p[6 + c] * u * (1 - v) + p[9 + c] * (1 - u) * v +
                                             (1.0 - p[3 + c]) * u * v);
                const bool on = c == 0 ? (((x / per[0]) + (y / per[0])) & 1) : c == 1 ? ((x / per[1]) & 1) : ((y / per[2]) & 1);
                const double edge = on ? 220.0 - 60.0 * p[c] : 30.0 + 60.0 * p[c + 1];
                // 3x3 box of the white noise = band-limited texture
                double nz = 0;
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        const uint32_t yy = std::min<int64_t>(std::max<int64_t>((int64_t)y + dy, 0), h - 1);
                        const uint32_t xx = std::min<int64_t>(std::max<int64_t>((int64_t)x + dx, 0), w - 1);
                        nz += noise[((size_t)yy * w + xx) * 3 + c];
                    }
                nz /= 3.0;
                const double tex = 0.6 * grad + 50.0 + nz;
                double val;
                switch (kind & 3) {
                case 0: val = grad; break;
                case 1: val = edge; break;
                case 2: val = tex; break;
                default: val = (u + v < 0.7) ? grad : (u - v > 0.1 ? edge : tex); val = 0.85 * val + 0.15 * tex; break;
                }
                px[c] = val;
            }
            uint8_t *o = &img.data[((size_t)y * w + x) * img.channels];
            for (int c = 0; c < 3; ++c) o[c] = (uint8_t)std::min(255.0, std::max(0.0, std::floor(px[c] + 0.5)));
            if (alpha) {
                const double dx = (x - w / 2.0) / (w / 2.0), dy = (y - h / 2.0) / (h / 2.0);
                o[3] = (uint8_t)std::min(255.0, std::max(0.0, 255.0 * (1.2 - std::sqrt(dx * dx + dy * dy))));
            }
        }
    img.file_bytes = img.data.size();
    return img;
}

std::vector<uint8_t> to_rgb8(const HostImage &img)
{  // io.zig:57-133, 8-bit branches (main.zig:86 aliases the data when channels == 3)
    if (img.channels == 3) return img.data;
    std::vector<uint8_t> rgb((size_t)img.w * img.h * 3);
    for (size_t i = 0, n = (size_t)img.w * img.h; i < n; ++i) {
        rgb[3 * i + 0] = img.data[4 * i + 0];
        rgb[3 * i + 1] = img.data[4 * i + 1];
        rgb[3 * i + 2] = img.data[4 * i + 2];
    }
    return rgb;
}

// ------------------------------------------------------------------------------------------------
// libavif glue

Decoded::~Decoded()
{
    if (decoder && lib) lib->avifDecoderDestroy(decoder);
}

std::vector<uint8_t> Codec::encode(const HostImage &src, uint32_t q, const EncOptions &o,
                                   const std::vector<uint16_t> *samples10) const
{
    const uint32_t depth = o.tenbit ? 10 : 8;  // io.zig:546 with an 8-bit source
    void *image = L.avifImageCreate(src.w, src.h, depth, 1 /* AVIF_PIXEL_FORMAT_YUV444 */);
    if (!image) throw std::runtime_error("avifImageCreate failed");
    struct ImgGuard {
        const LibAvif &L;
        void *p;
        ~ImgGuard() { L.avifImageDestroy(p); }
    } ig{L, image};
    AvifImageView v{static_cast<uint8_t *>(image)};
    v.colorPrimaries() = (uint16_t)o.color_primaries;
    v.transferCharacteristics() = (uint16_t)o.transfer_characteristics;
    v.matrixCoefficients() = (uint16_t)o.matrix_coefficients;

    AvifRGBImage rgb{};
    L.avifRGBImageSetDefaults(&rgb, image);
    rgb.format = src.channels == 4 ? 1 : 0;
    std::vector<uint16_t> scaled;
    if (depth == 10) {  // io.zig:566-579
        if (samples10 && samples10->size() == src.data.size()) {   // converted once per image on the device
            rgb.pixels = reinterpret_cast<uint8_t *>(const_cast<uint16_t *>(samples10->data()));
        } else {
            scaled.resize(src.data.size());
            for (size_t i = 0; i < scaled.size(); ++i) scaled[i] = (uint16_t)(((size_t)src.data[i] * 1023 + 127) / 255);
            rgb.pixels = reinterpret_cast<uint8_t *>(scaled.data());
        }
        rgb.rowBytes = src.w * src.channels * 2;
        rgb.depth = 10;
    } else {  // io.zig:611-617
        rgb.pixels = const_cast<uint8_t *>(src.data.data());
        rgb.rowBytes = src.w * src.channels;
        rgb.depth = 8;
    }
    if (L.avifImageRGBToYUV(image, &rgb) != 0) throw std::runtime_error("ConvertFailed");

    uint8_t *enc = static_cast<uint8_t *>(L.avifEncoderCreate());
    if (!enc) throw std::runtime_error("avifEncoderCreate failed");
    struct EncGuard {
        const LibAvif &L;
        void *p;
        ~EncGuard() { L.avifEncoderDestroy(p); }
    } eg{L, enc};
    auto i32 = [&](size_t off) -> int32_t & { return *reinterpret_cast<int32_t *>(enc + off); };
    // copyToEncoder, parse_args.zig:65-74
    i32(kEncQualityAlpha) = (int32_t)o.quality_alpha;
    i32(kEncSpeed) = (int32_t)o.speed;
    i32(kEncMaxThreads) = (int32_t)o.max_threads;
    i32(kEncTileRowsLog2) = (int32_t)o.tile_rows_log2;
    i32(kEncTileColsLog2) = (int32_t)o.tile_cols_log2;
    i32(kEncAutoTiling) = o.auto_tiling ? 1 : 0;
    if (L.avifEncoderSetCodecSpecificOption(enc, "tune", o.tune.c_str()) != 0) throw std::runtime_error("InvalidCodecOption");
    i32(kEncQuality) = (int32_t)q;  // io.zig:625-626
    i32(kEncQualityAlpha) = (int32_t)o.quality_alpha;
    AvifRWData out{nullptr, 0};
    int rc = L.avifEncoderAddImage(enc, image, 1, 2 /* AVIF_ADD_IMAGE_FLAG_SINGLE */);
    if (rc != 0) throw std::runtime_error(std::string("AddImageFailed: ") + L.avifResultToString(rc));
    rc = L.avifEncoderFinish(enc, &out);
    if (rc != 0) throw std::runtime_error(std::string("FinishFailed: ") + L.avifResultToString(rc));
    std::vector<uint8_t> bytes(out.data, out.data + out.size);
    L.avifRWDataFree(&out);
    return bytes;
}

Decoded Codec::decode(const std::vector<uint8_t> &avif) const
{
    Decoded d;
    d.lib = &L;
    d.decoder = L.avifDecoderCreate();
    if (!d.decoder) throw std::runtime_error("avifDecoderCreate failed");
    if (L.avifDecoderSetIOMemory(d.decoder, avif.data(), avif.size()) != 0) throw std::runtime_error("SetIOFailed");
    if (L.avifDecoderParse(d.decoder) != 0) throw std::runtime_error("ParseFailed");
    if (L.avifDecoderNextImage(d.decoder) != 0) throw std::runtime_error("DecodeImageFailed");
    d.img = AvifImageView{*reinterpret_cast<uint8_t **>(static_cast<uint8_t *>(d.decoder) + kDecImage)};
    if (!d.img.p || d.img.yuvFormat() != 1) throw std::runtime_error("decoder->image is not YUV444");
    return d;
}

std::vector<uint8_t> Codec::decode_to_rgb8(const std::vector<uint8_t> &avif) const
{  // the reference's own decode path (io.zig:468-481, 638-666), kept for comparisons
    Decoded d = decode(avif);
    AvifRGBImage rgb{};
    L.avifRGBImageSetDefaults(&rgb, d.img.p);
    rgb.depth = 8;
    rgb.format = d.img.alphaPlane() ? 1 : 0;
    if (L.avifRGBImageAllocatePixels(&rgb) != 0) throw std::runtime_error("AllocatePixelsFailed");
    if (L.avifImageYUVToRGB(d.img.p, &rgb) != 0) {
        L.avifRGBImageFreePixels(&rgb);
        throw std::runtime_error("ConvertToRGBFailed");
    }
    const uint32_t w = d.img.width(), h = d.img.height(), ch = rgb.format == 1 ? 4 : 3;
    std::vector<uint8_t> out((size_t)w * h * 3);
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t *row = rgb.pixels + (size_t)y * rgb.rowBytes;
        for (uint32_t x = 0; x < w; ++x)
            for (int c = 0; c < 3; ++c) out[((size_t)y * w + x) * 3 + c] = row[x * ch + c];
    }
    L.avifRGBImageFreePixels(&rgb);
    return out;
}

// ------------------------------------------------------------------------------------------------
// scorer

GpuScorer::GpuScorer(int device, uint32_t max_w, uint32_t max_h, uint32_t max_batch, int blur_mode, bool pinned_staging)
    : max_batch_(max_batch), pinned_(pinned_staging)
{
    if (oavif_ssimu2_ctx_create(device, max_w, max_h, max_batch, &ctx_) != 0)
        throw std::runtime_error(std::string("oavif_ssimu2_ctx_create: ") + oavif_ssimu2_last_error(nullptr));
    oavif_ssimu2_set_option(ctx_, OAVIF_SSIMU2_OPT_BLUR, blur_mode);
    if (pinned_) {
        plane_cap_ = ((size_t)max_w * max_h * 2 + 255) & ~(size_t)255;   // up to 10-bit samples
        stage_bytes_ = plane_cap_ * 3 * max_batch;
        stage_ = static_cast<uint8_t *>(oavif_ssimu2_pinned_alloc(stage_bytes_));
        if (!stage_) pinned_ = false;   // pinning refused (ulimit): fall back to the pageable hand-off
    }
}

GpuScorer::~GpuScorer()
{
    oavif_ssimu2_ctx_destroy(ctx_);
    oavif_ssimu2_pinned_free(stage_);
}

void GpuScorer::set_source(const uint8_t *rgb, uint32_t w, uint32_t h)
{
    src_samples_ = 0;
    if (oavif_ssimu2_set_source_rgb8(ctx_, rgb, w, h, (size_t)w * 3) != 0)
        throw std::runtime_error(std::string("set_source: ") + oavif_ssimu2_last_error(ctx_));
}

void GpuScorer::set_source_image(const HostImage &img)
{
    if (oavif_ssimu2_set_source_pixels(ctx_, img.data.data(), img.w, img.h, (size_t)img.w * img.channels, (int)img.channels, 8) != 0)
        throw std::runtime_error(std::string("set_source_pixels: ") + oavif_ssimu2_last_error(ctx_));
    src_samples_ = img.data.size();
}

bool GpuScorer::source_samples10(std::vector<uint16_t> &out)
{
    if (src_samples_ == 0) return false;
    out.resize(src_samples_);
    if (oavif_ssimu2_source_samples(ctx_, 10, out.data(), out.size() * sizeof(uint16_t)) != 0) {
        out.clear();   // e.g. an image below 8x8, which the library does not stage: the host loop converts it
        return false;
    }
    return true;
}

std::vector<double> GpuScorer::score(const std::vector<const Decoded *> &cands)
{
    std::vector<double> out(cands.size());
    for (size_t base = 0; base < cands.size(); base += max_batch_) {
        const size_t n = std::min<size_t>(max_batch_, cands.size() - base);
        std::vector<const void *> y(n), u(n), v(n);
        const AvifImageView &i0 = cands[base]->img;
        size_t strides[3] = {i0.rowBytes(0), i0.rowBytes(1), i0.rowBytes(2)};
        for (size_t i = 0; i < n; ++i) {
            const AvifImageView &im = cands[base + i]->img;
            if (im.depth() != i0.depth() || im.rowBytes(0) != i0.rowBytes(0) || im.rowBytes(1) != i0.rowBytes(1) ||
                im.rowBytes(2) != i0.rowBytes(2))
                throw std::runtime_error("batched candidates differ in layout");
            // The C ABI converts full-range planes only and has no range argument: refuse anything else here
            // rather than score it silently wrong (avifImageCreate's default is full range, io.zig:546).
            if (im.yuvRange() != 1) throw std::runtime_error("decoded image is limited-range YUV: not on the scored path");
            y[i] = im.plane(0);
            u[i] = im.plane(1);
            v[i] = im.plane(2);
        }
        if (pinned_) {   // decode hand-off through pinned staging: tight rows, one memcpy per plane row
            const size_t rb = (size_t)i0.width() * (i0.depth() > 8 ? 2 : 1);
            const uint32_t h = i0.height();
            if (rb * h <= plane_cap_) {
                for (size_t i = 0; i < n; ++i) {
                    const void **pl[3] = {&y[i], &u[i], &v[i]};
                    for (int p = 0; p < 3; ++p) {
                        uint8_t *dst = stage_ + (i * 3 + p) * plane_cap_;
                        const uint8_t *src = static_cast<const uint8_t *>(*pl[p]);
                        if (strides[p] == rb) memcpy(dst, src, rb * h);
                        else
                            for (uint32_t r = 0; r < h; ++r) memcpy(dst + r * rb, src + r * strides[p], rb);
                        *pl[p] = dst;
                    }
                }
                strides[0] = strides[1] = strides[2] = rb;
            }
        }
        const int rc = oavif_ssimu2_score_batch_yuv444(ctx_, (uint32_t)n, y.data(), u.data(), v.data(), strides[0],
                                                       strides[1], strides[2], (int)i0.depth(),
                                                       (int)i0.matrixCoefficients(), i0.alphaPlane() != nullptr,
                                                       out.data() + base);
        if (rc != 0) throw std::runtime_error(std::string("score: ") + oavif_ssimu2_last_error(ctx_));
        oavif_ssimu2_timing t;
        oavif_ssimu2_get_timing(ctx_, &t);
        device_ms += t.total_ms;
    }
    return out;
}

// ------------------------------------------------------------------------------------------------
// one image: main.zig:86-116

SearchResult search_image(const Codec &codec, ScorerIface &scorer, const HostImage &img, const EncOptions &o,
                          uint32_t batch_width, uint32_t host_threads)
{
    SearchResult R;
    const double t_start = now_ms();
    R.out_depth = o.tenbit ? 10 : 8;
    if (o.quality >= 0) {  // main.zig:93-100
        R.log += fmt("Encoding [q%d, speed %u, %u-bit]\n", o.quality, o.speed, R.out_depth);
        R.avif = codec.encode(img, (uint32_t)o.quality, o);
        R.size = R.avif.size();
        R.tq.q = (uint32_t)o.quality;
        R.log += fmt("Compressed to %zu bytes (%.3f bpp)\n", R.size, (double)(R.size * 8) / ((double)img.w * img.h));
        R.total_ms = now_ms() - t_start;
        return R;
    }
    R.log += fmt("Searching [tgt %g±%.1f, speed %u, %u-bit]\n", o.score_tgt, o.tolerance, o.speed, R.out_depth);
    scorer.set_source_image(img);
    // io.zig:566-579 converts the source to 10 bits in every pass; a scorer that staged the pixels does it once
    std::vector<uint16_t> samples10;
    const std::vector<uint16_t> *s10 = (o.tenbit && scorer.source_samples10(samples10)) ? &samples10 : nullptr;

    TQOptions topt;
    topt.score_tgt = o.score_tgt;
    topt.tolerance = o.tolerance;
    topt.max_pass = o.max_pass;

    // EncBuffer (main.zig:11-23): the bytes of the LAST pass the policy consumed
    uint32_t buf_q = 0;
    std::vector<uint8_t> buf;
    std::vector<std::pair<uint32_t, std::vector<uint8_t>>> spec_bytes;  // speculative encodes kept by q
    std::mutex mu;

    auto probe_many = [&](const std::vector<uint32_t> &qs) {
        std::vector<std::vector<uint8_t>> bytes(qs.size());
        std::vector<Decoded> dec(qs.size());
        std::vector<std::string> errs(qs.size());
        double enc_ms = 0, dec_ms = 0;
        auto work = [&](size_t i) {
            try {
                const double t0 = now_ms();
                bytes[i] = codec.encode(img, qs[i], o, s10);
                const double t1 = now_ms();
                dec[i] = codec.decode(bytes[i]);
                const double t2 = now_ms();
                std::lock_guard<std::mutex> lk(mu);
                enc_ms += t1 - t0;
                dec_ms += t2 - t1;
            } catch (const std::exception &e) {
                errs[i] = e.what();
            }
        };
        if (qs.size() == 1 || host_threads <= 1) {
            for (size_t i = 0; i < qs.size(); ++i) work(i);
        } else {
            std::vector<std::thread> th;
            std::atomic<size_t> next{0};
            const size_t nt = std::min<size_t>(host_threads, qs.size());
            for (size_t t = 0; t < nt; ++t)
                th.emplace_back([&] {
                    for (size_t i; (i = next.fetch_add(1)) < qs.size();) work(i);
                });
            for (auto &t : th) t.join();
        }
        for (const auto &e : errs)
            if (!e.empty()) throw std::runtime_error(e);
        R.encode_ms += enc_ms;
        R.decode_ms += dec_ms;
        std::vector<const Decoded *> ptrs;
        for (auto &d : dec) ptrs.push_back(&d);
        const double t0 = now_ms();
        std::vector<double> sc = scorer.score(ptrs);
        R.score_ms += now_ms() - t0;
        for (size_t i = 0; i < qs.size(); ++i) spec_bytes.emplace_back(qs[i], std::move(bytes[i]));
        return sc;
    };
    auto take_bytes = [&](uint32_t q) {  // tq.zig:31-35: the consumed pass becomes the cached buffer
        for (auto &p : spec_bytes)
            if (p.first == q) {
                buf = p.second;
                buf_q = q;
                return;
            }
    };

    if (batch_width <= 1) {
        R.tq = findTargetQuality(topt, [&](uint32_t q) {
            const double s = probe_many({q})[0];
            take_bytes(q);
            return s;
        });
    } else {
        R.tq = findTargetQualityBatched(topt, batch_width, probe_many, &R.batched);
        // the reference's cache holds the last pass of the sequential procedure
        if (!R.tq.history.empty()) take_bytes(R.tq.history.back().q);
    }
    R.margins = decisionMargins(topt, R.tq.history);
    R.log += fmt("Found q%u (score %.2f, %u passes)\n", R.tq.q, R.tq.score, R.tq.num_pass);
    if (buf_q == R.tq.q && !buf.empty()) {  // main.zig:109-112
        R.avif = buf;
        R.size = buf.size();
    } else {  // main.zig:113 -> encodeAvifToFile: one more encode, not counted in num_pass
        bool have = false;
        if (batch_width > 1)
            for (auto &p : spec_bytes)
                if (p.first == R.tq.q) {  // deterministic encoder: the speculative bytes ARE that encode
                    R.avif = p.second;
                    have = true;
                }
        if (!have) R.avif = codec.encode(img, R.tq.q, o, s10);
        R.size = R.avif.size();
        R.reencoded = true;
    }
    R.log += fmt("Compressed to %zu bytes (%.3f bpp)\n", R.size, (double)(R.size * 8) / ((double)img.w * img.h));
    R.total_ms = now_ms() - t_start;
    return R;
}

// ------------------------------------------------------------------------------------------------
// corpus driver: scripts/measure.py

namespace {
// CPUs the kernel reports as local to a GPU's PCI function; empty when sysfs has nothing useful
std::vector<int> gpu_local_cpus(int device)
{
    std::vector<int> cpus;
    char id[64] = {0};
    if (oavif_ssimu2_device_pci_bus_id(device, id, sizeof id) != 0) return cpus;
    const std::string path = std::string("/sys/bus/pci/devices/") + id + "/local_cpulist";
    FILE *f = fopen(path.c_str(), "r");
    if (!f) return cpus;
    char buf[512] = {0};
    const bool ok = fgets(buf, sizeof buf, f) != nullptr;
    fclose(f);
    if (!ok) return cpus;
    for (char *tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k == 1) b = a;
        if (k >= 1)
            for (int c = a; c <= b; ++c) cpus.push_back(c);
    }
    return cpus;
}

// Bind the calling worker next to its GPU before it allocates pinned staging: the GPU's local CPUs when sysfs
// names a proper subset of the machine, otherwise leave the scheduler alone (one visible NUMA node).
void bind_worker_near_gpu(int device)
{
    const std::vector<int> cpus = gpu_local_cpus(device);
    const long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    if (cpus.empty() || (long)cpus.size() >= ncpu) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    for (int c : cpus) CPU_SET(c, &set);
    sched_setaffinity(0, sizeof set, &set);
}
}  // namespace

std::vector<CorpusRow> run_corpus(const std::string &libavif_path, const CorpusSpec &spec, const EncOptions &o,
                                  CorpusStats *stats)
{
    const size_t n = spec.files.empty() ? spec.synth_count : spec.files.size();
    std::vector<CorpusRow> rows(n);
    LibAvif lib(libavif_path);
    Codec codec(lib);
    const int G = std::max(1, spec.n_gpus);
    const uint32_t W = std::max(1u, spec.workers_per_gpu);
    const double t0 = now_ms();
    std::vector<std::thread> workers;
    std::vector<double> device_ms((size_t)G * W, 0.0);
    std::atomic<size_t> next{0};   // the shared work counter: whoever is free takes the next image
    for (int g = 0; g < G; ++g)
        for (uint32_t wk = 0; wk < W; ++wk)
            workers.emplace_back([&, g, wk] {
                const int widx = g * (int)W + (int)wk;
                const bool gpu_arm = !spec.scorer_factory;
                if (gpu_arm) bind_worker_near_gpu(spec.first_gpu + g);
                std::unique_ptr<ScorerIface> scorer;  // one context per worker, sized on first use
                uint32_t cap_w = 0, cap_h = 0;
                for (size_t i; (i = next.fetch_add(1)) < n;) {
                    CorpusRow &r = rows[i];
                    r.gpu = gpu_arm ? spec.first_gpu + g : -1;
                    r.worker = widx;
                    try {
                        HostImage img = spec.files.empty()
                                            ? synth_image(spec.synth_w, spec.synth_h, (uint32_t)(i & 3), i)
                                            : load_pnm(spec.files[i]);
                        r.image = img.name;
                        r.orig_bytes = img.file_bytes;
                        if (!scorer || img.w > cap_w || img.h > cap_h) {
                            cap_w = std::max(cap_w, img.w);
                            cap_h = std::max(cap_h, img.h);
                            if (auto *gs = dynamic_cast<GpuScorer *>(scorer.get())) device_ms[widx] += gs->device_ms;
                            scorer.reset();
                            if (gpu_arm)
                                scorer.reset(new GpuScorer(spec.first_gpu + g, cap_w, cap_h, std::max(1u, spec.batch_width),
                                                           spec.blur_mode, spec.pinned_staging));
                            else
                                scorer = spec.scorer_factory(widx);
                        }
                        const double te = now_ms();  // measure.py times the oavif process: load excluded is closest
                        SearchResult sr = search_image(codec, *scorer, img, o, spec.batch_width, spec.batch_width);
                        r.encoding_time_ms = now_ms() - te;
                        r.final_bytes = sr.size;
                        r.passes = sr.tq.num_pass;
                        r.q = sr.tq.q;
                        r.score = sr.tq.score;
                        r.encode_ms = sr.encode_ms;
                        r.decode_ms = sr.decode_ms;
                        r.score_ms = sr.score_ms;
                        r.margin = minMargin(sr.margins);
                        for (const auto &h : sr.tq.history) r.trace += fmt("%s%u:%.6f", r.trace.empty() ? "" : " ", h.q, h.score);
                        r.status = sr.avif.empty() ? "no-output" : "ok";   // measure.py:90
                    } catch (const std::exception &e) {  // one bad image must not kill the sweep
                        r.status = "error";
                        r.error = e.what();
                        if (r.image.empty()) r.image = spec.files.empty() ? fmt("synth_%05zu", i) : spec.files[i];
                    }
                }
                if (auto *gs = dynamic_cast<GpuScorer *>(scorer.get())) device_ms[widx] += gs->device_ms;
            });
    for (auto &t : workers) t.join();
    if (stats) {
        stats->wall_s = (now_ms() - t0) / 1e3;
        stats->workers = (uint32_t)(G * W);
        stats->host_cpus = std::thread::hardware_concurrency();
        stats->scorer_device_ms = 0;
        for (double d : device_ms) stats->scorer_device_ms += d;
    }
    return rows;
}

std::string corpus_csv(const std::vector<CorpusRow> &rows)
{  // header and column formats of measure.py:180-206
    std::string s = "Image,Original Bytes,Final Bytes,Savings Bytes,Savings %,Encoding Time (ms),Passes,Status,Error\r\n";
    for (const auto &r : rows) {
        if (r.status == "ok") {
            const size_t sav = r.orig_bytes > r.final_bytes ? r.orig_bytes - r.final_bytes : 0;
            const double pct = r.orig_bytes ? 100.0 * (double)sav / (double)r.orig_bytes : 0.0;
            s += r.image + fmt(",%zu,%zu,%zu,%.2f,%.2f,%u,ok,\r\n", r.orig_bytes, r.final_bytes, sav, pct,
                               r.encoding_time_ms, r.passes);
        } else if (r.status == "no-output") {   // measure.py:70-91: time and passes known, no file
            s += r.image + fmt(",%zu,,,,%.2f,%u,no-output,\r\n", r.orig_bytes, r.encoding_time_ms, r.passes);
        } else {
            std::string e = r.error;
            std::replace(e.begin(), e.end(), ',', ';');
            s += r.image + fmt(",%zu,,,,,,", r.orig_bytes) + r.status + "," + e + "\r\n";
        }
    }
    return s;
}

std::string corpus_trace_csv(const std::vector<CorpusRow> &rows)
{
    std::string s = "Image,Q,Score,Passes,Final Bytes,Encode ms,Decode ms,Score ms,GPU,Worker,Margin,Trace\r\n";
    for (const auto &r : rows)
        s += r.image + fmt(",%u,%.6f,%u,%zu,%.2f,%.2f,%.3f,%d,%d,%.6f,", r.q, r.score, r.passes, r.final_bytes, r.encode_ms,
                           r.decode_ms, r.score_ms, r.gpu, r.worker, r.margin) + r.trace + "\r\n";
    return s;
}

namespace {
std::string human_bytes(double n)
{  // measure.py:31-38
    const char *units[] = {"B", "KiB", "MiB", "GiB", "TiB"};
    double size = n;
    for (int i = 0; i < 5; ++i) {
        if (size < 1024.0 || i == 4) return fmt("%.2f %s", size, units[i]);
        size /= 1024.0;
    }
    return "";
}
}  // namespace

std::string corpus_summary(const std::vector<CorpusRow> &rows, const CorpusStats &st)
{  // measure.py:208-269, line for line (rich markup dropped), then the figures this driver adds
    const double wall_s = st.wall_s;
    std::vector<double> t, p, ratios;
    size_t ok = 0, err = 0, no_out = 0, orig = 0, fin = 0;
    double enc = 0, dec = 0, sc = 0;
    for (const auto &r : rows) {
        if (r.status == "ok") {
            ++ok;
            t.push_back(r.encoding_time_ms);
            p.push_back(r.passes);
            orig += r.orig_bytes;
            fin += r.final_bytes;
            if (r.orig_bytes > 0) ratios.push_back((double)r.final_bytes / (double)r.orig_bytes);
            enc += r.encode_ms;
            dec += r.decode_ms;
            sc += r.score_ms;
        } else if (r.status == "no-output") {
            ++no_out;
        } else {
            ++err;
        }
    }
    auto mean = [](const std::vector<double> &v) {
        double s = 0;
        for (double x : v) s += x;
        return v.empty() ? 0.0 : s / v.size();
    };
    auto stdev = [&](const std::vector<double> &v) {
        if (v.size() < 2) return 0.0;
        const double m = mean(v);
        double s = 0;
        for (double x : v) s += (x - m) * (x - m);
        return std::sqrt(s / (v.size() - 1));
    };
    auto median = [](std::vector<double> v) {
        if (v.empty()) return 0.0;
        std::sort(v.begin(), v.end());
        return v.size() % 2 ? v[v.size() / 2] : 0.5 * (v[v.size() / 2 - 1] + v[v.size() / 2]);
    };
    const size_t savings = ok && orig > fin ? orig - fin : 0;
    std::string s = "\nRun Summary\n";
    s += fmt("Images: %zu ok, %zu no-output, %zu errors\n", ok, no_out, err);
    s += fmt("Total wall time: %.2f s\n", wall_s);
    s += fmt("Throughput: %.2f images/s\n", wall_s > 0 ? ok / wall_s : 0.0);
    s += "Input bytes throughput: " + human_bytes(std::floor(wall_s > 0 ? orig / wall_s : 0.0)) + "/s\n";
    s += "Output bytes throughput: " + human_bytes(std::floor(wall_s > 0 ? fin / wall_s : 0.0)) + "/s\n";
    s += "\nCompression Totals\n";
    s += fmt("Original total bytes: %zu (", orig) + human_bytes((double)orig) + ")\n";
    s += fmt("Final total bytes:    %zu (", fin) + human_bytes((double)fin) + ")\n";
    s += fmt("Savings (bytes):      %zu (", savings) + human_bytes((double)savings) + ")\n";
    s += fmt("%% saved (overall):    %.2f%%\n", orig ? 100.0 * (double)savings / orig : 0.0);
    bool geo_ok = !ratios.empty();
    double logsum = 0;
    for (double r : ratios) {
        if (r <= 0) geo_ok = false;   // statistics.geometric_mean raises on non-positive data: the line is skipped
        else logsum += std::log(r);
    }
    if (geo_ok) s += fmt("%% saved (geometric mean across files): %.2f%%\n", (1.0 - std::exp(logsum / ratios.size())) * 100.0);
    if (!t.empty()) {
        s += "\nTiming & Passes\n";
        s += fmt("Average encoding time: %.2f ms ± %.2f\n", mean(t), stdev(t));
        s += fmt("Median encoding time:  %.2f ms\n", median(t));
        const double mx = p.empty() ? 0 : *std::max_element(p.begin(), p.end());
        const double mn = p.empty() ? 0 : *std::min_element(p.begin(), p.end());
        s += fmt("Average passes:        %.2f ± %.2f (max: %.0f, min: %.0f)\n", mean(p), stdev(p), mx, mn);
    }
    // ---- not in measure.py: where the time went, and how close any decision came to flipping ----
    s += "\nDriver (oavif-b200)\n";
    s += fmt("Workers: %u host threads on %u CPUs\n", st.workers, st.host_cpus);
    if (ok) {
        s += fmt("Per image, mean: encode %.1f ms, decode %.1f ms, score %.2f ms (host wall, incl. upload)\n", enc / ok, dec / ok,
                 sc / ok);
        s += fmt("Scoring share of the search: %.2f%%\n", 100.0 * sc / std::max(1e-9, enc + dec + sc));
    }
    if (st.scorer_device_ms > 0 && wall_s > 0)
        s += fmt("Scorer device time: %.1f ms in total = %.3f%% of wall x workers' GPUs\n", st.scorer_device_ms,
                 100.0 * st.scorer_device_ms / (wall_s * 1e3));
    const double edges[] = {1e-4, 1e-3, 1e-2, 0.05, 0.1, 0.5, 1e30};
    const char *names[] = {"<= 1e-4", "<= 1e-3", "<= 0.01", "<= 0.05", "<= 0.1", "<= 0.5", "> 0.5"};
    size_t hist[7] = {0, 0, 0, 0, 0, 0, 0};
    for (const auto &r : rows)
        if (r.status == "ok")
            for (int b = 0; b < 7; ++b)
                if (r.margin <= edges[b]) {
                    ++hist[b];
                    break;
                }
    s += "Decision margin per image (smallest score change that alters the search):";
    for (int b = 0; b < 7; ++b) s += fmt(" [%s] %zu", names[b], hist[b]);
    s += "\n";
    return s;
}

}  // namespace oavif_host
