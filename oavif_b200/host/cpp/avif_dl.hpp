// avif_dl.hpp — libavif through dlopen, with the ABI subset the reference uses declared by hand
// (there are no avif.h headers in the build image).  Mirrors the calls of /root/reference/src/io.zig:
// encodeAvifToBuffer (544-636) and decodeAvifCommon (452-466).  Field offsets are those of libavif
// 1.4.x on x86-64 and are verified at load time (`self_check`).
#pragma once

#include <dlfcn.h>

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace oavif_host {

struct AvifRWData {
    uint8_t *data;
    size_t size;
};

struct AvifRGBImage {  // 64 bytes
    uint32_t width, height, depth, format, chromaUpsampling, chromaDownsampling;
    int32_t avoidLibYUV, ignoreAlpha, alphaPremultiplied, isFloat, maxThreads, pad_;
    uint8_t *pixels;
    uint32_t rowBytes, pad2_;
};
static_assert(sizeof(AvifRGBImage) == 64, "avifRGBImage layout");

// Read-only view of an avifImage (we never allocate one ourselves: avifImageCreate does).
struct AvifImageView {
    uint8_t *p;
    uint32_t &u32(size_t off) const { return *reinterpret_cast<uint32_t *>(p + off); }
    uint16_t &u16(size_t off) const { return *reinterpret_cast<uint16_t *>(p + off); }
    uint32_t width() const { return u32(0); }
    uint32_t height() const { return u32(4); }
    uint32_t depth() const { return u32(8); }
    uint32_t yuvFormat() const { return u32(12); }
    uint32_t yuvRange() const { return u32(16); }
    uint8_t *plane(int i) const { return *reinterpret_cast<uint8_t **>(p + 24 + 8 * i); }
    uint32_t rowBytes(int i) const { return u32(48 + 4 * i); }
    uint8_t *alphaPlane() const { return *reinterpret_cast<uint8_t **>(p + 64); }
    uint16_t &colorPrimaries() const { return u16(104); }
    uint16_t &transferCharacteristics() const { return u16(106); }
    uint16_t &matrixCoefficients() const { return u16(108); }
};

enum { kEncMaxThreads = 4, kEncSpeed = 8, kEncQuality = 32, kEncQualityAlpha = 36, kEncTileRowsLog2 = 56,
       kEncTileColsLog2 = 60, kEncAutoTiling = 64, kDecImage = 48 };

class LibAvif {
  public:
    explicit LibAvif(const std::string &path)
    {
        h_ = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!h_) throw std::runtime_error(std::string("dlopen libavif: ") + dlerror());
#define OAVIF_SYM(name) name = reinterpret_cast<decltype(name)>(sym(#name))
        OAVIF_SYM(avifVersion);
        OAVIF_SYM(avifImageCreate);
        OAVIF_SYM(avifImageDestroy);
        OAVIF_SYM(avifImageRGBToYUV);
        OAVIF_SYM(avifImageYUVToRGB);
        OAVIF_SYM(avifRGBImageSetDefaults);
        OAVIF_SYM(avifRGBImageAllocatePixels);
        OAVIF_SYM(avifRGBImageFreePixels);
        OAVIF_SYM(avifEncoderCreate);
        OAVIF_SYM(avifEncoderDestroy);
        OAVIF_SYM(avifEncoderAddImage);
        OAVIF_SYM(avifEncoderFinish);
        OAVIF_SYM(avifEncoderSetCodecSpecificOption);
        OAVIF_SYM(avifRWDataFree);
        OAVIF_SYM(avifDecoderCreate);
        OAVIF_SYM(avifDecoderDestroy);
        OAVIF_SYM(avifDecoderSetIOMemory);
        OAVIF_SYM(avifDecoderParse);
        OAVIF_SYM(avifDecoderNextImage);
        OAVIF_SYM(avifResultToString);
#undef OAVIF_SYM
        self_check();
    }
    ~LibAvif()
    {
        if (h_) dlclose(h_);
    }
    LibAvif(const LibAvif &) = delete;
    LibAvif &operator=(const LibAvif &) = delete;

    const char *(*avifVersion)() = nullptr;
    void *(*avifImageCreate)(uint32_t, uint32_t, uint32_t, int) = nullptr;
    void (*avifImageDestroy)(void *) = nullptr;
    int (*avifImageRGBToYUV)(void *, const AvifRGBImage *) = nullptr;
    int (*avifImageYUVToRGB)(const void *, AvifRGBImage *) = nullptr;
    void (*avifRGBImageSetDefaults)(AvifRGBImage *, const void *) = nullptr;
    int (*avifRGBImageAllocatePixels)(AvifRGBImage *) = nullptr;
    void (*avifRGBImageFreePixels)(AvifRGBImage *) = nullptr;
    void *(*avifEncoderCreate)() = nullptr;
    void (*avifEncoderDestroy)(void *) = nullptr;
    int (*avifEncoderAddImage)(void *, const void *, uint64_t, int) = nullptr;
    int (*avifEncoderFinish)(void *, AvifRWData *) = nullptr;
    int (*avifEncoderSetCodecSpecificOption)(void *, const char *, const char *) = nullptr;
    void (*avifRWDataFree)(AvifRWData *) = nullptr;
    void *(*avifDecoderCreate)() = nullptr;
    void (*avifDecoderDestroy)(void *) = nullptr;
    int (*avifDecoderSetIOMemory)(void *, const uint8_t *, size_t) = nullptr;
    int (*avifDecoderParse)(void *) = nullptr;
    int (*avifDecoderNextImage)(void *) = nullptr;
    const char *(*avifResultToString)(int) = nullptr;

  private:
    void *sym(const char *n)
    {
        void *p = dlsym(h_, n);
        if (!p) throw std::runtime_error(std::string("libavif lacks ") + n);
        return p;
    }
    // The hand-declared offsets are only trusted if a freshly created image/encoder shows the
    // documented defaults where we expect them.
    void self_check()
    {
        void *img = avifImageCreate(37, 21, 10, 1 /* YUV444 */);
        if (!img) throw std::runtime_error("avifImageCreate failed");
        AvifImageView v{static_cast<uint8_t *>(img)};
        const bool ok_img = v.width() == 37 && v.height() == 21 && v.depth() == 10 && v.yuvFormat() == 1 &&
                            v.yuvRange() == 1 && v.matrixCoefficients() == 2 /* unspecified */;
        avifImageDestroy(img);
        uint8_t *enc = static_cast<uint8_t *>(avifEncoderCreate());
        const auto i32 = [&](size_t off) { return *reinterpret_cast<int32_t *>(enc + off); };
        const bool ok_enc = i32(kEncMaxThreads) == 1 && i32(kEncSpeed) == -1 && i32(kEncQuality) == -1 &&
                            i32(kEncQualityAlpha) == -1 && i32(kEncAutoTiling) == 0 && i32(44) == 63;
        avifEncoderDestroy(enc);
        if (!ok_img || !ok_enc)
            throw std::runtime_error(std::string("libavif ") + avifVersion() + ": struct layout differs from 1.4.x");
    }
    void *h_ = nullptr;
};

}  // namespace oavif_host
