// png_io.cpp — host mirror of png_image.zig:19-148 without libpng (its headers are not in this image):
// a small PNG codec on top of zlib.  Supports what the reference supports: 8-bit RGB / RGBA,
// non-interlaced.  Ancillary chunks (gAMA, cHRM, ...) are ignored, as the reference never calls
// png_set_gamma.
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../../include/zrt_host.h"

namespace {

uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void put32(std::vector<uint8_t> &v, uint32_t x) {
    v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x);
}
int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
bool readAll(const char *path, std::vector<uint8_t> *out) {
    FILE *f = std::fopen(path, "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out->resize(n > 0 ? (size_t)n : 0);
    const bool ok = n >= 0 && std::fread(out->data(), 1, out->size(), f) == out->size();
    std::fclose(f);
    return ok;
}
void writeChunk(std::vector<uint8_t> &png, const char type[4], const std::vector<uint8_t> &data) {
    put32(png, (uint32_t)data.size());
    const size_t start = png.size();
    png.insert(png.end(), type, type + 4);
    png.insert(png.end(), data.begin(), data.end());
    put32(png, (uint32_t)crc32(0, png.data() + start, (uInt)(png.size() - start)));
}

} // namespace

static int pngRead(const char *path, uint8_t **pixels, uint32_t *width, uint32_t *height, uint32_t *channels);

extern "C" int zrt_host_png_read(const char *path, uint8_t **pixels, uint32_t *width, uint32_t *height, uint32_t *channels) {
    if (!path || !pixels || !width || !height || !channels) return ZRT_ERR_INVALID;
    try { // nothing throws across the C ABI (an IHDR can ask for more memory than there is)
        return pngRead(path, pixels, width, height, channels);
    } catch (const std::bad_alloc &) {
        return ZRT_ERR_OOM;
    } catch (...) {
        return ZRT_ERR_IO;
    }
}

static int pngRead(const char *path, uint8_t **pixels, uint32_t *width, uint32_t *height, uint32_t *channels) {
    std::vector<uint8_t> file;
    if (!readAll(path, &file)) return ZRT_ERR_IO;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) return ZRT_ERR_IO; // PngError.BadPngFile
    uint32_t w = 0, h = 0, ch = 0;
    std::vector<uint8_t> idat;
    size_t pos = 8;
    while (pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        const uint8_t *type = &file[pos + 4];
        const uint8_t *data = &file[pos + 8];
        if (pos + 12 + len > file.size()) return ZRT_ERR_IO;
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len < 13) return ZRT_ERR_IO;
            w = be32(data);
            h = be32(data + 4);
            const int bit_depth = data[8], color_type = data[9], interlace = data[12];
            // png_image.zig:45-52: only RGB (2) / RGBA (6), 8 bits
            if (bit_depth != 8 || (color_type != 2 && color_type != 6) || interlace != 0) return ZRT_ERR_INVALID;
            ch = color_type == 2 ? 3 : 4;
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + len;
    }
    if (w == 0 || h == 0 || ch == 0) return ZRT_ERR_IO;
    if ((uint64_t)w * h > (1ull << 28)) return ZRT_ERR_INVALID; // an untrusted header does not get to size the allocation
    const size_t stride = (size_t)w * ch;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) return ZRT_ERR_IO;
    uint8_t *out = (uint8_t *)std::malloc(stride * h);
    if (!out) return ZRT_ERR_OOM;
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t *line = &raw[(stride + 1) * y];
        const int filter = line[0];
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= ch ? cur[i - ch] : 0, b = prev[i], c = i >= ch ? prev[i - ch] : 0;
            int v = line[1 + i];
            switch (filter) {
            case 0: break;
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) >> 1; break;
            case 4: v += paeth(a, b, c); break;
            default: std::free(out); return ZRT_ERR_IO;
            }
            cur[i] = (uint8_t)v;
        }
        std::memcpy(out + stride * (h - y - 1), cur.data(), stride); // png_image.zig:86 row flip
        prev.swap(cur);
    }
    *pixels = out; *width = w; *height = h; *channels = ch;
    return ZRT_OK;
}

static int writePngRaw(const char *path, const std::vector<uint8_t> &raw, uint32_t width, uint32_t height);

extern "C" int zrt_host_png_write(const char *path, const float *rgb, uint32_t width, uint32_t height) {
    if (!path || !rgb || width == 0 || height == 0) return ZRT_ERR_INVALID;
    try {
    const size_t stride = (size_t)width * 3;
    std::vector<uint8_t> raw((stride + 1) * height);
    for (uint32_t y = 0; y < height; y++) {
        uint8_t *line = &raw[(stride + 1) * y];
        line[0] = 0; // filter none
        const float *src = rgb + (size_t)(height - y - 1) * stride; // png_image.zig:133
        for (size_t i = 0; i < stride; i++) {
            float v = 255.999f * src[i]; // png_image.zig:136-140
            v = v < 255.0f ? v : 255.0f;
            v = 0.0f > v ? 0.0f : v;
            line[1 + i] = (uint8_t)v;
        }
    }
    return writePngRaw(path, raw, width, height);
    } catch (const std::bad_alloc &) {
        return ZRT_ERR_OOM;
    }
}

extern "C" int zrt_host_png_write_rgb8(const char *path, const uint8_t *rgb8, uint32_t width, uint32_t height) {
    if (!path || !rgb8 || width == 0 || height == 0) return ZRT_ERR_INVALID;
    try {
    const size_t stride = (size_t)width * 3;
    std::vector<uint8_t> raw((stride + 1) * height);
    for (uint32_t y = 0; y < height; y++) {
        raw[(stride + 1) * y] = 0; // filter none
        std::memcpy(&raw[(stride + 1) * y + 1], rgb8 + stride * y, stride);
    }
    return writePngRaw(path, raw, width, height);
    } catch (const std::bad_alloc &) {
        return ZRT_ERR_OOM;
    }
}

static int writePngRaw(const char *path, const std::vector<uint8_t> &raw, uint32_t width, uint32_t height) {
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return ZRT_ERR_IO;
    z.resize(zlen);
    std::vector<uint8_t> png = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    put32(ihdr, width);
    put32(ihdr, height);
    ihdr.insert(ihdr.end(), {8, 2, 0, 0, 0}); // 8-bit RGB, png_image.zig:116-126
    writeChunk(png, "IHDR", ihdr);
    writeChunk(png, "IDAT", z);
    writeChunk(png, "IEND", {});
    FILE *f = std::fopen(path, "wb");
    if (!f) return ZRT_ERR_IO; // PngError.FailedToOpenFile
    const bool ok = std::fwrite(png.data(), 1, png.size(), f) == png.size();
    std::fclose(f);
    return ok ? ZRT_OK : ZRT_ERR_IO;
}

extern "C" void zrt_host_free(void *p) { std::free(p); }
