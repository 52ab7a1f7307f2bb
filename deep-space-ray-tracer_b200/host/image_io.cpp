// image_io.cpp -- frame export on the host.
//
// PPM: byte-for-byte what /root/reference/src/gpu_render.cu:1100-1107 writes
// ("P6\n<W> <H>\n255\n" + W*H*3 raw bytes, top row first).
// PNG: replaces the `magick ppm png` subprocess of /root/reference/src/main.cpp:28-36 with an
// in-process encoder (8-bit RGB, filter 0 on every row, one zlib stream); ImageMagick is not
// installed on the GPU boxes.
#include "dsrt.h"

#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <vector>

extern "C" int dsrt_write_ppm(const char* path, const unsigned char* rgb, int width, int height) {
    if (!path || !rgb || width <= 0 || height <= 0) return DSRT_EINVAL;
    FILE* f = std::fopen(path, "wb");
    if (!f) return DSRT_EIO;
    std::fprintf(f, "P6\n%d %d\n255\n", width, height);
    const size_t bytes = (size_t)width * (size_t)height * 3;
    const size_t wrote = std::fwrite(rgb, 1, bytes, f);
    const int rc = std::fclose(f);
    return (wrote == bytes && rc == 0) ? DSRT_OK : DSRT_EIO;
}

namespace {

void put_be32(std::vector<unsigned char>& v, uint32_t x) {
    v.push_back((unsigned char)(x >> 24)); v.push_back((unsigned char)(x >> 16));
    v.push_back((unsigned char)(x >> 8));  v.push_back((unsigned char)x);
}

void put_chunk(std::vector<unsigned char>& png, const char type[4], const unsigned char* data, size_t n) {
    put_be32(png, (uint32_t)n);
    const size_t start = png.size();
    png.insert(png.end(), type, type + 4);
    if (n) png.insert(png.end(), data, data + n);
    const uint32_t crc = (uint32_t)crc32(0L, png.data() + start, (uInt)(png.size() - start));
    put_be32(png, crc);
}

}  // namespace

extern "C" int dsrt_write_png(const char* path, const unsigned char* rgb, int width, int height) {
    if (!path || !rgb || width <= 0 || height <= 0) return DSRT_EINVAL;
    const size_t row = (size_t)width * 3;
    std::vector<unsigned char> raw((row + 1) * (size_t)height);
    for (int y = 0; y < height; ++y) {
        raw[(row + 1) * y] = 0;   // filter type None
        std::memcpy(&raw[(row + 1) * y + 1], rgb + row * y, row);
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<unsigned char> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 3) != Z_OK) return DSRT_ENOMEM;

    std::vector<unsigned char> png;
    png.reserve(zlen + 128);
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    png.insert(png.end(), sig, sig + 8);
    std::vector<unsigned char> ihdr;
    put_be32(ihdr, (uint32_t)width);
    put_be32(ihdr, (uint32_t)height);
    ihdr.push_back(8);  // bit depth
    ihdr.push_back(2);  // colour type: truecolour
    ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    put_chunk(png, "IHDR", ihdr.data(), ihdr.size());
    put_chunk(png, "IDAT", z.data(), zlen);
    put_chunk(png, "IEND", nullptr, 0);

    FILE* f = std::fopen(path, "wb");
    if (!f) return DSRT_EIO;
    const size_t wrote = std::fwrite(png.data(), 1, png.size(), f);
    const int rc = std::fclose(f);
    return (wrote == png.size() && rc == 0) ? DSRT_OK : DSRT_EIO;
}
