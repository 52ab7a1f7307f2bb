// image_decode.h -- texture file decoding for the OBJ/MTL loader (host side, see image_decode.cpp).
#pragma once

#include <stdint.h>

#include <vector>

namespace dsrt_host {

struct DecodedImage {
    int w = 0, h = 0;
    std::vector<unsigned char> rgb;      // w*h*3, top row first
};

// jpeg_decode.cpp
bool decode_jpeg(const std::vector<unsigned char>& file, DecodedImage& out);

// Decodes PNG / JPEG / binary PNM / BMP / TGA into 8-bit RGB.  false = unreadable or unsupported; *exists (may be NULL)
// tells the two apart: the file could be read but is not an image this decoder understands.
bool decode_image_file(const char* path, DecodedImage& out, bool* exists = nullptr);

}  // namespace dsrt_host
