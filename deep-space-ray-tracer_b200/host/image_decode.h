// image_decode.h -- texture file decoding for the OBJ/MTL loader (host side, see image_decode.cpp).
#pragma once

#include <stdint.h>

#include <vector>

namespace dsrt_host {

struct DecodedImage {
    int w = 0, h = 0;
    std::vector<unsigned char> rgb;      // w*h*3, top row first
};

// Decodes PNG / binary PNM / BMP / TGA into 8-bit RGB.  false = unreadable or unsupported.
bool decode_image_file(const char* path, DecodedImage& out);

}  // namespace dsrt_host
