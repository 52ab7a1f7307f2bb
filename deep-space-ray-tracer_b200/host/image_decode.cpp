// image_decode.cpp -- texture file decoding for the OBJ/MTL loader (host side).
//
// The reference decodes `map_Kd` images with the vendored stb_image
// (/root/reference/src/gpu_scene_builder.cpp:214: stbi_load(path, &w, &h, &n, 3), forced RGB).
// This file is an independent decoder with the same observable result for the formats it
// covers -- 8-bit RGB, top row first:
//   PNG   all colour types and bit depths, Adam7; alpha dropped, 16-bit -> high byte,
//         1/2/4-bit grey scaled to 0..255 (what stb's conversion to 3 x 8-bit yields)
//   PNM   binary P5 / P6, maxval <= 65535 (16-bit -> high byte)
//   BMP   uncompressed 8-bit paletted / 24-bit / 32-bit, bottom-up or top-down
//   TGA   true-colour and grey, raw or RLE (types 2, 3, 10, 11), 8/24/32 bits
//   JPEG  baseline and progressive Huffman DCT, grey or YCbCr/RGB (jpeg_decode.cpp: stb_image's
//         integer arithmetic restated, so the bytes are the reference's)
// Anything else (GIF, PSD, HDR, CMYK or arithmetic-coded JPEG ...) is reported as undecodable; the
// loader then takes the reference's own failure path, a 1x1 white texture
// (gpu_scene_builder.cpp:215-221), and counts the file so that the CLI can refuse to render.
#include "image_decode.h"

#include <zlib.h>

#include <cstdio>
#include <cstring>

namespace dsrt_host {

namespace {

bool read_file(const char* path, std::vector<unsigned char>& out) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (n < 0) { std::fclose(f); return false; }
    out.resize((size_t)n);
    const size_t got = n ? std::fread(out.data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    return got == (size_t)n;
}

inline uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline uint32_t le32(const unsigned char* p) { return ((uint32_t)p[3] << 24) | ((uint32_t)p[2] << 16) | ((uint32_t)p[1] << 8) | p[0]; }
inline uint32_t le16(const unsigned char* p) { return ((uint32_t)p[1] << 8) | p[0]; }

// ---- PNG -------------------------------------------------------------------------------------
inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
    if (pa <= pb && pa <= pc) return a;
    return pb <= pc ? b : c;
}

// Undoes the scanline filters of one (sub)image in place; returns false on a bad filter byte.
bool png_unfilter(unsigned char* data, size_t rows, size_t row_bytes, size_t bpp) {
    for (size_t y = 0; y < rows; ++y) {
        unsigned char* cur = data + y * (row_bytes + 1);
        const unsigned char* prev = y ? cur - (row_bytes + 1) + 1 : nullptr;
        const int ft = cur[0];
        unsigned char* r = cur + 1;
        for (size_t i = 0; i < row_bytes; ++i) {
            const int a = i >= bpp ? r[i - bpp] : 0;
            const int b = prev ? prev[i] : 0;
            const int c = (prev && i >= bpp) ? prev[i - bpp] : 0;
            int add;
            switch (ft) {
                case 0: add = 0; break;
                case 1: add = a; break;
                case 2: add = b; break;
                case 3: add = (a + b) >> 1; break;
                case 4: add = paeth(a, b, c); break;
                default: return false;
            }
            r[i] = (unsigned char)(r[i] + add);
        }
    }
    return true;
}

bool decode_png(const std::vector<unsigned char>& f, DecodedImage& out) {
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (f.size() < 8 + 25 || std::memcmp(f.data(), sig, 8) != 0) return false;
    uint32_t W = 0, H = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<unsigned char> idat, plte;
    size_t pos = 8;
    bool have_ihdr = false, done = false;
    while (!done && pos + 12 <= f.size()) {
        const uint32_t len = be32(&f[pos]);
        const unsigned char* type = &f[pos + 4];
        if (pos + 12 + (size_t)len > f.size()) return false;
        const unsigned char* data = &f[pos + 8];
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len < 13) return false;
            W = be32(data); H = be32(data + 4);
            depth = data[8]; ctype = data[9]; interlace = data[12];
            if (data[10] != 0 || data[11] != 0 || interlace > 1) return false;
            have_ihdr = true;
        } else if (!std::memcmp(type, "PLTE", 4)) {
            plte.assign(data, data + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            done = true;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || W == 0 || H == 0 || W > (1u << 24) || H > (1u << 24)) return false;
    int channels;
    switch (ctype) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: return false;
    }
    if (!(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) return false;
    if ((ctype == 2 || ctype == 4 || ctype == 6) && depth < 8) return false;
    if (ctype == 3 && (depth == 16 || plte.size() < 3)) return false;
    const size_t bits_pp = (size_t)channels * depth;
    const size_t bpp = bits_pp >= 8 ? bits_pp / 8 : 1;

    // passes: {x0, y0, dx, dy}
    static const int adam7[7][4] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    const int npass = interlace ? 7 : 1;
    size_t total = 0;
    size_t pw[7], ph[7], prow[7];
    for (int p = 0; p < npass; ++p) {
        if (interlace) {
            pw[p] = (W > (uint32_t)adam7[p][0]) ? (W - adam7[p][0] + adam7[p][2] - 1) / adam7[p][2] : 0;
            ph[p] = (H > (uint32_t)adam7[p][1]) ? (H - adam7[p][1] + adam7[p][3] - 1) / adam7[p][3] : 0;
        } else { pw[p] = W; ph[p] = H; }
        prow[p] = (pw[p] * bits_pp + 7) / 8;
        if (pw[p] && ph[p]) total += (prow[p] + 1) * ph[p];
    }
    // a deflate stream expands at most ~1032:1, so a header that promises more than the IDAT data can hold is
    // malformed: refuse before allocating (W and H come from the file and may be anything up to 2^24 each)
    if (total > idat.size() * 1040u + 4096u || (uint64_t)W * H > (1ull << 28)) return false;
    std::vector<unsigned char> raw(total);
    {
        uLongf dlen = (uLongf)raw.size();
        const int rc = uncompress(raw.data(), &dlen, idat.data(), (uLong)idat.size());
        if (rc != Z_OK || dlen != raw.size()) return false;
    }
    out.w = (int)W; out.h = (int)H;
    out.rgb.assign((size_t)W * H * 3, 0);
    const int gray_scale = depth == 1 ? 255 : depth == 2 ? 85 : depth == 4 ? 17 : 1;
    size_t off = 0;
    for (int p = 0; p < npass; ++p) {
        if (!pw[p] || !ph[p]) continue;
        unsigned char* sub = raw.data() + off;
        if (!png_unfilter(sub, ph[p], prow[p], bpp)) return false;
        for (size_t y = 0; y < ph[p]; ++y) {
            const unsigned char* row = sub + y * (prow[p] + 1) + 1;
            const size_t oy = interlace ? (size_t)adam7[p][1] + y * adam7[p][3] : y;
            for (size_t x = 0; x < pw[p]; ++x) {
                const size_t ox = interlace ? (size_t)adam7[p][0] + x * adam7[p][2] : x;
                unsigned char* o = &out.rgb[(oy * W + ox) * 3];
                auto sample = [&](size_t idx) -> unsigned {      // idx-th sample of this row, as 8 bits
                    if (depth == 8) return row[idx];
                    if (depth == 16) return row[idx * 2];        // high byte (stb: v >> 8)
                    const size_t bit = idx * depth;
                    return (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
                };
                if (ctype == 3) {
                    unsigned k = sample(x);
                    if ((size_t)k * 3 + 2 >= plte.size()) k = 0;
                    o[0] = plte[k * 3]; o[1] = plte[k * 3 + 1]; o[2] = plte[k * 3 + 2];
                } else if (channels <= 2) {
                    unsigned g = sample(x * channels);
                    if (depth < 8) g *= gray_scale;
                    o[0] = o[1] = o[2] = (unsigned char)g;
                } else {
                    o[0] = (unsigned char)sample(x * channels);
                    o[1] = (unsigned char)sample(x * channels + 1);
                    o[2] = (unsigned char)sample(x * channels + 2);
                }
            }
        }
        off += (prow[p] + 1) * ph[p];
    }
    return true;
}

// ---- PNM (binary P5 / P6) ----------------------------------------------------------------------
bool pnm_int(const std::vector<unsigned char>& f, size_t& pos, long& v) {
    for (;;) {
        while (pos < f.size() && (f[pos] == ' ' || f[pos] == '\t' || f[pos] == '\n' || f[pos] == '\r' || f[pos] == '\f' || f[pos] == '\v')) ++pos;
        if (pos < f.size() && f[pos] == '#') { while (pos < f.size() && f[pos] != '\n' && f[pos] != '\r') ++pos; continue; }
        break;
    }
    if (pos >= f.size() || f[pos] < '0' || f[pos] > '9') return false;
    v = 0;
    while (pos < f.size() && f[pos] >= '0' && f[pos] <= '9') { v = v * 10 + (f[pos] - '0'); if (v > (1L << 30)) return false; ++pos; }
    return true;
}

bool decode_pnm(const std::vector<unsigned char>& f, DecodedImage& out) {
    if (f.size() < 7 || f[0] != 'P' || (f[1] != '5' && f[1] != '6')) return false;
    const int ch = f[1] == '6' ? 3 : 1;
    size_t pos = 2;
    long W, H, maxv;
    if (!pnm_int(f, pos, W) || !pnm_int(f, pos, H) || !pnm_int(f, pos, maxv)) return false;
    if (W <= 0 || H <= 0 || maxv <= 0 || maxv > 65535) return false;
    ++pos;                                                    // the single whitespace after maxval
    const size_t bps = maxv > 255 ? 2 : 1;
    if (pos + (size_t)W * H * ch * bps > f.size()) return false;
    out.w = (int)W; out.h = (int)H;
    out.rgb.resize((size_t)W * H * 3);
    const unsigned char* p = &f[pos];
    for (size_t i = 0; i < (size_t)W * H; ++i)
        for (int c = 0; c < 3; ++c)
            out.rgb[i * 3 + c] = p[(i * ch + (ch == 3 ? c : 0)) * bps];     // 16-bit: big-endian high byte
    return true;
}

// ---- BMP ---------------------------------------------------------------------------------------
bool decode_bmp(const std::vector<unsigned char>& f, DecodedImage& out) {
    if (f.size() < 54 || f[0] != 'B' || f[1] != 'M') return false;
    const uint32_t data_off = le32(&f[10]), hdr = le32(&f[14]);
    if (hdr < 40) return false;
    const int32_t W = (int32_t)le32(&f[18]);
    int32_t H = (int32_t)le32(&f[22]);
    const uint32_t bpp = le16(&f[28]), comp = le32(&f[30]);
    if (W <= 0 || H == 0 || (comp != 0 && !(comp == 3 && bpp == 32))) return false;
    const bool top_down = H < 0;
    if (top_down) H = -H;
    if (bpp != 8 && bpp != 24 && bpp != 32) return false;
    const size_t row = (((size_t)W * bpp + 31) / 32) * 4;
    if ((size_t)data_off + row * (size_t)H > f.size()) return false;
    const unsigned char* pal = &f[14 + hdr];
    uint32_t ncol = le32(&f[46]);
    if (bpp == 8) { if (ncol == 0) ncol = 256; if (14 + (size_t)hdr + ncol * 4 > f.size()) return false; }
    out.w = W; out.h = H;
    out.rgb.resize((size_t)W * H * 3);
    for (int y = 0; y < H; ++y) {
        const unsigned char* src = &f[data_off + row * (size_t)(top_down ? y : H - 1 - y)];
        unsigned char* dst = &out.rgb[(size_t)y * W * 3];
        for (int x = 0; x < W; ++x) {
            if (bpp == 8) {
                uint32_t k = src[x]; if (k >= ncol) k = 0;
                dst[x * 3] = pal[k * 4 + 2]; dst[x * 3 + 1] = pal[k * 4 + 1]; dst[x * 3 + 2] = pal[k * 4];
            } else {
                const unsigned char* q = src + (size_t)x * (bpp / 8);
                dst[x * 3] = q[2]; dst[x * 3 + 1] = q[1]; dst[x * 3 + 2] = q[0];
            }
        }
    }
    return true;
}

// ---- TGA ---------------------------------------------------------------------------------------
bool decode_tga(const std::vector<unsigned char>& f, DecodedImage& out) {
    if (f.size() < 18) return false;
    const int idlen = f[0], cmap = f[1], type = f[2];
    const int W = (int)le16(&f[12]), H = (int)le16(&f[14]), bpp = f[16], desc = f[17];
    if (cmap != 0 || !(type == 2 || type == 3 || type == 10 || type == 11)) return false;
    const bool gray = (type == 3 || type == 11), rle = type >= 10;
    if (W <= 0 || H <= 0 || (gray ? bpp != 8 && bpp != 16 : bpp != 24 && bpp != 32)) return false;
    const size_t px = (size_t)bpp / 8;
    size_t pos = 18 + (size_t)idlen;
    // the header may promise up to 65535 x 65535 pixels: check against what the file can hold before allocating
    // (a run-length packet expands at most 128:1)
    if (pos > f.size() || (size_t)W * H > (f.size() - pos) * (rle ? 128u : 1u)) return false;
    std::vector<unsigned char> pix((size_t)W * H * px);
    if (!rle) {
        if (pos + pix.size() > f.size()) return false;
        std::memcpy(pix.data(), &f[pos], pix.size());
    } else {
        size_t n = 0;
        while (n < (size_t)W * H) {
            if (pos >= f.size()) return false;
            const int c = f[pos++];
            const size_t cnt = (size_t)(c & 127) + 1;
            if (n + cnt > (size_t)W * H) return false;
            if (c & 128) {
                if (pos + px > f.size()) return false;
                for (size_t k = 0; k < cnt; ++k) std::memcpy(&pix[(n + k) * px], &f[pos], px);
                pos += px;
            } else {
                if (pos + cnt * px > f.size()) return false;
                std::memcpy(&pix[n * px], &f[pos], cnt * px);
                pos += cnt * px;
            }
            n += cnt;
        }
    }
    const bool top_down = (desc & 0x20) != 0, right_left = (desc & 0x10) != 0;
    out.w = W; out.h = H;
    out.rgb.resize((size_t)W * H * 3);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const unsigned char* q = &pix[((size_t)(top_down ? y : H - 1 - y) * W + (right_left ? W - 1 - x : x)) * px];
            unsigned char* o = &out.rgb[((size_t)y * W + x) * 3];
            if (gray) { o[0] = o[1] = o[2] = q[0]; }
            else { o[0] = q[2]; o[1] = q[1]; o[2] = q[0]; }
        }
    return true;
}

}  // namespace

bool decode_image_file(const char* path, DecodedImage& out, bool* exists) {
    std::vector<unsigned char> f;
    if (exists) *exists = false;
    if (!read_file(path, f)) return false;
    if (exists) *exists = true;
    if (f.size() < 4) return false;
    if (f[0] == 0x89 && f[1] == 'P') return decode_png(f, out);
    if (f[0] == 0xFF && f[1] == 0xD8) return decode_jpeg(f, out);
    if (f[0] == 'P' && (f[1] == '5' || f[1] == '6')) return decode_pnm(f, out);
    if (f[0] == 'B' && f[1] == 'M') return decode_bmp(f, out);
    const size_t L = std::strlen(path);
    // TGA has no magic number: it is tried last, on anything (its header checks reject what is not one)
    (void)L;
    return decode_tga(f, out);
}

}  // namespace dsrt_host
