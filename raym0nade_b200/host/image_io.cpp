// image_io.cpp — texture file decoders for the host's Model: PNG and DDS to 8-bit RGBA, top row first.
//
// The reference decodes PNG through libpng (src/material.cpp:220-286) and DDS through an embedded Python interpreter
// running imageio's FreeImage plugin (scripts/dds_to_array.py); neither library is part of this image or of the hot
// path, so the two formats are restated here from their published specifications:
//   PNG  (ISO/IEC 15948): zlib-compressed scanlines, filter types 0-4, colour types 0/2/3/4/6, bit depths 1-16,
//        tRNS; Adam7 interlacing is refused.
//   DDS  (DirectDraw Surface, legacy header): DXT1 / DXT3 / DXT5 block compression (S3TC: two RGB565 endpoints and
//        2-bit indices per 4x4 block; explicit 4-bit or interpolated 3-bit alpha) and uncompressed 24/32-bit RGB(A)
//        through the channel masks.  Only the top mip level is read: the mip chain is rebuilt by the library with the
//        reference's box filter (generateMipmaps, src/material.cpp:113-148), as the reference does after every load.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include <zlib.h>

#include "host.hpp"

namespace {

bool readFile(const std::string &path, std::vector<uint8_t> &out) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    in.seekg(0, std::ios::end);
    const std::streamoff n = in.tellg();
    if (n < 0) return false;
    in.seekg(0);
    out.resize(size_t(n));
    if (n) in.read(reinterpret_cast<char *>(out.data()), n);
    return bool(in);
}

uint32_t be32(const uint8_t *p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
uint32_t le32(const uint8_t *p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24); }

// ------------------------------------------------------------------------------------------------ PNG
int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

bool decodePng(const std::vector<uint8_t> &f, int &w, int &h, std::vector<uint8_t> &rgba, std::string &why) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    if (f.size() < 8 + 25 || std::memcmp(f.data(), sig, 8) != 0) { why = "not a PNG file"; return false; }
    size_t pos = 8;
    int depth = 0, ctype = 0, lace = 0;
    std::vector<uint8_t> idat, plte, trns;
    bool seenHeader = false, seenEnd = false;
    while (pos + 12 <= f.size() && !seenEnd) {
        const uint32_t n = be32(&f[pos]);
        if (n > f.size() - pos - 12) { why = "truncated chunk"; return false; }
        const uint8_t *type = &f[pos + 4], *data = &f[pos + 8];
        if (be32(data + n) != uint32_t(crc32(crc32(0L, type, 4), data, n))) { why = "chunk CRC mismatch"; return false; }
        if (!std::memcmp(type, "IHDR", 4)) {
            if (n != 13) { why = "bad IHDR"; return false; }
            w = int(be32(data)); h = int(be32(data + 4));
            depth = data[8]; ctype = data[9]; lace = data[12];
            if (data[10] != 0 || data[11] != 0) { why = "unknown compression/filter method"; return false; }
            seenHeader = true;
        } else if (!std::memcmp(type, "PLTE", 4)) plte.assign(data, data + n);
        else if (!std::memcmp(type, "tRNS", 4)) trns.assign(data, data + n);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + n);
        else if (!std::memcmp(type, "IEND", 4)) seenEnd = true;
        pos += 12 + size_t(n);
    }
    if (!seenHeader || idat.empty()) { why = "missing IHDR or IDAT"; return false; }
    if (w <= 0 || h <= 0 || int64_t(w) * h > (int64_t(1) << 28)) { why = "bad image size"; return false; }
    if (lace != 0) { why = "Adam7 interlaced PNGs are not supported"; return false; }
    const int samples = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    const bool depthOk = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
                         (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                         ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
    if (!samples || !depthOk) { why = "unsupported colour type / bit depth"; return false; }
    if (ctype == 3 && plte.size() < 3) { why = "palette image without PLTE"; return false; }
    const size_t bpp = std::max<size_t>(1, size_t(samples) * depth / 8);              // bytes per complete pixel, for the filters
    const size_t stride = (size_t(w) * samples * depth + 7) / 8;
    std::vector<uint8_t> raw((stride + 1) * size_t(h));
    uLongf got = uLongf(raw.size());
    if (uncompress(raw.data(), &got, idat.data(), uLong(idat.size())) != Z_OK || got != raw.size()) { why = "zlib stream does not match the image size"; return false; }
    // undo the scanline filters in place
    std::vector<uint8_t> zero(stride, 0);
    for (int y = 0; y < h; y++) {
        uint8_t *cur = &raw[(stride + 1) * y + 1];
        const uint8_t *up = y ? &raw[(stride + 1) * (y - 1) + 1] : zero.data();
        const int ft = cur[-1];
        if (ft > 4) { why = "unknown filter type"; return false; }
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0;
            const int pred = ft == 0 ? 0 : ft == 1 ? a : ft == 2 ? b : ft == 3 ? (a + b) / 2 : paeth(a, b, c);
            cur[i] = uint8_t(cur[i] + pred);
        }
    }
    rgba.resize(size_t(w) * h * 4);
    const int maxv = (1 << std::min(depth, 8)) - 1;
    for (int y = 0; y < h; y++) {
        const uint8_t *row = &raw[(stride + 1) * y + 1];
        for (int x = 0; x < w; x++) {
            // sample k of pixel x: 16-bit samples keep their high byte (png_set_strip_16), sub-byte samples are unpacked MSB first
            auto sample = [&](int k) -> int {
                const size_t idx = size_t(x) * samples + k;
                if (depth == 16) return row[idx * 2];
                if (depth == 8) return row[idx];
                const size_t bit = idx * depth;
                return (row[bit / 8] >> (8 - depth - bit % 8)) & maxv;
            };
            auto sample16 = [&](int k) -> int {              // full-precision value, for tRNS colour-key comparison
                const size_t idx = size_t(x) * samples + k;
                return depth == 16 ? (row[idx * 2] << 8 | row[idx * 2 + 1]) : sample(k);
            };
            uint8_t *o = &rgba[(size_t(y) * w + x) * 4];
            o[3] = 255;
            if (ctype == 3) {
                const int i = sample(0);
                if (size_t(i) * 3 + 2 >= plte.size()) { why = "palette index out of range"; return false; }
                o[0] = plte[i * 3]; o[1] = plte[i * 3 + 1]; o[2] = plte[i * 3 + 2];
                if (size_t(i) < trns.size()) o[3] = trns[i];
            } else if (ctype == 0 || ctype == 4) {
                const int g = depth < 8 ? sample(0) * 255 / maxv : sample(0);      // png_set_expand_gray_1_2_4_to_8
                o[0] = o[1] = o[2] = uint8_t(g);
                if (ctype == 4) o[3] = uint8_t(sample(1));
                else if (trns.size() >= 2 && sample16(0) == ((trns[0] << 8) | trns[1])) o[3] = 0;
            } else {
                o[0] = uint8_t(sample(0)); o[1] = uint8_t(sample(1)); o[2] = uint8_t(sample(2));
                if (ctype == 6) o[3] = uint8_t(sample(3));
                else if (trns.size() >= 6 && sample16(0) == ((trns[0] << 8) | trns[1]) && sample16(1) == ((trns[2] << 8) | trns[3]) &&
                         sample16(2) == ((trns[4] << 8) | trns[5])) o[3] = 0;
            }
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ DDS
void rgb565(uint16_t c, int out[3]) {
    const int r = (c >> 11) & 31, g = (c >> 5) & 63, b = c & 31;
    out[0] = (r << 3) | (r >> 2); out[1] = (g << 2) | (g >> 4); out[2] = (b << 3) | (b >> 2);
}

// colour part of a DXT block (8 bytes) into a 4x4 RGBA tile; `punch` = DXT1's 1-bit alpha mode when c0 <= c1
void dxtColour(const uint8_t *b, bool punch, uint8_t tile[16][4]) {
    const uint16_t c0 = uint16_t(b[0] | (b[1] << 8)), c1 = uint16_t(b[2] | (b[3] << 8));
    int p[4][4];
    rgb565(c0, p[0]); rgb565(c1, p[1]);
    p[0][3] = p[1][3] = p[2][3] = p[3][3] = 255;
    if (c0 > c1 || !punch) {
        for (int k = 0; k < 3; k++) { p[2][k] = (2 * p[0][k] + p[1][k]) / 3; p[3][k] = (p[0][k] + 2 * p[1][k]) / 3; }
    } else {
        for (int k = 0; k < 3; k++) { p[2][k] = (p[0][k] + p[1][k]) / 2; p[3][k] = 0; }
        p[3][3] = 0;
    }
    const uint32_t idx = le32(b + 4);
    for (int i = 0; i < 16; i++) {
        const int s = (idx >> (2 * i)) & 3;
        for (int k = 0; k < 4; k++) tile[i][k] = uint8_t(p[s][k]);
    }
}

void dxt5Alpha(const uint8_t *b, uint8_t tile[16][4]) {
    int a[8];
    a[0] = b[0]; a[1] = b[1];
    if (a[0] > a[1]) for (int k = 1; k < 7; k++) a[1 + k] = ((7 - k) * a[0] + k * a[1]) / 7;
    else { for (int k = 1; k < 5; k++) a[1 + k] = ((5 - k) * a[0] + k * a[1]) / 5; a[6] = 0; a[7] = 255; }
    uint64_t bits = 0;
    for (int k = 0; k < 6; k++) bits |= uint64_t(b[2 + k]) << (8 * k);
    for (int i = 0; i < 16; i++) tile[i][3] = uint8_t(a[(bits >> (3 * i)) & 7]);
}

int maskShift(uint32_t m) { int s = 0; while (m && !(m & 1)) { m >>= 1; s++; } return s; }
int maskBits(uint32_t m) { m >>= maskShift(m); int n = 0; while (m & 1) { m >>= 1; n++; } return n; }

bool decodeDds(const std::vector<uint8_t> &f, int &w, int &h, std::vector<uint8_t> &rgba, std::string &why) {
    if (f.size() < 128 || std::memcmp(f.data(), "DDS ", 4) != 0 || le32(&f[4]) != 124) { why = "not a DDS file"; return false; }
    h = int(le32(&f[12])); w = int(le32(&f[16]));
    if (w <= 0 || h <= 0 || int64_t(w) * h > (int64_t(1) << 28)) { why = "bad image size"; return false; }
    const uint32_t pfFlags = le32(&f[80]), bitCount = le32(&f[88]);
    const uint8_t *fourcc = &f[84];
    const uint8_t *data = f.data() + 128;
    const size_t avail = f.size() - 128;
    rgba.assign(size_t(w) * h * 4, 255);
    if (pfFlags & 0x4) {                                       // DDPF_FOURCC
        const bool d1 = !std::memcmp(fourcc, "DXT1", 4), d3 = !std::memcmp(fourcc, "DXT3", 4), d5 = !std::memcmp(fourcc, "DXT5", 4);
        if (!d1 && !d3 && !d5) { why = std::string("unsupported DDS format ") + std::string(reinterpret_cast<const char *>(fourcc), 4); return false; }
        const size_t bw = (size_t(w) + 3) / 4, bh = (size_t(h) + 3) / 4, bsz = d1 ? 8 : 16;
        if (bw * bh * bsz > avail) { why = "truncated DDS data"; return false; }
        for (size_t by = 0; by < bh; by++)
            for (size_t bx = 0; bx < bw; bx++) {
                const uint8_t *b = data + (by * bw + bx) * bsz;
                uint8_t tile[16][4];
                dxtColour(d1 ? b : b + 8, d1, tile);
                if (d3) for (int i = 0; i < 16; i++) { const int a4 = (b[i / 2] >> (4 * (i & 1))) & 15; tile[i][3] = uint8_t(a4 * 17); }
                if (d5) dxt5Alpha(b, tile);
                for (int i = 0; i < 16; i++) {
                    const size_t x = bx * 4 + (i & 3), y = by * 4 + (i >> 2);
                    if (x < size_t(w) && y < size_t(h)) std::memcpy(&rgba[(y * w + x) * 4], tile[i], 4);
                }
            }
        return true;
    }
    if ((pfFlags & 0x40) && (bitCount == 32 || bitCount == 24)) {          // DDPF_RGB through the channel masks
        const uint32_t masks[4] = {le32(&f[92]), le32(&f[96]), le32(&f[100]), (pfFlags & 0x1) ? le32(&f[104]) : 0u};
        const size_t px = bitCount / 8;
        if (size_t(w) * h * px > avail) { why = "truncated DDS data"; return false; }
        for (size_t i = 0; i < size_t(w) * h; i++) {
            uint32_t v = 0;
            for (size_t k = 0; k < px; k++) v |= uint32_t(data[i * px + k]) << (8 * k);
            for (int c = 0; c < 4; c++) {
                if (!masks[c]) { rgba[i * 4 + c] = c == 3 ? 255 : 0; continue; }
                const int bits = maskBits(masks[c]);
                const uint32_t s = (v & masks[c]) >> maskShift(masks[c]);
                rgba[i * 4 + c] = uint8_t(bits >= 8 ? s >> (bits - 8) : s * 255 / ((1u << bits) - 1));
            }
        }
        return true;
    }
    why = "unsupported DDS pixel format";
    return false;
}

}  // namespace

bool loadImageRGBA(const std::string &path, int &width, int &height, std::vector<uint8_t> &rgba, std::string &why) {
    std::vector<uint8_t> f;
    if (!readFile(path, f)) { why = "cannot open"; return false; }
    if (f.size() >= 4 && !std::memcmp(f.data(), "DDS ", 4)) return decodeDds(f, width, height, rgba, why);
    if (f.size() >= 4 && !std::memcmp(f.data() + 1, "PNG", 3)) return decodePng(f, width, height, rgba, why);
    why = "unsupported file format (PNG and DDS are read)";           // src/material.cpp:295
    return false;
}
