// image_out.cpp — the output stage after write_color: the reference prints a P3 text image on
// stdout (header src/main.rs:472, one "r g b" line per pixel src/math.rs:127-131, rows top to bottom :591-596).
// rtw_write_ppm writes exactly those bytes to a file; rtw_write_png writes the same pixels as an 8-bit RGB PNG
// (stored deflate blocks: no compression library needed, every PNG reader accepts it).  Host only, no CUDA.
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../include/rtw.h"

namespace {

uint32_t crc_table[256];
bool crc_ready = false;
uint32_t crc32_update(uint32_t c, const uint8_t* p, size_t n) {
    if (!crc_ready) {
        for (uint32_t i = 0; i < 256; ++i) { uint32_t k = i; for (int j = 0; j < 8; ++j) k = (k & 1) ? 0xEDB88320u ^ (k >> 1) : k >> 1; crc_table[i] = k; }
        crc_ready = true;
    }
    for (size_t i = 0; i < n; ++i) c = crc_table[(c ^ p[i]) & 0xff] ^ (c >> 8);
    return c;
}
void be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
void chunk(std::vector<uint8_t>& out, const char type[4], const std::vector<uint8_t>& data) {
    be32(out, (uint32_t)data.size());
    const size_t at = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data.begin(), data.end());
    be32(out, crc32_update(0xffffffffu, out.data() + at, out.size() - at) ^ 0xffffffffu);
}

}  // namespace

extern "C" int rtw_write_ppm(const char* path, const uint8_t* rgb8, int32_t width, int32_t height) {
    if (!path || !rgb8 || width < 1 || height < 1) return RTW_ERR_INVALID_ARG;
    FILE* f = std::fopen(path, "wb");
    if (!f) return RTW_ERR_INVALID_ARG;
    std::fprintf(f, "P3\n%d %d\n255\n\n", width, height);                       // println!("P3\n{} {}\n255\n") :472
    const size_t n = (size_t)width * height;
    for (size_t i = 0; i < n; ++i) std::fprintf(f, "%d %d %d\n", rgb8[3 * i], rgb8[3 * i + 1], rgb8[3 * i + 2]);
    return std::fclose(f) == 0 ? RTW_OK : RTW_ERR_INVALID_ARG;
}

extern "C" int rtw_write_png(const char* path, const uint8_t* rgb8, int32_t width, int32_t height) {
    if (!path || !rgb8 || width < 1 || height < 1) return RTW_ERR_INVALID_ARG;
    std::vector<uint8_t> raw;                                                   // scanlines, filter byte 0 each
    raw.reserve(((size_t)width * 3 + 1) * height);
    for (int y = 0; y < height; ++y) { raw.push_back(0); raw.insert(raw.end(), rgb8 + (size_t)y * width * 3, rgb8 + (size_t)(y + 1) * width * 3); }
    std::vector<uint8_t> z = {0x78, 0x01};                                      // zlib header, then stored blocks of <= 65535 bytes
    uint32_t a = 1, b = 0;
    for (size_t off = 0; off < raw.size() || off == 0; ) {
        const size_t len = raw.size() - off < 65535 ? raw.size() - off : 65535;
        z.push_back(off + len >= raw.size() ? 1 : 0);
        z.push_back(len & 0xff); z.push_back(len >> 8); z.push_back(~len & 0xff); z.push_back((~len >> 8) & 0xff);
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + len);
        for (size_t i = 0; i < len; ++i) { a = (a + raw[off + i]) % 65521u; b = (b + a) % 65521u; }   // Adler-32
        off += len;
        if (len == 0) break;
    }
    be32(z, (b << 16) | a);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    be32(ihdr, (uint32_t)width); be32(ihdr, (uint32_t)height);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 8-bit, colour type 2 (RGB)
    chunk(out, "IHDR", ihdr);
    chunk(out, "IDAT", z);
    chunk(out, "IEND", {});
    FILE* f = std::fopen(path, "wb");
    if (!f) return RTW_ERR_INVALID_ARG;
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    return (std::fclose(f) == 0 && ok) ? RTW_OK : RTW_ERR_INVALID_ARG;
}
