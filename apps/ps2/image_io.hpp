// 8-bit image files for the ps2 executable: PNG (read: gray / RGB / RGBA / gray+alpha, 8-bit,
// non-interlaced; write: gray) over zlib, and binary PGM/PPM.  Stands in for cv::imread(...,
// IMREAD_UNCHANGED) / cv::imwrite (common/include/common/BasicConfig.h:54-72, main.cpp:94-99); the
// build image has zlib but neither libpng nor OpenCV C++.
// Channel order follows OpenCV: a colour file is returned as B,G,R(,A) interleaved.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace ps2 {

struct Image8 {
    int rows = 0, cols = 0, channels = 0;     // interleaved, row-major, no padding
    std::vector<uint8_t> data;
    bool empty() const { return data.empty(); }
    uint8_t* row(int r) { return data.data() + size_t(r) * cols * channels; }
    const uint8_t* row(int r) const { return data.data() + size_t(r) * cols * channels; }
};

namespace detail {
inline std::vector<uint8_t> read_file(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("could not open " + path);
    std::vector<uint8_t> buf;
    uint8_t tmp[65536];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    std::fclose(f);
    return buf;
}
inline uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
inline void put_be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
} // namespace detail

inline Image8 read_png(const std::vector<uint8_t>& file, const std::string& name) {
    using namespace detail;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) throw std::runtime_error(name + ": not a PNG file");
    uint32_t w = 0, h = 0; int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat;
    size_t pos = 8;
    while (pos + 12 <= file.size()) {
        const uint32_t len = be32(&file[pos]);
        const char* type = reinterpret_cast<const char*>(&file[pos + 4]);
        if (pos + 12 + len > file.size()) throw std::runtime_error(name + ": truncated PNG chunk");
        const uint8_t* body = &file[pos + 8];
        if (!std::memcmp(type, "IHDR", 4)) {
            w = be32(body); h = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + len;
    }
    if (depth != 8 || interlace != 0) throw std::runtime_error(name + ": only 8-bit non-interlaced PNGs are supported");
    int ch;
    switch (ctype) { case 0: ch = 1; break; case 2: ch = 3; break; case 4: ch = 2; break; case 6: ch = 4; break;
    default: throw std::runtime_error(name + ": unsupported PNG colour type " + std::to_string(ctype)); }
    const size_t stride = size_t(w) * ch;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf raw_len = raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), idat.size()) != Z_OK || raw_len != raw.size())
        throw std::runtime_error(name + ": PNG inflate failed");
    std::vector<uint8_t> px(stride * h);
    for (uint32_t y = 0; y < h; ++y) {                       // undo the per-row filters
        const uint8_t f = raw[y * (stride + 1)];
        const uint8_t* in = &raw[y * (stride + 1) + 1];
        uint8_t* out = &px[y * stride];
        const uint8_t* up = y ? &px[(y - 1) * stride] : nullptr;
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= size_t(ch) ? out[i - ch] : 0, b = up ? up[i] : 0, c = (up && i >= size_t(ch)) ? up[i - ch] : 0;
            int v = in[i];
            switch (f) { case 0: break; case 1: v += a; break; case 2: v += b; break; case 3: v += (a + b) >> 1; break;
            case 4: v += paeth(a, b, c); break; default: throw std::runtime_error(name + ": bad PNG filter"); }
            out[i] = uint8_t(v);
        }
    }
    Image8 img;
    img.rows = int(h); img.cols = int(w);
    // OpenCV's IMREAD_UNCHANGED: gray stays 1 channel, gray+alpha becomes BGRA, RGB -> BGR, RGBA -> BGRA
    img.channels = ch == 1 ? 1 : (ch == 3 ? 3 : 4);
    img.data.resize(size_t(h) * w * img.channels);
    for (size_t i = 0; i < size_t(w) * h; ++i) {
        const uint8_t* s = &px[i * ch];
        uint8_t* d = &img.data[i * img.channels];
        if (ch == 1) d[0] = s[0];
        else if (ch == 2) { d[0] = d[1] = d[2] = s[0]; d[3] = s[1]; }
        else { d[0] = s[2]; d[1] = s[1]; d[2] = s[0]; if (ch == 4) d[3] = s[3]; }
    }
    return img;
}

inline Image8 read_pnm(const std::vector<uint8_t>& file, const std::string& name) {
    size_t pos = 0;
    auto token = [&]() {
        std::string t;
        while (pos < file.size()) {
            const char c = char(file[pos]);
            if (c == '#') { while (pos < file.size() && file[pos] != '\n') ++pos; continue; }
            if (std::isspace(static_cast<unsigned char>(c))) { ++pos; if (!t.empty()) break; continue; }
            t.push_back(c); ++pos;
        }
        return t;
    };
    const std::string magic = token();
    if (magic != "P5" && magic != "P6") throw std::runtime_error(name + ": only binary PGM (P5) / PPM (P6) are supported");
    const int w = std::stoi(token()), h = std::stoi(token()), maxv = std::stoi(token());
    if (maxv != 255) throw std::runtime_error(name + ": only maxval 255 is supported");
    const int ch = magic == "P5" ? 1 : 3;
    if (pos + size_t(w) * h * ch > file.size()) throw std::runtime_error(name + ": truncated PNM");
    Image8 img;
    img.rows = h; img.cols = w; img.channels = ch;
    img.data.assign(file.begin() + pos, file.begin() + pos + size_t(w) * h * ch);
    if (ch == 3) for (size_t i = 0; i < size_t(w) * h; ++i) std::swap(img.data[3 * i], img.data[3 * i + 2]);   // RGB -> BGR
    return img;
}

// cv::imread(path, IMREAD_UNCHANGED); throws on failure (the reference logs and fails the config load).
inline Image8 imread(const std::string& path) {
    const std::vector<uint8_t> file = detail::read_file(path);
    if (file.size() >= 2 && file[0] == 'P' && (file[1] == '5' || file[1] == '6')) return read_pnm(file, path);
    return read_png(file, path);
}

// cv::imwrite for a single-channel 8-bit image: .png (deflate) or .pgm by extension.
inline void imwrite_gray(const std::string& path, const uint8_t* data, int rows, int cols, size_t step) {
    using namespace detail;
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("could not write " + path);
    const bool pgm = path.size() >= 4 && path.compare(path.size() - 4, 4, ".pgm") == 0;
    if (pgm) {
        std::fprintf(f, "P5\n%d %d\n255\n", cols, rows);
        for (int r = 0; r < rows; ++r) std::fwrite(data + size_t(r) * step, 1, cols, f);
        std::fclose(f);
        return;
    }
    std::vector<uint8_t> raw(size_t(rows) * (cols + 1));
    for (int r = 0; r < rows; ++r) { raw[size_t(r) * (cols + 1)] = 0; std::memcpy(&raw[size_t(r) * (cols + 1) + 1], data + size_t(r) * step, cols); }
    uLongf zlen = compressBound(raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), raw.size(), 6) != Z_OK) { std::fclose(f); throw std::runtime_error("deflate failed for " + path); }
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    auto chunk = [&](const char* type, const uint8_t* body, uint32_t len) {
        put_be32(out, len);
        const size_t start = out.size();
        out.insert(out.end(), type, type + 4);
        if (len) out.insert(out.end(), body, body + len);
        put_be32(out, uint32_t(crc32(0L, &out[start], uInt(len + 4))));
    };
    uint8_t ihdr[13];
    ihdr[0] = cols >> 24; ihdr[1] = cols >> 16; ihdr[2] = cols >> 8; ihdr[3] = cols;
    ihdr[4] = rows >> 24; ihdr[5] = rows >> 16; ihdr[6] = rows >> 8; ihdr[7] = rows;
    ihdr[8] = 8; ihdr[9] = 0; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", z.data(), uint32_t(zlen));
    chunk("IEND", nullptr, 0);
    std::fwrite(out.data(), 1, out.size(), f);
    std::fclose(f);
}

} // namespace ps2
