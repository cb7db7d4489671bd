#include "png_writer.hpp"

#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace rfk {

namespace {
void put_u32(std::vector<unsigned char>& v, std::uint32_t x) {
    v.push_back((x >> 24) & 0xff); v.push_back((x >> 16) & 0xff); v.push_back((x >> 8) & 0xff); v.push_back(x & 0xff);
}
void write_chunk(std::FILE* f, const char type[4], const unsigned char* data, std::size_t len) {
    std::vector<unsigned char> head;
    put_u32(head, (std::uint32_t)len);
    head.insert(head.end(), type, type + 4);
    uLong crc = crc32(0L, reinterpret_cast<const Bytef*>(type), 4);
    if (len) crc = crc32(crc, data, (uInt)len);
    std::vector<unsigned char> tail;
    put_u32(tail, (std::uint32_t)crc);
    if (std::fwrite(head.data(), 1, head.size(), f) != head.size() || (len && std::fwrite(data, 1, len, f) != len) ||
        std::fwrite(tail.data(), 1, tail.size(), f) != tail.size())
        throw std::runtime_error("png: short write");
}
}  // namespace

void write_png_rgba8(const std::string& path, const std::uint8_t* rgba, std::size_t width, std::size_t height, int compression_level) {
    if (!rgba || !width || !height || width > 0x7fffffff / 4 || height > 0x7fffffff) throw std::runtime_error("png: bad image");
    // filter type 0 (none) in front of every row
    const std::size_t stride = width * 4;
    std::vector<unsigned char> raw((stride + 1) * height);
    for (std::size_t y = 0; y < height; y++) {
        raw[y * (stride + 1)] = 0;
        std::copy(rgba + y * stride, rgba + (y + 1) * stride, raw.begin() + y * (stride + 1) + 1);
    }
    uLongf bound = compressBound((uLong)raw.size());
    std::vector<unsigned char> z(bound);
    if (compress2(z.data(), &bound, raw.data(), (uLong)raw.size(), compression_level) != Z_OK) throw std::runtime_error("png: deflate failed");

    std::FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("png: cannot open " + path);
    try {
        static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
        if (std::fwrite(sig, 1, 8, f) != 8) throw std::runtime_error("png: short write");
        std::vector<unsigned char> ihdr;
        put_u32(ihdr, (std::uint32_t)width);
        put_u32(ihdr, (std::uint32_t)height);
        ihdr.push_back(8);  // bit depth
        ihdr.push_back(6);  // colour type RGBA
        ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
        write_chunk(f, "IHDR", ihdr.data(), ihdr.size());
        write_chunk(f, "IDAT", z.data(), bound);
        write_chunk(f, "IEND", nullptr, 0);
    } catch (...) {
        std::fclose(f);
        throw;
    }
    if (std::fclose(f) != 0) throw std::runtime_error("png: close failed for " + path);
}

namespace {
void put_le32(std::vector<unsigned char>& v, std::uint32_t x) { for (int k = 0; k < 4; k++) v.push_back((x >> (8 * k)) & 0xff); }
void put_le64(std::vector<unsigned char>& v, std::uint64_t x) { for (int k = 0; k < 8; k++) v.push_back((x >> (8 * k)) & 0xff); }
void put_str(std::vector<unsigned char>& v, const char* s) { while (*s) v.push_back((unsigned char)*s++); v.push_back(0); }
void put_f32(std::vector<unsigned char>& v, float f) { std::uint32_t u; std::memcpy(&u, &f, 4); put_le32(v, u); }
void put_attr(std::vector<unsigned char>& v, const char* name, const char* type, const std::vector<unsigned char>& value) {
    put_str(v, name); put_str(v, type); put_le32(v, (std::uint32_t)value.size());
    v.insert(v.end(), value.begin(), value.end());
}
}  // namespace

void write_exr_rgba32f(const std::string& path, const float* rgba, std::size_t width, std::size_t height) {
    if (!rgba || !width || !height || width > 0x3fffffff / 16 || height > 0x3fffffff) throw std::runtime_error("exr: bad image");
    std::vector<unsigned char> head;
    put_le32(head, 20000630u);  // magic 0x76 0x2f 0x31 0x01
    put_le32(head, 2u);         // version 2, scanline, single part
    {
        std::vector<unsigned char> ch;  // channels in alphabetical order: A B G R, each FLOAT (2), linear 0, sampling 1 x 1
        for (const char* name : {"A", "B", "G", "R"}) {
            put_str(ch, name); put_le32(ch, 2u); ch.push_back(0); ch.push_back(0); ch.push_back(0); ch.push_back(0); put_le32(ch, 1u); put_le32(ch, 1u);
        }
        ch.push_back(0);
        put_attr(head, "channels", "chlist", ch);
    }
    put_attr(head, "compression", "compression", {0});  // none
    std::vector<unsigned char> window;
    put_le32(window, 0); put_le32(window, 0); put_le32(window, (std::uint32_t)(width - 1)); put_le32(window, (std::uint32_t)(height - 1));
    put_attr(head, "dataWindow", "box2i", window);
    put_attr(head, "displayWindow", "box2i", window);
    put_attr(head, "lineOrder", "lineOrder", {0});  // increasing y
    { std::vector<unsigned char> f; put_f32(f, 1.0f); put_attr(head, "pixelAspectRatio", "float", f); }
    { std::vector<unsigned char> f; put_f32(f, 0.0f); put_f32(f, 0.0f); put_attr(head, "screenWindowCenter", "v2f", f); }
    { std::vector<unsigned char> f; put_f32(f, 1.0f); put_attr(head, "screenWindowWidth", "float", f); }
    head.push_back(0);  // end of header
    const std::uint64_t row_bytes = 8 + 16 * (std::uint64_t)width, first = head.size() + 8 * (std::uint64_t)height;
    for (std::size_t y = 0; y < height; y++) put_le64(head, first + y * row_bytes);  // offset table
    std::FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("exr: cannot open " + path);
    bool ok = std::fwrite(head.data(), 1, head.size(), f) == head.size();
    std::vector<unsigned char> row;
    std::vector<float> plane(4 * width);
    for (std::size_t y = 0; y < height && ok; y++) {
        row.clear();
        put_le32(row, (std::uint32_t)y);
        put_le32(row, (std::uint32_t)(16 * width));
        const float* src = rgba + 4 * width * y;
        static const int order[4] = {3, 2, 1, 0};  // A B G R from R G B A
        for (int c = 0; c < 4; c++)
            for (std::size_t x = 0; x < width; x++) plane[c * width + x] = src[4 * x + order[c]];
        ok = std::fwrite(row.data(), 1, row.size(), f) == row.size() && std::fwrite(plane.data(), sizeof(float), plane.size(), f) == plane.size();  // little-endian host
    }
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) throw std::runtime_error("exr: short write to " + path);
}

}  // namespace rfk
