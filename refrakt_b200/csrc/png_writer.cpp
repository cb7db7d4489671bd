#include "png_writer.hpp"

#include <zlib.h>

#include <cstdio>
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

}  // namespace rfk
