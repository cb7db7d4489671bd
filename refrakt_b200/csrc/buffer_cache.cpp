#include "buffer_cache.hpp"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <filesystem>
#include <fstream>
#include <stdexcept>

namespace rfk::buffer_cache {

buffer_group::buffer_group(const std::string& root, const std::string& type, const std::string& group) {
    path_ = (root.empty() ? std::string(".") : root) + "/cache/" + type + "/" + group + "/";
    std::filesystem::create_directories(path_);
}

std::string buffer_group::write_buffer(const void* data, std::size_t bytes, std::string name) const {
    if (name.empty()) {
        std::uint64_t h0 = 0xcbf29ce484222325ull, h1 = 0x84222325cbf29ce4ull;
        const unsigned char* p = static_cast<const unsigned char*>(data);
        for (std::size_t i = 0; i < bytes; i++) {
            h0 = (h0 ^ p[i]) * 0x100000001b3ull;
            h1 = (h1 ^ p[bytes - 1 - i]) * 0x100000001b3ull;
        }
        char buf[40];
        std::snprintf(buf, sizeof buf, "%016llX%016llX", (unsigned long long)h0, (unsigned long long)h1);
        name = buf;
    }
    std::ofstream f(path_ + name + ".bin", std::ios::binary | std::ios::out);
    if (!f) throw std::runtime_error("buffer_cache: cannot write " + path_ + name + ".bin");
    const std::size_t total_size = bytes;
    f.write(reinterpret_cast<const char*>(&total_size), sizeof(std::size_t));
    f.write(static_cast<const char*>(data), (std::streamsize)bytes);
    if (!f) throw std::runtime_error("buffer_cache: short write to " + path_ + name + ".bin");
    return name;
}

std::vector<char> buffer_group::read_buffer(const std::string& name) const {
    std::ifstream f(path_ + name + ".bin", std::ios::binary | std::ios::in);
    if (!f) throw std::runtime_error("buffer_cache: cannot read " + path_ + name + ".bin");
    std::size_t total_size = 0;
    f.read(reinterpret_cast<char*>(&total_size), sizeof(std::size_t));
    if (!f || total_size > (std::size_t(1) << 40)) throw std::runtime_error("buffer_cache: bad header in " + path_ + name + ".bin");
    std::vector<char> out(total_size);
    f.read(out.data(), (std::streamsize)total_size);
    if ((std::size_t)f.gcount() != total_size) throw std::runtime_error("buffer_cache: truncated payload in " + path_ + name + ".bin");
    return out;
}

std::vector<std::string> buffer_group::cached_buffers() const {
    std::vector<std::string> result;
    for (auto& f : std::filesystem::directory_iterator(path_))
        if (f.is_regular_file() && f.path().extension() == ".bin") result.push_back(f.path().stem().string());
    std::sort(result.begin(), result.end());
    return result;
}

}  // namespace rfk::buffer_cache
