#include "buffer_cache.hpp"

#include <dlfcn.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <filesystem>
#include <fstream>
#include <stdexcept>

namespace rfk::buffer_cache {

namespace {
struct xxh128 { std::uint64_t low64, high64; };  // XXH128_hash_t
using xxh3_128_fn = xxh128 (*)(const void*, std::size_t);
xxh3_128_fn system_xxh3_128() {
    static const xxh3_128_fn fn = [] {
        void* lib = dlopen("libxxhash.so.0", RTLD_NOW | RTLD_LOCAL);
        return lib ? reinterpret_cast<xxh3_128_fn>(dlsym(lib, "XXH3_128bits")) : nullptr;
    }();
    return fn;
}
}  // namespace

buffer_group::buffer_group(const std::string& root, const std::string& type, const std::string& group) {
    path_ = (root.empty() ? std::string(".") : root) + "/cache/" + type + "/" + group + "/";
    std::filesystem::create_directories(path_);
}

std::string buffer_group::write_buffer(const void* data, std::size_t bytes, std::string name) const {
    if (name.empty()) {
        // the reference's name (buffer_cache.cpp:10-11): XXH3_128bits of the payload as {high64:016X}{low64:016X}. xxHash is a
        // system library here as it is a dependency there (xxhash 0.8); it is looked up at run time, and a machine without it
        // gets two FNV-1a hashes instead — the reference lists the directory and never derives a name, so either is readable
        std::uint64_t h0 = 0xcbf29ce484222325ull, h1 = 0x84222325cbf29ce4ull;
        if (const xxh3_128_fn xxh3 = system_xxh3_128()) {
            const xxh128 h = xxh3(data, bytes);
            h0 = h.high64; h1 = h.low64;
        } else {
            const unsigned char* p = static_cast<const unsigned char*>(data);
            for (std::size_t i = 0; i < bytes; i++) {
                h0 = (h0 ^ p[i]) * 0x100000001b3ull;
                h1 = (h1 ^ p[bytes - 1 - i]) * 0x100000001b3ull;
            }
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
