// On-disk buffer cache in the reference's file format (src/buffer_cache.hpp:12-60, src/buffer_cache.cpp:7-21):
//   <root>/cache/<type>/<group>/<name>.bin  =  size_t byte count, then the raw payload.
// The reference names new files by the XXH3-128 of the payload and finds buffers by listing the directory, so any
// unique name interoperates; names written here are the same XXH3-128 where the system carries libxxhash (looked up at run
// time), else a 128-bit FNV-1a pair in the same 32-hex-digit form.
// Lets this core produce shuffle / RNG-state buffers a real refrakt build consumes, and consume the ones it cached
// (seeded A/B runs). The hot path does not need the cache: seeding runs on the device.
#pragma once
#include <cstddef>
#include <string>
#include <vector>

namespace rfk::buffer_cache {

class buffer_group {
public:
    // root: directory that holds "cache/" (the reference uses the working directory, i.e. ".")
    buffer_group(const std::string& root, const std::string& type, const std::string& group);
    // returns the buffer name (file stem); an empty `name` derives one from the payload
    std::string write_buffer(const void* data, std::size_t bytes, std::string name = {}) const;
    std::vector<char> read_buffer(const std::string& name) const;
    std::vector<std::string> cached_buffers() const;  // sorted, for reproducibility
    const std::string& path() const { return path_; }

private:
    std::string path_;
};

}  // namespace rfk::buffer_cache
