// TEST INFRASTRUCTURE. C wrappers over reference code compiled from /root/reference
// (see Makefile): make_sample_points (src/hammersley.cpp:29), jsf32::warmup_ctx (src/util.hpp:90),
// replace_macro / find_macros (src/util.cpp:6-23) and the header-inline reader of src/buffer_cache.hpp. Used only to pin the oracle's restatement of those functions.
#include <array>
#include <cstdint>
#include <vector>

#include "util.hpp"  // the reference's header: jsf32
#include "buffer_cache.hpp"  // the reference's header-inline reader: buffer_group::read_buffer / cached_buffers

std::vector<std::array<float, 4>> make_sample_points(std::uint32_t count);  // hammersley.cpp

#include <cstring>

extern "C" {
// src/util.cpp:6-23 (compiled from the reference): the macro grammar of the flame compiler
int ref_replace_macro(const char* str, const char* name, const char* value, char* out, int cap) {
    std::string r = replace_macro(str, name, value);
    if ((int)r.size() + 1 > cap) return -1;
    std::memcpy(out, r.c_str(), r.size() + 1);
    return (int)r.size();
}
int ref_find_macros(const char* str, char* out, int cap) {  // names joined by '\n', std::set order
    std::string joined;
    for (auto& m : find_macros(str)) joined += m + "\n";
    if ((int)joined.size() + 1 > cap) return -1;
    std::memcpy(out, joined.c_str(), joined.size() + 1);
    return (int)joined.size();
}
// src/buffer_cache.hpp:12-60 (header-inline, compiled from the reference): lists and reads cache/<type>/<group>/*.bin
// relative to the working directory — used to check that files written by the product are readable by refrakt
int ref_cache_list(const char* type, const char* group, char* out, int cap) {
    std::string joined;
    for (auto& n : buffer_cache::buffer_group(type, group).cached_buffers()) joined += n + "\n";
    if ((int)joined.size() + 1 > cap) return -1;
    std::memcpy(out, joined.c_str(), joined.size() + 1);
    return (int)joined.size();
}
long ref_cache_read_u32(const char* type, const char* group, const char* name, std::uint32_t* out, long cap) {
    auto data = buffer_cache::buffer_group(type, group).read_buffer<std::uint32_t>(name);
    if ((long)data.size() > cap) return -1;
    std::memcpy(out, data.data(), data.size() * 4);
    return (long)data.size();
}
void ref_make_sample_points(std::uint32_t count, float* out) {
    auto pts = make_sample_points(count);
    for (std::uint32_t i = 0; i < count; i++)
        for (int k = 0; k < 4; k++) out[4 * i + k] = pts[i][k];
}
void ref_jsf32_warmup(std::uint32_t seed, std::uint32_t* out4) {
    jsf32::ctx c;
    jsf32::warmup_ctx(c, seed);
    out4[0] = c.a; out4[1] = c.b; out4[2] = c.c; out4[3] = c.d;
}
std::uint32_t ref_jsf32_ranval(std::uint32_t* state4) {
    jsf32::ctx c{state4[0], state4[1], state4[2], state4[3]};
    std::uint32_t r = jsf32::ranval(c);
    state4[0] = c.a; state4[1] = c.b; state4[2] = c.c; state4[3] = c.d;
    return r;
}
}
