// TEST INFRASTRUCTURE. Stand-in for fmt 7 (the reference uses one fmt::format call with "{}" placeholders,
// src/variation_table.cpp:242; buffer_cache.cpp's use is not compiled).
#pragma once
#include <sstream>
#include <string>
namespace fmt {
inline void format_rest(std::ostringstream& o, const char* f) { o << f; }
template <typename T, typename... R>
void format_rest(std::ostringstream& o, const char* f, const T& v, const R&... rest) {
    for (; *f; f++) {
        if (f[0] == '{' && f[1] == '}') { o << v; format_rest(o, f + 2, rest...); return; }
        o << *f;
    }
}
template <typename... A>
std::string format(const char* f, const A&... args) { std::ostringstream o; format_rest(o, f, args...); return o.str(); }
}  // namespace fmt
