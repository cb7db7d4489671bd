#include "textutil.hpp"

#include <fstream>
#include <iterator>
#include <sstream>

namespace rfk {

static inline bool is_ident(char c) {
    return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_';
}

std::string replace_macro(const std::string& str, const std::string& name, const std::string& value) {
    std::string out;
    out.reserve(str.size() + 16);
    const std::string needle = "$" + name;
    std::size_t i = 0;
    while (i < str.size()) {
        if (str.compare(i, needle.size(), needle) == 0) {
            std::size_t after = i + needle.size();
            if (after < str.size() && !is_ident(str[after])) {
                out += value;
                out += str[after];
                i = after + 1;  // the delimiter is part of the match
                continue;
            }
        }
        out += str[i++];
    }
    return out;
}

std::set<std::string> find_macros(const std::string& str) {
    std::set<std::string> result;
    std::size_t i = 0;
    while (i < str.size()) {
        if (str[i] == '$') {
            std::size_t j = i + 1;
            while (j < str.size() && ((str[j] >= 'a' && str[j] <= 'z') || (str[j] >= '0' && str[j] <= '9') || str[j] == '_')) j++;
            if (j > i + 1) {
                result.insert(str.substr(i + 1, j - i - 1));
                i = j;
                continue;
            }
        }
        i++;
    }
    return result;
}

std::string replace_all(std::string str, const std::string& from, const std::string& to) {
    if (from.empty()) return str;
    std::size_t pos = 0;
    while ((pos = str.find(from, pos)) != std::string::npos) {
        str.replace(pos, from.size(), to);
        pos += to.size();
    }
    return str;
}

std::string read_file(const std::string& path, bool* ok) {
    std::ifstream f(path, std::ios::binary);
    if (ok) *ok = bool(f);
    if (!f) return {};
    return std::string(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}

std::vector<std::string> split_ws(const std::string& s) {
    std::vector<std::string> out;
    std::istringstream ss{s};
    std::string v;
    while (ss >> v) out.push_back(v);
    return out;
}

}  // namespace rfk
