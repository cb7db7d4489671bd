#include "xml_lite.hpp"

#include <cstdlib>
#include <stdexcept>

namespace rfk::xml {

const std::string* element::attribute(const std::string& key) const {
    for (auto& a : attributes)
        if (a.first == key) return &a.second;
    return nullptr;
}

const element* element::child(const std::string& key) const {
    for (auto& c : children)
        if (c.name == key) return &c;
    return nullptr;
}

namespace {

struct cursor {
    const std::string& s;
    std::size_t i = 0;

    [[noreturn]] void fail(const std::string& what) const {
        std::size_t line = 1;
        for (std::size_t k = 0; k < i && k < s.size(); k++)
            if (s[k] == '\n') line++;
        throw std::runtime_error("xml: line " + std::to_string(line) + ": " + what);
    }
    bool eof() const { return i >= s.size(); }
    bool starts(const char* lit) const { return s.compare(i, std::char_traits<char>::length(lit), lit) == 0; }
    void skip_ws() {
        while (!eof() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) i++;
    }
    void skip_until(const char* lit) {
        std::size_t p = s.find(lit, i);
        if (p == std::string::npos) fail(std::string("missing `") + lit + "`");
        i = p + std::char_traits<char>::length(lit);
    }
};

bool name_char(char c) {
    return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' || c == '-' || c == ':' || c == '.';
}

std::string decode_entities(const std::string& v) {
    if (v.find('&') == std::string::npos) return v;
    std::string out;
    for (std::size_t i = 0; i < v.size(); i++) {
        if (v[i] != '&') { out += v[i]; continue; }
        std::size_t semi = v.find(';', i);
        if (semi == std::string::npos) { out += v[i]; continue; }
        std::string ent = v.substr(i + 1, semi - i - 1);
        if (ent == "amp") out += '&';
        else if (ent == "lt") out += '<';
        else if (ent == "gt") out += '>';
        else if (ent == "quot") out += '"';
        else if (ent == "apos") out += '\'';
        else if (!ent.empty() && ent[0] == '#') {
            long code = (ent.size() > 1 && (ent[1] == 'x' || ent[1] == 'X')) ? std::strtol(ent.c_str() + 2, nullptr, 16) : std::strtol(ent.c_str() + 1, nullptr, 10);
            if (code > 0 && code < 128) out += char(code);
        } else { out += v.substr(i, semi - i + 1); }
        i = semi;
    }
    return out;
}

// Skips comments / PIs / doctype / text until the next element start or end tag.
void skip_misc(cursor& c) {
    for (;;) {
        while (!c.eof() && c.s[c.i] != '<') c.i++;
        if (c.eof()) return;
        if (c.starts("<!--")) c.skip_until("-->");
        else if (c.starts("<?")) c.skip_until("?>");
        else if (c.starts("<![CDATA[")) c.skip_until("]]>");
        else if (c.starts("<!")) c.skip_until(">");
        else return;
    }
}

element parse_element(cursor& c) {
    // at '<' of a start tag
    c.i++;
    element e;
    while (!c.eof() && name_char(c.s[c.i])) e.name += c.s[c.i++];
    if (e.name.empty()) c.fail("expected element name");
    for (;;) {
        c.skip_ws();
        if (c.eof()) c.fail("unterminated start tag");
        if (c.starts("/>")) { c.i += 2; return e; }
        if (c.s[c.i] == '>') { c.i++; break; }
        std::string key;
        while (!c.eof() && name_char(c.s[c.i])) key += c.s[c.i++];
        if (key.empty()) c.fail("expected attribute name");
        c.skip_ws();
        if (c.eof() || c.s[c.i] != '=') c.fail("expected `=` after attribute name");
        c.i++;
        c.skip_ws();
        if (c.eof() || (c.s[c.i] != '"' && c.s[c.i] != '\'')) c.fail("expected quoted attribute value");
        char q = c.s[c.i++];
        std::size_t end = c.s.find(q, c.i);
        if (end == std::string::npos) c.fail("unterminated attribute value");
        e.attributes.emplace_back(key, decode_entities(c.s.substr(c.i, end - c.i)));
        c.i = end + 1;
    }
    // content
    for (;;) {
        skip_misc(c);
        if (c.eof()) c.fail("missing end tag for <" + e.name + ">");
        if (c.starts("</")) {
            c.skip_until(">");
            return e;
        }
        e.children.push_back(parse_element(c));
    }
}

}  // namespace

element parse(const std::string& text) {
    cursor c{text};
    skip_misc(c);
    if (c.eof()) throw std::runtime_error("xml: no root element");
    return parse_element(c);
}

float as_float(const std::string* v) { return v ? (float)std::strtod(v->c_str(), nullptr) : 0.0f; }
int as_int(const std::string* v) { return v ? (int)std::strtol(v->c_str(), nullptr, 10) : 0; }
unsigned long long as_ullong(const std::string* v) { return v ? std::strtoull(v->c_str(), nullptr, 10) : 0ull; }

}  // namespace rfk::xml
