#include "yaml_lite.hpp"

#include <stdexcept>

namespace rfk::yaml {

const node* node::find(const std::string& key) const {
    if (type != kind::map) return nullptr;
    for (auto& e : entries)
        if (e.first == key) return &e.second;
    return nullptr;
}

std::string node::as_string(const std::string& fallback) const {
    return type == kind::scalar ? scalar : fallback;
}

namespace {

struct line {
    int indent;        // leading spaces; -1 for blank / comment-only lines
    std::string text;  // content after the indent, trailing whitespace removed
    std::string raw;   // full line (for literal blocks)
    int number;
};

[[noreturn]] void fail(int ln, const std::string& what) {
    throw std::runtime_error("yaml: line " + std::to_string(ln) + ": " + what);
}

std::string rstrip(std::string s) {
    while (!s.empty() && (s.back() == ' ' || s.back() == '\t' || s.back() == '\r')) s.pop_back();
    return s;
}

std::string lstrip(const std::string& s) {
    std::size_t i = 0;
    while (i < s.size() && (s[i] == ' ' || s[i] == '\t')) i++;
    return s.substr(i);
}

// Removes a trailing ` # comment` from a plain (unquoted) value.
std::string strip_comment(const std::string& s) {
    for (std::size_t i = 0; i < s.size(); i++)
        if (s[i] == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return rstrip(s.substr(0, i));
    return s;
}

std::string unquote_double(const std::string& s, int ln) {
    std::string out;
    for (std::size_t i = 1; i < s.size(); i++) {
        char c = s[i];
        if (c == '"') return out;
        if (c == '\\' && i + 1 < s.size()) {
            char n = s[++i];
            switch (n) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case '"': out += '"'; break;
                case '\\': out += '\\'; break;
                default: out += n; break;
            }
        } else {
            out += c;
        }
    }
    fail(ln, "unterminated double-quoted scalar");
}

std::string unquote_single(const std::string& s, int ln) {
    std::string out;
    for (std::size_t i = 1; i < s.size(); i++) {
        if (s[i] == '\'') {
            if (i + 1 < s.size() && s[i + 1] == '\'') { out += '\''; i++; continue; }
            return out;
        }
        out += s[i];
    }
    fail(ln, "unterminated single-quoted scalar");
}

node parse_flow(const std::string& v, int ln) {
    node n;
    if (v.front() == '{') {
        if (rstrip(v).back() != '}') fail(ln, "unterminated flow map");
        std::string inner = lstrip(rstrip(v.substr(1, rstrip(v).size() - 2)));
        n.type = node::kind::map;
        if (!inner.empty()) fail(ln, "non-empty flow maps are not supported");
        return n;
    }
    if (rstrip(v).back() != ']') fail(ln, "unterminated flow sequence");
    std::string inner = rstrip(v).substr(1, rstrip(v).size() - 2);
    n.type = node::kind::seq;
    std::size_t start = 0;
    while (start <= inner.size()) {
        std::size_t comma = inner.find(',', start);
        std::string item = rstrip(lstrip(inner.substr(start, comma == std::string::npos ? std::string::npos : comma - start)));
        if (!item.empty()) {
            node s;
            s.type = node::kind::scalar;
            s.scalar = item;
            n.items.push_back(s);
        }
        if (comma == std::string::npos) break;
        start = comma + 1;
    }
    return n;
}

struct parser {
    std::vector<line> lines;
    std::size_t pos = 0;

    void skip_blank() {
        while (pos < lines.size() && lines[pos].indent < 0) pos++;
    }

    // literal block scalar whose header sat on a line indented `parent_indent`
    node parse_literal(bool strip, int parent_indent) {
        int block_indent = -1;
        std::string out;
        std::size_t pending_blank = 0;
        while (pos < lines.size()) {
            const line& l = lines[pos];
            std::string raw = l.raw;  // literal blocks keep trailing spaces
            while (!raw.empty() && raw.back() == '\r') raw.pop_back();
            std::size_t lead = 0;
            while (lead < raw.size() && raw[lead] == ' ') lead++;
            bool blank = rstrip(raw).empty();
            if (blank) { pending_blank++; pos++; continue; }
            if (block_indent < 0) {
                if ((int)lead <= parent_indent) break;
                block_indent = (int)lead;
            }
            if ((int)lead < block_indent) break;
            out.append(pending_blank, '\n');
            pending_blank = 0;
            out += raw.substr(block_indent);
            out += '\n';
            pos++;
        }
        if (strip) while (!out.empty() && out.back() == '\n') out.pop_back();
        node n;
        n.type = node::kind::scalar;
        n.scalar = out;
        return n;
    }

    node parse_value(const std::string& v, int ln, int key_indent) {
        node n;
        if (v.empty()) {
            // nested block or null
            skip_blank();
            if (pos < lines.size() && lines[pos].indent > key_indent) return parse_map(lines[pos].indent);
            return n;
        }
        if (v[0] == '|') {
            bool strip = v.size() > 1 && v[1] == '-';
            return parse_literal(strip, key_indent);
        }
        if (v[0] == '"') { n.type = node::kind::scalar; n.scalar = unquote_double(v, ln); return n; }
        if (v[0] == '\'') { n.type = node::kind::scalar; n.scalar = unquote_single(v, ln); return n; }
        if (v[0] == '{' || v[0] == '[') return parse_flow(v, ln);
        n.type = node::kind::scalar;
        n.scalar = strip_comment(v);
        // plain multi-line continuation: following lines indented deeper than the key
        for (;;) {
            std::size_t save = pos;
            skip_blank();
            if (pos < lines.size() && lines[pos].indent > key_indent) {
                n.scalar += " " + strip_comment(lines[pos].text);
                pos++;
            } else {
                pos = save;
                break;
            }
        }
        return n;
    }

    node parse_map(int indent) {
        node n;
        n.type = node::kind::map;
        for (;;) {
            skip_blank();
            if (pos >= lines.size()) break;
            const line& l = lines[pos];
            if (l.indent < indent) break;
            if (l.indent > indent) fail(l.number, "unexpected indentation");
            // key: value  (key is a plain identifier-like token)
            std::size_t colon = std::string::npos;
            for (std::size_t i = 0; i < l.text.size(); i++) {
                if (l.text[i] == ':' && (i + 1 == l.text.size() || l.text[i + 1] == ' ')) { colon = i; break; }
            }
            if (colon == std::string::npos) fail(l.number, "expected `key:`");
            std::string key = rstrip(l.text.substr(0, colon));
            std::string val = lstrip(l.text.substr(colon + 1));
            if (!val.empty() && val[0] == '#') val.clear();
            int ln = l.number;
            pos++;
            n.entries.emplace_back(key, parse_value(val, ln, indent));
        }
        return n;
    }
};

}  // namespace

node parse(const std::string& text) {
    parser p;
    std::size_t start = 0;
    int number = 0;
    while (start <= text.size()) {
        std::size_t nl = text.find('\n', start);
        std::string raw = text.substr(start, nl == std::string::npos ? std::string::npos : nl - start);
        number++;
        line l;
        l.raw = raw;
        l.number = number;
        std::string r = rstrip(raw);
        std::size_t lead = 0;
        while (lead < r.size() && r[lead] == ' ') lead++;
        l.text = r.substr(lead);
        l.indent = (l.text.empty() || l.text[0] == '#') ? -1 : (int)lead;
        p.lines.push_back(l);
        if (nl == std::string::npos) break;
        start = nl + 1;
    }
    p.skip_blank();
    if (p.pos >= p.lines.size()) return node{};
    return p.parse_map(p.lines[p.pos].indent);
}

}  // namespace rfk::yaml
