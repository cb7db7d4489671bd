// Minimal YAML reader: the subset variations.yaml is written in (block maps by
// indentation, plain / double-quoted / single-quoted scalars, `|` and `|-` literal
// blocks, `{}` and `[a, b]` flow collections, `#` comments). Stands in for
// yaml-cpp 0.6.3, which the reference uses only to read that file
// (src/variation_table.cpp:183-214). Map entries keep document order.
#pragma once
#include <string>
#include <utility>
#include <vector>

namespace rfk::yaml {

struct node {
    enum class kind { null, scalar, map, seq };
    kind type = kind::null;
    std::string scalar;
    std::vector<std::pair<std::string, node>> entries;  // map, document order
    std::vector<node> items;                            // seq

    bool is_null() const { return type == kind::null; }
    bool is_map() const { return type == kind::map; }
    // nullptr when absent or when this node is not a map
    const node* find(const std::string& key) const;
    // scalar text or `fallback` (mirrors YAML::Node::as<std::string>(fallback))
    std::string as_string(const std::string& fallback = "") const;
};

// Throws std::runtime_error with a line number on malformed input.
node parse(const std::string& text);

}  // namespace rfk::yaml
