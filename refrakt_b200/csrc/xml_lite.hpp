// Minimal XML element/attribute reader for flam3 genomes. Stands in for pugixml
// v1.10, which the reference uses only to walk elements and attributes
// (src/flame.cpp:161-217). Attributes and children keep document order.
#pragma once
#include <string>
#include <utility>
#include <vector>

namespace rfk::xml {

struct element {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attributes;
    std::vector<element> children;

    // nullptr when absent
    const std::string* attribute(const std::string& key) const;
    const element* child(const std::string& key) const;
};

// Parses the first root element of the document. Throws std::runtime_error on
// malformed input. Text content, comments, CDATA and processing instructions are skipped.
element parse(const std::string& text);

// pugixml attribute conversions (as_float / as_int / as_ullong): strtod / strtol
// style, 0 on an absent or non-numeric attribute.
float as_float(const std::string* v);
int as_int(const std::string* v);
unsigned long long as_ullong(const std::string* v);

}  // namespace rfk::xml
