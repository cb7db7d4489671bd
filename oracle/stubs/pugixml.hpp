// TEST INFRASTRUCTURE. Stand-in for pugixml 1.10 over the product's XML reader (refrakt_b200/csrc/xml_lite.*, itself
// checked against ElementTree in tests/test_host_cpu.py): the node / attribute surface src/flame.cpp:161-217 uses, with
// pugixml's own conversion rules (as_float = (float)strtod, as_int / as_ullong = strtol-style, 0 on a missing attribute).
#pragma once
#include <cstdlib>
#include <string>
#include <vector>
#include "../../refrakt_b200/csrc/xml_lite.hpp"
#include "../../refrakt_b200/csrc/textutil.hpp"
namespace pugi {
class xml_attribute {
public:
    xml_attribute() = default;
    explicit xml_attribute(const std::pair<std::string, std::string>* a) : a_(a) {}
    const char* name() const { return a_ ? a_->first.c_str() : ""; }
    const char* value() const { return a_ ? a_->second.c_str() : ""; }
    const char* as_string() const { return value(); }
    float as_float() const { return a_ ? (float)std::strtod(a_->second.c_str(), nullptr) : 0.0f; }
    int as_int() const { return a_ ? (int)std::strtol(a_->second.c_str(), nullptr, 10) : 0; }
    unsigned long long as_ullong() const { return a_ ? std::strtoull(a_->second.c_str(), nullptr, 10) : 0ull; }
    explicit operator bool() const { return a_ != nullptr; }
private:
    const std::pair<std::string, std::string>* a_ = nullptr;
};
class xml_node {
public:
    xml_node() = default;
    explicit xml_node(const rfk::xml::element* e) : e_(e) {}
    const char* name() const { return e_ ? e_->name.c_str() : ""; }
    xml_attribute attribute(const char* key) const {
        if (e_) for (auto& a : e_->attributes) if (a.first == key) return xml_attribute(&a);
        return xml_attribute();
    }
    xml_node child(const char* key) const {
        if (e_) for (auto& c : e_->children) if (c.name == key) return xml_node(&c);
        return xml_node();
    }
    std::vector<xml_node> children() const {
        std::vector<xml_node> out;
        if (e_) for (auto& c : e_->children) out.emplace_back(&c);
        return out;
    }
    std::vector<xml_attribute> attributes() const {
        std::vector<xml_attribute> out;
        if (e_) for (auto& a : e_->attributes) out.emplace_back(&a);
        return out;
    }
    explicit operator bool() const { return e_ != nullptr; }
protected:
    const rfk::xml::element* e_ = nullptr;
};
struct xml_parse_result { bool ok = false; explicit operator bool() const { return ok; } };
class xml_document : public xml_node {
public:
    xml_parse_result load_file(const char* path) {
        bool ok = false;
        std::string text = rfk::read_file(path, &ok);
        if (!ok) return {};
        try {
            doc_.name = "#document";
            doc_.children.clear();
            doc_.children.push_back(rfk::xml::parse(text));
        } catch (...) { return {}; }
        e_ = &doc_;
        return {true};
    }
private:
    rfk::xml::element doc_;
};
}  // namespace pugi
