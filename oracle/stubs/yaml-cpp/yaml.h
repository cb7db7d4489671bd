// TEST INFRASTRUCTURE. Stand-in for yaml-cpp 0.6.3 over the product's YAML-subset reader (refrakt_b200/csrc/yaml_lite.*,
// itself checked against PyYAML in tests/test_host_cpu.py): just the Node surface src/variation_table.cpp:183-214 uses.
#pragma once
#include <string>
#include <utility>
#include <vector>
#include "../../../refrakt_b200/csrc/yaml_lite.hpp"
#include "../../../refrakt_b200/csrc/textutil.hpp"
namespace YAML {
class Node;
struct iterator_value;
class Node {
public:
    Node() = default;
    explicit Node(const rfk::yaml::node* n) : n_(n) {}
    Node operator[](const char* key) const { return Node(n_ ? n_->find(key) : nullptr); }
    template <typename T> T as() const;
    template <typename T, typename S> T as(const S& fallback) const;
    class const_iterator;
    const_iterator begin() const;
    const_iterator end() const;
    const rfk::yaml::node* raw() const { return n_; }
private:
    const rfk::yaml::node* n_ = nullptr;
};
struct iterator_value : public Node, public std::pair<Node, Node> {
    iterator_value() = default;
    explicit iterator_value(const Node& v) : Node(v) {}
    iterator_value(const Node& k, const Node& v) : std::pair<Node, Node>(k, v) {}
};
class Node::const_iterator {
public:
    const_iterator(const rfk::yaml::node* n, std::size_t i) : n_(n), i_(i) {}
    bool operator!=(const const_iterator& o) const { return i_ != o.i_; }
    const_iterator& operator++() { i_++; return *this; }
    const_iterator operator++(int) { auto c = *this; i_++; return c; }
    iterator_value operator*() const {
        if (n_->type == rfk::yaml::node::kind::map) {
            key_.type = rfk::yaml::node::kind::scalar;
            key_.scalar = n_->entries[i_].first;
            return iterator_value(Node(&key_), Node(&n_->entries[i_].second));
        }
        return iterator_value(Node(&n_->items[i_]));
    }
    struct proxy { iterator_value v; iterator_value* operator->() { return &v; } };
    proxy operator->() const { return proxy{**this}; }
private:
    const rfk::yaml::node* n_;
    std::size_t i_;
    mutable rfk::yaml::node key_;
};
inline Node::const_iterator Node::begin() const { return const_iterator(n_, 0); }
inline Node::const_iterator Node::end() const {
    if (!n_) return const_iterator(n_, 0);
    return const_iterator(n_, n_->type == rfk::yaml::node::kind::map ? n_->entries.size() : n_->items.size());
}
template <> inline std::string Node::as<std::string>() const { return n_ ? n_->as_string() : std::string(); }
template <> inline Node Node::as<Node>() const { return *this; }
template <> inline std::string Node::as<std::string, char[1]>(const char (&fallback)[1]) const {
    return (n_ && n_->type == rfk::yaml::node::kind::scalar) ? n_->scalar : std::string(fallback);
}
inline Node LoadFile(const std::string& path) {
    static std::vector<rfk::yaml::node*> keep;  // documents live for the process
    bool ok = false;
    std::string text = rfk::read_file(path, &ok);
    if (!ok) throw std::runtime_error("yaml stand-in: cannot read " + path);
    keep.push_back(new rfk::yaml::node(rfk::yaml::parse(text)));
    return Node(keep.back());
}
}  // namespace YAML
