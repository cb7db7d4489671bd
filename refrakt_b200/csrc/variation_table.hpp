// flame_compiler: loads variations.yaml and generates the per-genome dispatch()
// and get_xform_id() functions. Restates src/variation_table.{hpp,cpp}; emits the
// reference's GLSL text verbatim (compile_flame_xforms) and a CUDA dialect of the
// same text (compile_flame_cuda) that the sm_100a kernels are built from.
#pragma once
#include <map>
#include <set>
#include <string>
#include <vector>

#include "flame.hpp"

namespace rfk {

// src/variation_table.hpp:12-18
struct variation_definition {
    std::string source;
    std::string result;
    std::vector<std::string> param;
    std::set<std::string> flags;
};

class flame_compiler {
public:
    // Reads `path` (the reference reads "variations.yaml" from the CWD,
    // src/variation_table.cpp:183). Throws std::runtime_error on a missing or
    // malformed file, as yaml-cpp does.
    explicit flame_compiler(const std::string& path);
    // Same, from text already in memory.
    static flame_compiler from_text(const std::string& yaml_text);
    // Adds / replaces definitions from a second file of the same format (used for
    // the corrected definitions of the variations that do not compile in the
    // reference's table).
    void load_overlay_text(const std::string& yaml_text);

    bool is_param(const std::string& name) const { return param_owners_.count(name) != 0; }
    bool is_variation(const std::string& name) const { return vars_.count(name) != 0; }
    bool is_common(const std::string& name) const { return common_.count(name) != 0; }

    std::string param_owner(const std::string& param) const { return param_owners_.at(param); }
    const std::vector<std::string>& get_parameters_for_variation(const std::string& name) const { return vars_.at(name).param; }
    const variation_definition& variation(const std::string& name) const { return vars_.at(name); }
    const std::map<std::string, variation_definition>& variations() const { return vars_; }
    std::string common(const std::string& name) const { return common_.at(name); }

    // src/variation_table.cpp:217-265 — GLSL text of get_xform_id() + dispatch().
    std::string compile_flame_xforms(const flame& f) const;
    // The same functions in the CUDA dialect (float literals suffixed, swizzles as
    // calls, randf() draws of one statement sequenced left to right, explicit
    // fp / rng / first_run parameters).
    std::string compile_flame_cuda(const flame& f) const;

private:
    struct empty_tag {};
    explicit flame_compiler(empty_tag) {}
    void load_text(const std::string& yaml_text);

    std::map<std::string, std::string> common_;
    std::map<std::string, variation_definition> vars_;
    std::map<std::string, std::string> param_owners_;
};

// GLSL → CUDA-dialect rewrites, exposed for tests.
std::string suffix_float_literals(const std::string& glsl);
std::string swizzles_to_calls(const std::string& glsl);
std::string sequence_randf(const std::string& body, int& counter);

}  // namespace rfk
