// TEST INFRASTRUCTURE. Stand-in for inja 3.1. The two templates the reference renders (xform_select.tpl.glsl,
// animate.tpl.glsl) are not evaluated in C++: render() returns the JSON data and the template between sentinels, and
// oracle/softgl/glsl_to_cpp.py (render_inja) evaluates the subset of inja they use when it turns the shader into a library.
#pragma once
#include <nlohmann/json.hpp>
#include <string>
namespace inja {
inline std::string render(const std::string& tpl, const nlohmann::json& data) {
    return "\x01INJA\x02" + data.dump() + "\x02" "DATA\x02" + tpl + "\x03";
}
}  // namespace inja
