// TEST INFRASTRUCTURE. Stand-in for inja 3.1: the two templates the reference renders (xform_select.tpl.glsl,
// animate.tpl.glsl) are not evaluated here; render() returns a marker naming the template's first line, so the text the
// reference builds AROUND the rendered template can be compared (the templates are restated in the oracle / product).
#pragma once
#include <nlohmann/json.hpp>
#include <string>
namespace inja {
inline std::string render(const std::string& tpl, const nlohmann::json&) {
    return "/*inja:" + tpl.substr(0, tpl.find('\n')) + "*/";
}
}  // namespace inja
