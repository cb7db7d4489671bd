#include "variation_table.hpp"

#include <algorithm>
#include <cstdlib>
#include <functional>
#include <regex>
#include <set>
#include <stdexcept>

#include "textutil.hpp"
#include "yaml_lite.hpp"

namespace rfk {

namespace {
inline bool ident_char(char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_'; }
inline bool digit_char(char c) { return c >= '0' && c <= '9'; }

std::string slot_str(int slot) { return "fp[" + std::to_string(slot) + "]"; }  // variation_table.cpp:21

// name -> the names its definition mentions
using dependency_map = std::map<std::string, std::set<std::string>>;

// The order in which an xform's precalcs (`r`, `rsq`, `phi`, ... of variations.yaml `common`) are declared: every name after
// the names its own definition uses, alphabetical otherwise — the order the reference's generated text has (it pushes a
// depth-first walk on a stack and prepends while popping, variation_table.cpp:25-47 and :130-147; the emitted GLSL has to
// be character-identical, so the order is part of the contract).
std::vector<std::string> declaration_order(const dependency_map& uses) {
    std::vector<std::string> order;
    std::set<std::string> placed;
    std::function<void(const std::string&)> place = [&](const std::string& name) {
        if (!placed.insert(name).second) return;
        for (const std::string& needed : uses.at(name)) place(needed);
        order.push_back(name);
    };
    for (const auto& entry : uses) place(entry.first);
    return order;
}

// $cCR -> affine slot C*2+R (variation_table.cpp:49-60); $pCR likewise for post (:62-73).
void resolve_coefs(std::string& src, const std::array<int, 6>& slots, char prefix) {
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 2; r++) {
            std::string name = std::string(1, prefix) + std::to_string(c) + std::to_string(r);
            src = replace_macro(src, name, slot_str(slots[c * 2 + r]));
        }
}

bool is_linear_only(const flame_xform& x) {  // variation_table.cpp:85-87
    return !x.post && x.variations.size() == 1 && x.variations.begin()->first == "linear";
}

// The reference's two "optimizers" (variation_table.cpp:78-169): returns {inlined, text}.
std::pair<bool, std::string> make_xform_text(const flame_xform& x, const xform_slots& xmap, const flame_compiler& vt) {
    if (is_linear_only(x)) {
        std::string src = "$weight * vec2(fma($c00, v.x, fma($c10, v.y, $c20)), fma($c01, v.x, fma($c11, v.y, $c21)))";
        src = replace_macro(src, "weight", slot_str(xmap.variations.at("linear")));
        resolve_coefs(src, xmap.affine, 'c');
        return {true, src};
    }

    std::string xform_result;
    bool first_var = true;
    std::string affine = "\tv.xy = vec2(fma($c00, v.x, fma($c10, v.y, $c20)), fma($c01, v.x, fma($c11, v.y, $c21)));\n";

    for (auto& [var_name, weight] : x.variations) {
        const auto& vd = vt.variation(var_name);
        const std::string wslot = slot_str(xmap.variations.at(var_name));
        std::string var_src;
        var_src += "// variation: " + var_name + "\n";
        if (!vd.source.empty()) var_src += replace_macro(vd.source, "weight", wslot) + "\n";

        std::string weight_str = vd.flags.count("no_weight_mul") ? "" : "$weight *";

        if (vd.flags.count("pre_xform")) {
            affine += replace_macro(("v.xy += " + weight_str) + vd.result + ";", "weight", wslot) + "\n";
        } else {
            var_src += replace_macro((first_var ? "vec2 result = " + weight_str : "result += " + weight_str) + vd.result + ";", "weight", wslot) + "\n";
            xform_result += var_src;
            first_var = false;
        }
    }

    for (auto& [p_name, val] : x.var_param) xform_result = replace_macro(xform_result, p_name, slot_str(xmap.param.at(p_name)));

    // precalc macros in dependency order (variation_table.cpp:130-147)
    auto macros = find_macros(xform_result);
    std::erase_if(macros, [&vt](const std::string& name) { return !vt.is_common(name); });

    dependency_map uses;
    for (auto& m : macros) {
        auto deps = find_macros(vt.common(m));
        for (auto& d : deps)
            if (vt.is_common(d) && !uses.count(d)) uses[d] = find_macros(vt.common(m));  // (sic: the dependent's set, as variation_table.cpp:133-137 stores it)
        uses[m] = find_macros(vt.common(m));
    }

    std::string declarations;
    for (const std::string& name : declaration_order(uses)) declarations += "float " + name + " = " + vt.common(name) + ";\n";
    xform_result = declarations + xform_result;

    xform_result = affine + xform_result;
    resolve_coefs(xform_result, xmap.affine, 'c');

    if (x.post) {
        xform_result += "\tresult = vec2(fma($p00, result.x, fma($p10, result.y, $p20)), fma($p01, result.x, fma($p11, result.y, $p21)));\n";
        resolve_coefs(xform_result, xmap.post, 'p');
    }

    xform_result = replace_all(xform_result, "\n", "\n\t");
    xform_result = replace_all(xform_result, "$", "");
    return {false, xform_result};
}

// The substitutions every snippet of variations.yaml gets when the table is loaded: the particle's coordinates by name
// ($x, $y, $v) and $result, in this order (what variation_table.cpp:174-180 does before anything else sees the text).
void expand_particle_macros(std::string& text) {
    static const std::pair<const char*, const char*> table[] = {{"x", "v.x"}, {"y", "v.y"}, {"v", "v.xy"}, {"result", "result"}};
    for (const auto& [name, value] : table) text = replace_macro(text, name, value);
}

std::string xform_select_text(const buffer_map_t& map, bool cuda) {  // shaders/templates/xform_select.tpl.glsl
    std::string s = cuda ? "__device__ __forceinline__ int get_xform_id(float ratio) {\n\n"
                         : "int get_xform_id(float ratio) {\n\n";
    const int n = (int)map.xforms.size();
    for (int i = 0; i < n; i++) {
        const std::string w = slot_str(map.xforms[i].weight);
        if (i == 0) {
            s += "\tfloat sum = " + w + ";\n\tif(sum >= ratio) return 0;\n";
        } else if (i == n - 1) {
            s += "\treturn " + std::to_string(i) + ";\n";
        } else {
            s += "\tsum += " + w + ";\n\tif(sum >= ratio) return " + std::to_string(i) + ";\n";
        }
    }
    // a one-xform genome leaves the template without a final return; every pick is xform 0
    if (cuda && n <= 1) s += "\treturn 0;\n";
    s += "}";
    return s;
}

enum class dialect { glsl, cuda };

// The affine lines the emitter writes (pre-affine of every xform, post-affine),
//   vec2(fma(a, X, fma(c, Y, e)), fma(b, X, fma(d, Y, f)))
// become rfk_affine(a, b, c, d, e, f, X, Y): the same nested fmas, two lanes per instruction (device_prelude.cuh).
std::string pack_affines(const std::string& s) {
    static const std::regex re(
        R"(vec2\(fma\(([^,()]+), ([A-Za-z_][\w.]*), fma\(([^,()]+), ([A-Za-z_][\w.]*), ([^,()]+)\)\), fma\(([^,()]+), \2, fma\(([^,()]+), \4, ([^,()]+)\)\)\))");
    return std::regex_replace(s, re, "rfk_affine($1, $6, $3, $7, $5, $8, $2, $4)");
}

std::string emit(const flame& f, const flame_compiler& vt, dialect d) {
    const auto& buf_map = f.buffer_map();
    std::string disp_func = d == dialect::glsl
        ? std::string("vec4 dispatch(vec3 v, int xform){\n").append("switch(xform){\n")
        : std::string("template <bool first_run>\n__device__ __forceinline__ vec4 dispatch_a(vec3 v, int xform, rfk_rng& rs, const float4 rfk_A){\n").append("switch(xform){\n");
    int rf_counter = 0;
    std::vector<std::pair<int, std::string>> cuda_cases;  // (xform index, body)

    for (int i = -1; i < (int)f.xforms.size(); i++) {
        if (i == -1 && !f.final_xform) continue;
        const auto& xform = (i == -1) ? f.final_xform.value() : f.xforms.at(i);
        const auto& xmap = (i == -1) ? buf_map.final_xform.value() : buf_map.xforms.at(i);

        auto [inlined, xform_src] = make_xform_text(xform, xmap, vt);

        // colour blend mix(z, colour, speed) = z * (1 - speed) + colour * speed. The CUDA dialect spells it RFK_MIXC with the
        // slots of the two derived constants 1 - speed and colour * speed (uploaded by the host behind the slots and their
        // reciprocals: rfk_cfp[2 * size + speed], rfk_cfp[3 * size + speed]; needs colour and speed in adjacent slots):
        // one FFMA instead of FADD + FMUL + FFMA on warp-uniform values in every iteration.
        const int size = std::max(1, buf_map.size);
        std::string blend = "mix(((first_run)? randf(): v.z), " + slot_str(xmap.color) + ", " + slot_str(xmap.color_speed) + ")";
        if (d == dialect::cuda && xmap.color_speed == xmap.color + 1)
            blend = "RFK_MIXC(((first_run)? randf(): v.z), " + slot_str(xmap.color) + ", " + slot_str(xmap.color_speed) + ", rfk_cfp[" +
                    std::to_string(2 * size + xmap.color_speed) + "], rfk_cfp[" + std::to_string(3 * size + xmap.color_speed) + "])";
        std::string dispatch_invoke = "return vec4(" + std::string(inlined ? xform_src : "result") + ", " + blend + ", " + slot_str(xmap.opacity) + ");\n";
        if (!inlined) dispatch_invoke = xform_src + dispatch_invoke;

        if (d == dialect::cuda) {
            dispatch_invoke = suffix_float_literals(dispatch_invoke);
            dispatch_invoke = swizzles_to_calls(dispatch_invoke);
            dispatch_invoke = sequence_randf(dispatch_invoke, rf_counter);
            dispatch_invoke = pack_affines(dispatch_invoke);
        }

        if (d == dialect::cuda) cuda_cases.emplace_back(i, dispatch_invoke);
        if (i + 1 == (int)f.xforms.size()) disp_func += "default: {\n" + dispatch_invoke + "\n}}\n";
        else disp_func += "case " + std::to_string(i) + ": {\n" + dispatch_invoke + "\n}\n";
    }
    // CUDA dialect. Every xform's text becomes a function of its own, rfk_xform_<k> (k = m1 for the final xform), and two
    // dispatchers call them: dispatch_a(v, xform, ...) for one particle and dispatch2_a(v0, v1, xform, ...) for the two
    // particles a thread of the paired kernels holds — one pick, one walk to the case, the case applied twice.
    // When a few xforms take most of the picks the walk is a chain of `if (xform == k)` in order of decreasing weight instead
    // of a switch (whose compare tree + jump table costs ~9 instructions per pick whatever the weights): E[tests] = sum over
    // the order of position x weight, two instructions per test. The pick is warp-uniform, so either form is one taken path
    // per warp. Same cases, same text inside them.
    if (d == dialect::cuda) {
        auto fn_name = [](int index) { return index < 0 ? std::string("rfk_xform_m1") : "rfk_xform_" + std::to_string(index); };
        std::string functions;
        for (const auto& c : cuda_cases)
            functions += "template <bool first_run>\n__device__ __forceinline__ vec4 " + fn_name(c.first) + "(vec3 v, rfk_rng& rs, const float4 rfk_A){\n" + c.second + "\n}\n";
        std::vector<std::pair<double, int>> order;
        double total = 0.0;
        for (std::size_t i = 0; i < f.xforms.size(); i++) total += std::max(0.0f, f.xforms[i].weight);
        for (std::size_t i = 0; i < f.xforms.size(); i++) order.emplace_back(total > 0.0 ? std::max(0.0f, f.xforms[i].weight) / total : 0.0, (int)i);
        std::stable_sort(order.begin(), order.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
        double expected_tests = 0.0;
        for (std::size_t k = 0; k < order.size(); k++) expected_tests += order[k].first * (double)std::min(k + 1, order.size() - 1);
        const char* force = std::getenv("RFK_DISPATCH");  // "switch" / "chain": A/B runs
        const bool chain_wanted = force ? std::string(force) == "chain" : (f.xforms.size() >= 3 && expected_tests <= 3.5);
        // `one` / `two`: the statement that applies case k to one particle / to both
        auto one = [&](int k) { return "return " + fn_name(k) + "<first_run>(v, rs, rfk_A);"; };
        auto two = [&](int k) { return "o0 = " + fn_name(k) + "<first_run>(v0, rs0, rfk_A); o1 = " + fn_name(k) + "<first_run>(v1, rs1, rfk_A); return;"; };
        auto walk = [&](auto&& apply, const std::string& nothing) {
            std::string w;
            if (f.xforms.empty()) {
                if (f.final_xform) w += "if (xform == -1) { " + apply(-1) + " }\n";
                return w + nothing + "\n";
            }
            if (chain_wanted) {
                // RFK_OPAQUE between the tests: the optimiser would otherwise fold the chain back into a switch
                w += "int rfk_pick = xform;\n";
                if (f.final_xform) w += "if (rfk_pick == -1) { " + apply(-1) + " }\n";
                for (std::size_t k = 0; k + 1 < order.size(); k++)
                    w += "RFK_OPAQUE(rfk_pick);\nif (rfk_pick == " + std::to_string(order[k].second) + ") { " + apply(order[k].second) + " }\n";
                return w + "{ " + apply(order.back().second) + " }\n";
            }
            w += "switch(xform){\n";
            if (f.final_xform) w += "case -1: { " + apply(-1) + " }\n";
            for (int i = 0; i + 1 < (int)f.xforms.size(); i++) w += "case " + std::to_string(i) + ": { " + apply(i) + " }\n";
            return w + "default: { " + apply((int)f.xforms.size() - 1) + " }}\n";
        };
        disp_func = functions +
                    "template <bool first_run>\n__device__ __forceinline__ vec4 dispatch_a(vec3 v, int xform, rfk_rng& rs, const float4 rfk_A){\n" +
                    walk(one, "return vec4(v.xy, v.z, 0.0f);") + "}\n" +
                    "template <bool first_run>\n__device__ __forceinline__ void dispatch2_a(vec3 v0, vec3 v1, int xform, rfk_rng& rs0, rfk_rng& rs1, const float4 rfk_A, vec4& o0, vec4& o1){\n" +
                    walk(two, "o0 = vec4(v0.xy, v0.z, 0.0f); o1 = vec4(v1.xy, v1.z, 0.0f); return;");
    }
    if (f.xforms.empty() && d == dialect::glsl) disp_func += "default: { return vec4(v.xy, v.z, 0.0); }}\n";

    std::string xid_func = xform_select_text(buf_map, d == dialect::cuda);

    disp_func = replace_all(disp_func, "\n", "\n\t");
    disp_func += "\n}";
    if (d == dialect::cuda) {
        // Parameter slots that are the same for every temporal sample — all but the four rotated affine coefficients of each
        // xform (animate.tpl.glsl:42-49) — are read from constant memory: `fp[N]` becomes `rfk_cfp[N]`, which the compiler
        // folds into the arithmetic instruction as a constant-bank operand instead of a shared-memory load. The host uploads
        // the array at warmup (it holds the same binary32 values as fp_inflated).
        // The four rotated coefficients (a, b, c, d: consecutive slots) become RFK_AFF(xform, x|y|z|w): a component of the
        // float4 the kernel loads for the picked xform with ONE 128-bit shared-memory read before it enters the switch
        // (rfk_aff[] in device_prelude.cuh, staged per CTA from its temporal sample's row by rfk_stage_params()).
        // rfk_affine_slot[1 + i] is the slot of xform i's coefficient a; [0] that of the final xform (-1: none).
        std::map<int, std::pair<int, int>> per_sample;  // slot -> (xform index, component)
        auto mark = [&](const xform_slots& m, int index) { for (int a = 0; a < 4; a++) per_sample[m.affine[a]] = {index, a}; };
        for (std::size_t i = 0; i < buf_map.xforms.size(); i++) mark(buf_map.xforms[i], (int)i);
        if (buf_map.final_xform) mark(*buf_map.final_xform, -1);
        std::string slots = "__constant__ int rfk_affine_slot[" + std::to_string(buf_map.xforms.size() + 1) + "] = {" +
                            std::to_string(buf_map.final_xform ? buf_map.final_xform->affine[0] : -1);
        for (auto& m : buf_map.xforms) slots += ", " + std::to_string(m.affine[0]);
        slots += "};\n";
        std::string body = xid_func + "\n" + disp_func + "\n", out;
        out.reserve(body.size() + 1024);
        for (std::size_t i = 0; i < body.size();) {
            if (body.compare(i, 3, "fp[") == 0 && (i == 0 || !ident_char(body[i - 1]))) {
                std::size_t j = i + 3, k = j;
                while (k < body.size() && body[k] >= '0' && body[k] <= '9') k++;
                if (k > j && k < body.size() && body[k] == ']') {
                    const auto owner = per_sample.find(std::stoi(body.substr(j, k - j)));
                    if (owner == per_sample.end()) {
                        out += "rfk_cfp[";
                        i = j;
                    } else {
                        out += "RFK_AFF(" + std::to_string(owner->second.first) + ", " + std::string(1, "xyzw"[owner->second.second]) + ")";
                        i = k + 1;
                    }
                    continue;
                }
            }
            out += body[i++];
        }
        // A division whose divisor is one of those slots, `e / rfk_cfp[N]`, multiplies by the slot's reciprocal instead:
        // the host stores 1 / fp[N] (IEEE) in the upper half of the array, rfk_cfp[size + N], so the MUFU.RCP the kernel
        // would issue for a warp-uniform value on every iteration is gone. `/` and `*` associate alike, so the rewrite
        // is local to the token pair. Math mode 0 (IEEE division) keeps the division (RFK_DIVC in device_prelude.cuh).
        const int size = std::max(1, buf_map.size);
        std::string folded;
        folded.reserve(out.size() + 256);
        for (std::size_t i = 0; i < out.size();) {
            if (out[i] == '/' && i + 1 < out.size() && out[i + 1] != '/' && out[i + 1] != '*' && (i == 0 || out[i - 1] != '/')) {
                std::size_t j = i + 1;
                while (j < out.size() && (out[j] == ' ' || out[j] == '\t')) j++;
                if (out.compare(j, 8, "rfk_cfp[") == 0) {
                    std::size_t k = j + 8, e = k;
                    while (e < out.size() && digit_char(out[e])) e++;
                    // the slot must be the whole divisor: not `rfk_cfp[N].x`, not a call or index on it
                    if (e > k && e < out.size() && out[e] == ']' && (e + 1 == out.size() || (out[e + 1] != '.' && out[e + 1] != '[' && out[e + 1] != '('))) {
                        const int slot = std::stoi(out.substr(k, e - k));
                        folded += " RFK_DIVC(" + std::to_string(slot) + ", " + std::to_string(size + slot) + ")";
                        i = e + 1;
                        continue;
                    }
                }
            }
            folded += out[i++];
        }
        // C linkage: the host finds it with cuModuleGetGlobal("rfk_cfp") although the text sits inside namespace rfk_glsl
        return slots + "extern \"C\" { __constant__ float rfk_cfp[" + std::to_string(4 * size) + "]; }\n" + folded;
    }
    return xid_func + disp_func;
}

inline bool digit(char c) { return c >= '0' && c <= '9'; }

}  // namespace

// GLSL floating literals are single precision; unsuffixed they would be doubles in CUDA.
std::string suffix_float_literals(const std::string& s) {
    std::string out;
    out.reserve(s.size() + 64);
    std::size_t i = 0;
    while (i < s.size()) {
        char c = s[i];
        // line comments pass through untouched
        if (c == '/' && i + 1 < s.size() && s[i + 1] == '/') {
            std::size_t nl = s.find('\n', i);
            if (nl == std::string::npos) nl = s.size();
            out.append(s, i, nl - i);
            i = nl;
            continue;
        }
        bool starts_number = (digit(c) || (c == '.' && i + 1 < s.size() && digit(s[i + 1]))) && (i == 0 || !ident_char(s[i - 1])) &&
                             !(i > 0 && s[i - 1] == '.' && c != '.');
        if (!starts_number) {
            out += c;
            i++;
            continue;
        }
        std::size_t j = i;
        bool is_float = false;
        while (j < s.size() && digit(s[j])) j++;
        if (j < s.size() && s[j] == '.') {
            is_float = true;
            j++;
            while (j < s.size() && digit(s[j])) j++;
        }
        if (j < s.size() && (s[j] == 'e' || s[j] == 'E')) {
            std::size_t k = j + 1;
            if (k < s.size() && (s[k] == '+' || s[k] == '-')) k++;
            if (k < s.size() && digit(s[k])) {
                is_float = true;
                while (k < s.size() && digit(s[k])) k++;
                j = k;
            }
        }
        out.append(s, i, j - i);
        if (is_float) {
            if (j < s.size() && (s[j] == 'f' || s[j] == 'F')) { out += 'f'; j++; }
            else if (j + 1 < s.size() && (s[j] == 'l' || s[j] == 'L') && (s[j + 1] == 'f' || s[j + 1] == 'F')) { j += 2; }  // GLSL double suffix
            else out += 'f';
        }
        i = j;
    }
    return out;
}

// `.yx` / `.xy` on an rvalue cannot be a data member in C++: they become calls. `v.xy`
// (the particle position, an lvalue) stays a member of the CUDA-side vec3.
std::string swizzles_to_calls(const std::string& s) {
    std::string out;
    out.reserve(s.size() + 16);
    for (std::size_t i = 0; i < s.size(); i++) {
        if (s[i] == '.' && i + 2 < s.size() && s[i + 1] == 'y' && s[i + 2] == 'x' && (i + 3 == s.size() || !ident_char(s[i + 3])) && i > 0 &&
            (ident_char(s[i - 1]) || s[i - 1] == ')' || s[i - 1] == ']')) {
            out += ".yx()";
            i += 2;
            continue;
        }
        out += s[i];
    }
    return out;
}

// GLSL leaves the evaluation order of operands unspecified; so does C++. A statement
// that draws two or more random numbers gets its draws hoisted into declarations, in
// textual order, so that every implementation consumes the stream identically.
// Statements containing `?` keep their draws in place (a draw under a conditional
// must stay conditional).
std::string sequence_randf(const std::string& body, int& counter) {
    static const std::string call = "randf()";
    std::string out;
    std::size_t start = 0;
    auto flush = [&](std::size_t end) {  // [start, end) is one statement chunk, delimiter included
        std::string chunk = body.substr(start, end - start);
        std::size_t n = 0;
        for (std::size_t p = chunk.find(call); p != std::string::npos; p = chunk.find(call, p + call.size())) n++;
        if (n >= 2 && chunk.find('?') == std::string::npos) {
            std::string decl = "float ";
            std::string rewritten;
            std::size_t p = 0, k = 0;
            for (;;) {
                std::size_t q = chunk.find(call, p);
                if (q == std::string::npos) { rewritten.append(chunk, p, std::string::npos); break; }
                rewritten.append(chunk, p, q - p);
                std::string name = "_rf" + std::to_string(counter++);
                rewritten += name;
                decl += (k++ ? ", " : "") + name + " = randf()";
                p = q + call.size();
            }
            // keep leading whitespace / comment lines in front of the declaration
            std::size_t ins = 0;
            for (;;) {
                while (ins < rewritten.size() && (rewritten[ins] == ' ' || rewritten[ins] == '\t' || rewritten[ins] == '\n')) ins++;
                if (rewritten.compare(ins, 2, "//") == 0) {
                    std::size_t nl = rewritten.find('\n', ins);
                    if (nl == std::string::npos) break;
                    ins = nl + 1;
                } else break;
            }
            out += rewritten.substr(0, ins) + decl + "; " + rewritten.substr(ins);
        } else {
            out += chunk;
        }
        start = end;
    };
    for (std::size_t i = 0; i < body.size(); i++) {
        char c = body[i];
        if (c == '/' && i + 1 < body.size() && body[i + 1] == '/') {  // skip comment text
            std::size_t nl = body.find('\n', i);
            if (nl == std::string::npos) break;
            i = nl;
            continue;
        }
        if (c == ';' || c == '{' || c == '}') flush(i + 1);
    }
    flush(body.size());
    return out;
}

void flame_compiler::load_text(const std::string& yaml_text) {
    auto defs = yaml::parse(yaml_text);

    if (const auto* variations = defs.find("variations"); variations && variations->is_map()) {
        for (auto& [name, def] : variations->entries) {
            const auto* src_n = def.find("src");
            const auto* res_n = def.find("result");
            std::string src = src_n ? src_n->as_string("") : "";
            expand_particle_macros(src);
            std::string result = res_n ? res_n->as_string("") : "";
            expand_particle_macros(result);

            auto& var = vars_[name];
            var = variation_definition{};
            var.source = src;
            var.result = result;

            if (const auto* param = def.find("param"); param && param->is_map())
                for (auto& [pname, unused] : param->entries) {
                    var.param.push_back(pname);
                    param_owners_[pname] = name;
                }
            if (const auto* flags = def.find("flags"); flags)
                for (auto& item : flags->items) var.flags.insert(item.as_string());
        }
    }

    if (const auto* common = defs.find("common"); common && common->is_map()) {
        for (auto& [name, val] : common->entries) {
            std::string src = val.as_string();
            expand_particle_macros(src);
            common_[name] = src;
        }
    }
}

flame_compiler::flame_compiler(const std::string& path) {
    bool ok = false;
    std::string text = read_file(path, &ok);
    if (!ok) throw std::runtime_error("flame_compiler: cannot read " + path);
    load_text(text);
}

flame_compiler flame_compiler::from_text(const std::string& yaml_text) {
    flame_compiler c{empty_tag{}};
    c.load_text(yaml_text);
    return c;
}

void flame_compiler::load_overlay_text(const std::string& yaml_text) { load_text(yaml_text); }

std::string flame_compiler::compile_flame_xforms(const flame& f) const { return emit(f, *this, dialect::glsl); }

std::string flame_compiler::compile_flame_cuda(const flame& f) const { return emit(f, *this, dialect::cuda); }

}  // namespace rfk
