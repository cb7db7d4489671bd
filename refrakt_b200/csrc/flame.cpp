// Genome model, flam3 parser and parameter-buffer layout (host only, no CUDA here).
// Restates src/flame.cpp:33-103 and :160-226 of the reference.
#include "flame.hpp"

#include <cmath>
#include <cstdlib>
#include <sstream>
#include <stdexcept>

#include "textutil.hpp"
#include "variation_table.hpp"
#include "xml_lite.hpp"

namespace rfk {

namespace {
thread_local std::string g_last_error;

// util.hpp:12-25: up to N whitespace-separated tokens, converted one by one; missing ones stay 0
template <typename T, std::size_t N, typename Conv>
std::array<T, N> parse_strings(const std::string* in, Conv&& cv) {
    std::array<T, N> ret{};
    if (!in) return ret;
    auto tokens = split_ws(*in);
    for (std::size_t i = 0; i < N && i < tokens.size(); i++) ret[i] = cv(tokens[i]);
    return ret;
}

void json_xform(std::ostringstream& o, const xform_slots& x, const char* indent) {
    auto arr = [&](const std::array<int, 6>& a) {
        o << "[";
        for (int i = 0; i < 6; i++) o << (i ? ", " : "") << a[i];
        o << "]";
    };
    auto obj = [&](const std::map<std::string, int>& m) {
        o << "{";
        bool first = true;
        for (auto& [k, v] : m) { o << (first ? "" : ", ") << "\"" << k << "\": " << v; first = false; }
        o << "}";
    };
    // keys in nlohmann::json's (alphabetical) order
    o << "{\n" << indent << " \"affine\": "; arr(x.affine);
    o << ",\n" << indent << " \"color\": " << x.color;
    o << ",\n" << indent << " \"color_speed\": " << x.color_speed;
    o << ",\n" << indent << " \"meta\": {\"end\": " << x.end << ", \"size\": " << x.size << ", \"start\": " << x.start << "}";
    o << ",\n" << indent << " \"opacity\": " << x.opacity;
    o << ",\n" << indent << " \"param\": "; obj(x.param);
    if (x.has_post) { o << ",\n" << indent << " \"post\": "; arr(x.post); }
    o << ",\n" << indent << " \"rotation_frequency\": " << x.rotation_frequency;
    o << ",\n" << indent << " \"variations\": "; obj(x.variations);
    o << ",\n" << indent << " \"weight\": " << x.weight;
    o << "\n" << indent << "}";
}
}  // namespace

void set_last_error(const std::string& msg) { g_last_error = msg; }
const std::string& flame::last_error() { return g_last_error; }

std::string buffer_map_t::dump_json() const {
    std::ostringstream o;
    o << "{\n";
    if (final_xform) { o << " \"final_xform\": "; json_xform(o, *final_xform, " "); o << ",\n"; }
    o << " \"size\": " << size << ",\n \"xforms\": [";
    for (std::size_t i = 0; i < xforms.size(); i++) {
        o << (i ? ",\n  " : "\n  ");
        json_xform(o, xforms[i], "  ");
    }
    o << "\n ]\n}";
    return o.str();
}

void flame::make_shader_buffer_map() {
    buffer_map_t map;
    int counter = 0;

    auto make_xform_map = [&counter](const flame_xform& xform) {
        xform_slots m;
        int start = counter;
        m.start = start;
        m.weight = counter++;
        for (int a = 0; a < 6; a++) m.affine[a] = counter++;
        if (xform.post) {
            m.has_post = true;
            for (int a = 0; a < 6; a++) m.post[a] = counter++;
        }
        for (auto& [k, v] : xform.variations) m.variations[k] = counter++;
        for (auto& [k, v] : xform.var_param) m.param[k] = counter++;
        m.color = counter++;
        m.color_speed = counter++;
        m.opacity = counter++;
        m.rotation_frequency = counter++;
        m.end = counter - 1;
        m.size = counter - start;
        return m;
    };

    for (const auto& xform : xforms) map.xforms.push_back(make_xform_map(xform));
    if (final_xform) map.final_xform = make_xform_map(final_xform.value());
    map.size = counter++;
    buffer_map_ = map;
}

std::array<float, flame::PARAM_BUFFER> flame::copy_flame_data_to_buffer() const {
    std::array<float, PARAM_BUFFER> buf{};

    float normal_weight = 0.0;
    for (auto& x : xforms) normal_weight += x.weight;

    auto push_xform = [&](const flame_xform& xform, const xform_slots& xmap) {
        buf[xmap.weight] = xform.weight / normal_weight;
        for (int a = 0; a < 6; a++) buf[xmap.affine[a]] = xform.affine[a];
        for (auto& [n, w] : xform.variations) buf[xmap.variations.at(n)] = w;
        for (auto& [n, v] : xform.var_param) buf[xmap.param.at(n)] = v;
        if (xform.post) for (int a = 0; a < 6; a++) buf[xmap.post[a]] = xform.post.value()[a];
        buf[xmap.color] = xform.color;
        buf[xmap.opacity] = xform.opacity;
        buf[xmap.color_speed] = xform.color_speed;
        buf[xmap.rotation_frequency] = xform.rotation_frequency;
    };

    for (std::size_t i = 0; i < xforms.size(); i++) push_xform(xforms[i], buffer_map_.xforms.at(i));
    if (final_xform) push_xform(final_xform.value(), buffer_map_.final_xform.value());
    return buf;
}

flame_xform::affine_t flame::rotate_affine(const flame_xform::affine_t& a, float deg) {
    float rad = 0.01745329251f * deg;
    float sino = sinf(rad);
    float coso = cosf(rad);

    flame_xform::affine_t ret = a;
    ret[0] = a[0] * coso + a[2] * sino;
    ret[1] = a[1] * coso + a[3] * sino;
    ret[2] = a[2] * coso - a[0] * sino;
    ret[3] = a[3] * coso - a[1] * sino;
    return ret;
}

flame_xform::affine_t flame::scale_affine(const flame_xform::affine_t& a, float scale) {
    flame_xform::affine_t ret = a;
    ret[0] = a[0] * scale;
    ret[1] = a[1] * scale;
    ret[2] = a[2] * scale;
    ret[3] = a[3] * scale;
    return ret;
}

flame_xform::affine_t flame::translate_affine(const flame_xform::affine_t& a, const std::array<float, 2>& t) {
    flame_xform::affine_t ret = a;
    ret[4] = ret[0] * t[0] + ret[2] * t[1] + a[4];
    ret[5] = ret[1] * t[0] + ret[3] * t[1] + a[5];
    return ret;
}

flame_xform::affine_t flame::screen_space_affine(std::size_t bins_width, std::size_t bins_height) const {
    std::size_t target_dims[2] = {bins_width, bins_height};
    flame_xform::affine_t base{1, 0, 0, 1, 0, 0};
    base = translate_affine(base, {target_dims[0] / 2.0f, target_dims[1] / 2.0f});
    base = scale_affine(base, scale * float(target_dims[1]) / float(size[1]));
    base = rotate_affine(base, rotate);
    base = translate_affine(base, {-center[0], -center[1]});
    return base;
}

std::unique_ptr<flame> flame::load_flame(const std::string& path, const flame_compiler& vt) {
    bool ok = false;
    std::string text = read_file(path, &ok);
    if (!ok) {
        set_last_error("cannot read flame file " + path);
        return nullptr;
    }
    return load_flame_string(text, path, vt);
}

std::unique_ptr<flame> flame::load_flame_string(const std::string& xml_text, const std::string& origin, const flame_compiler& vt) {
    xml::element root;
    try {
        root = xml::parse(xml_text);
    } catch (const std::exception& e) {
        set_last_error(std::string(e.what()) + " in flame " + origin);
        return nullptr;
    }
    // the reference takes the document's <flame> child; a <flames> wrapper is also accepted
    const xml::element* flame_node = root.name == "flame" ? &root : root.child("flame");
    if (!flame_node) {
        set_last_error("no <flame> element in " + origin);
        return nullptr;
    }

    std::unique_ptr<flame> f{new flame{}};
    std::string errors;
    try {
        f->center = parse_strings<float, 2>(flame_node->attribute("center"), [](auto& v) { return std::stof(v); });
        f->scale = xml::as_float(flame_node->attribute("scale"));
        f->rotate = xml::as_float(flame_node->attribute("rotate"));
        f->estimator_curve = xml::as_float(flame_node->attribute("estimator_curve"));
        f->estimator_min = xml::as_int(flame_node->attribute("estimator_min"));  // sic: flam3 writes estimator_minimum
        f->estimator_radius = xml::as_int(flame_node->attribute("estimator_radius"));
        f->brightness = xml::as_float(flame_node->attribute("brightness"));
        f->gamma = xml::as_float(flame_node->attribute("gamma"));
        f->vibrancy = xml::as_float(flame_node->attribute("vibrancy"));
        f->size = parse_strings<unsigned int, 2>(flame_node->attribute("size"), [](auto& v) { return (unsigned int)std::stoi(v); });

        for (const auto& node : flame_node->children) {
            const std::string& node_name = node.name;
            if (node_name == "xform" || node_name == "finalxform") {
                flame_xform xform{};
                for (const auto& [name, value] : node.attributes) {
                    if (name == "weight") xform.weight = xml::as_float(&value);
                    else if (name == "color") xform.color = xml::as_float(&value);
                    else if (name == "color_speed") xform.color_speed = xml::as_float(&value);
                    // sic: the node is called "finalxform", so a final xform with animate > 0 rotates too
                    else if (name == "animate") xform.rotation_frequency = (xml::as_float(&value) > 0 && node_name != "final_xform") ? 1.0f : 0.0f;
                    else if (name == "opacity") xform.opacity = xml::as_float(&value);
                    else if (vt.is_param(name)) xform.var_param[name] = xml::as_float(&value);
                    else if (vt.is_variation(name)) xform.variations[name] = xml::as_float(&value);
                    else if (name == "coefs") xform.affine = parse_strings<float, 6>(&value, [](auto& v) { return (float)std::stod(v); });
                    else if (name == "post") xform.post = parse_strings<float, 6>(&value, [](auto& v) { return (float)std::stod(v); });
                    else errors += "Unknown attribute " + name + " in flame " + origin + "\n";
                }
                // src/flame.cpp:199-210 (commented out there): every attribute of a <motion> child other than its frequency and
                // function names an animated field of this xform
                for (const auto& motion : node.children) {
                    if (motion.name != "motion") continue;
                    motion_info m{};
                    m.freq = xml::as_float(motion.attribute("motion_frequency"));
                    if (const std::string* fn = motion.attribute("motion_function")) m.function = *fn;
                    for (const auto& [name, value] : motion.attributes) {
                        if (name == "motion_frequency" || name == "motion_function") continue;
                        m.amplitude = xml::as_float(&value);
                        xform.motion[name] = m;
                    }
                }
                if (node_name == "finalxform") f->final_xform = xform;
                else f->xforms.push_back(xform);
            } else if (node_name == "color") {
                auto idx = xml::as_ullong(node.attribute("index"));
                if (idx >= f->palette.size()) { errors += "palette index " + std::to_string(idx) + " out of range in flame " + origin + "\n"; continue; }
                f->palette[idx] = parse_strings<float, 4>(node.attribute("rgb"), [](auto& v) { return std::stoi(v) / 256.0f; });
                f->palette[idx][3] = 1.0f;
            }
        }
    } catch (const std::exception& e) {  // std::stof / stoi / stod on a non-numeric token
        set_last_error(std::string("bad numeric value (") + e.what() + ") in flame " + origin);
        return nullptr;
    }

    if (!errors.empty()) {
        while (!errors.empty() && errors.back() == '\n') errors.pop_back();
        set_last_error(errors);
        return nullptr;
    }
    if (f->xforms.empty()) {
        set_last_error("flame " + origin + " has no xforms");
        return nullptr;
    }
    if (!f->do_common_init(vt)) return nullptr;
    return f;
}

// flam3's motion functions of the phase x = frequency * time (flam3.c, motion_funcs): sin: sin(2 pi x); triangle: a
// triangle wave through 0 at x = 0 with peaks +-1 at x = 1/4, 3/4; hill: (1 - cos(2 pi x)) / 2.
float flame::motion_function(const std::string& name, float x) {
    const double pi = 3.14159265358979323846;
    if (name == "sin") return (float)std::sin(2.0 * pi * (double)x);
    if (name == "hill") return (float)((1.0 - std::cos(2.0 * pi * (double)x)) * 0.5);
    if (name == "triangle") {
        double fr = std::fmod((double)x, 1.0);
        if (fr < 0.0) fr += 1.0;
        if (fr <= 0.25) return (float)(4.0 * fr);
        if (fr <= 0.75) return (float)(-4.0 * fr + 2.0);
        return (float)(4.0 * fr - 4.0);
    }
    return 0.0f;
}

int flame::apply_motion(float time) {
    if (motion_base_.empty()) {
        motion_base_ = xforms;
        if (final_xform) motion_base_.push_back(*final_xform);
    }
    int written = 0;
    auto apply = [&](flame_xform& x, const flame_xform& base) {
        for (const auto& [name, m] : base.motion) {
            const float delta = m.amplitude * motion_function(m.function, m.freq * time);
            auto set = [&](float& field, float from) { field = from + delta; written++; };
            if (name == "weight") set(x.weight, base.weight);
            else if (name == "color") set(x.color, base.color);
            else if (name == "color_speed") set(x.color_speed, base.color_speed);
            else if (name == "opacity") set(x.opacity, base.opacity);
            else if (auto v = x.variations.find(name); v != x.variations.end()) set(v->second, base.variations.at(name));
            else if (auto q = x.var_param.find(name); q != x.var_param.end()) set(q->second, base.var_param.at(name));
            // anything else (a variation the xform does not use: the structure is fixed after load) is ignored
        }
    };
    for (std::size_t i = 0; i < xforms.size() && i < motion_base_.size(); i++) apply(xforms[i], motion_base_[i]);
    if (final_xform && motion_base_.size() == xforms.size() + 1) apply(*final_xform, motion_base_.back());
    if (written) needs_update_ = true;
    return written;
}

}  // namespace rfk
