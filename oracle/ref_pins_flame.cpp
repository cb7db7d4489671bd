// TEST INFRASTRUCTURE. The pure affine helpers of the reference's src/flame.hpp:97-128, compiled from the
// reference source where it lies (declaration-only GL/glm stand-ins under oracle/stubs/, nlohmann/json.hpp from the
// image's cudnn_frontend), as pins for the oracle's and the product's restatements. See Makefile.
#include <set>
#include <map>
#include <optional>
#include <memory>
#include <string>
#include <vector>
#include <array>
#include <regex>
namespace jsf32 { struct ctx; }
#include "util.hpp"
#include "flame.hpp"

extern "C" {
void ref_rotate_affine(const float* a, float deg, float* out) {
    flame_xform::affine_t in{a[0], a[1], a[2], a[3], a[4], a[5]};
    auto r = flame::rotate_affine(in, deg);
    for (int i = 0; i < 6; i++) out[i] = r[i];
}
void ref_scale_affine(const float* a, float s, float* out) {
    flame_xform::affine_t in{a[0], a[1], a[2], a[3], a[4], a[5]};
    auto r = flame::scale_affine(in, s);
    for (int i = 0; i < 6; i++) out[i] = r[i];
}
void ref_translate_affine(const float* a, const float* t, float* out) {
    flame_xform::affine_t in{a[0], a[1], a[2], a[3], a[4], a[5]};
    auto r = flame::translate_affine(in, {t[0], t[1]});
    for (int i = 0; i < 6; i++) out[i] = r[i];
}
// src/flame.cpp:289-296 composed from the reference's own helpers
void ref_screen_space_affine(float scale, float rotate, float cx, float cy, unsigned size_y, unsigned long W, unsigned long H, float* out) {
    std::size_t target_dims[2] = {W, H};
    flame_xform::affine_t base{1, 0, 0, 1, 0, 0};
    base = flame::translate_affine(base, {target_dims[0] / 2.0f, target_dims[1] / 2.0f});
    base = flame::scale_affine(base, scale * float(target_dims[1]) / float(size_y));
    base = flame::rotate_affine(base, rotate);
    base = flame::translate_affine(base, {-cx, -cy});
    for (int i = 0; i < 6; i++) out[i] = base[i];
}
}
