// rfk_render — headless command-line renderer over the C ABI (the reference only has an
// interactive screenshot button, src/main.cpp:590-593). Renders stills or an animation (per-frame rotation of
// src/main.cpp:383-395: 18 deg/s * rotation_frequency) to PNG files.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/refrakt_b200.h"

static void usage() {
    std::fprintf(stderr,
                 "usage: rfk_render --genome FILE.flam3 --variations variations.yaml [--overlay FILE.yaml] --out OUT.png\n"
                 "  [--width 1280] [--height 720] [--quality 2000 (samples/pixel)] [--passes 128] [--warmup 16]\n"
                 "  [--particles 2097152] [--temporal-samples 512] [--tss-width 0.02] [--seed 0] [--device 0]\n"
                 "  [--frames N --fps 60] (OUT.png takes a %%d / %%04d frame number) [--deterministic] [--math-mode 0|1|2]\n"
                 "  [--supersample 1] [--filter 1.0] (histogram at supersample x the image size, spatial filter radius in pixels;\n"
                 "   --quality is samples per histogram bin)\n");
}

int main(int argc, char** argv) {
    std::string genome, variations, overlay, out;
    unsigned width = 1280, height = 720, quality = 2000, passes = 128, warmup = 16, frames = 1;
    size_t particles = 2048 * 1024, ts = 512;
    float tss_width = 1.2f / 60.0f, fps = 60.0f;
    unsigned long long seed = 0;
    int device = 0, deterministic = 0, math_mode = -1;
    unsigned supersample = 1;
    float filter_radius = 1.0f;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { usage(); std::exit(2); } return argv[++i]; };
        if (a == "--genome") genome = next();
        else if (a == "--variations") variations = next();
        else if (a == "--overlay") overlay = next();
        else if (a == "--out") out = next();
        else if (a == "--width") width = std::strtoul(next(), nullptr, 10);
        else if (a == "--height") height = std::strtoul(next(), nullptr, 10);
        else if (a == "--quality") quality = std::strtoul(next(), nullptr, 10);
        else if (a == "--passes") passes = std::strtoul(next(), nullptr, 10);
        else if (a == "--warmup") warmup = std::strtoul(next(), nullptr, 10);
        else if (a == "--particles") particles = std::strtoull(next(), nullptr, 10);
        else if (a == "--temporal-samples") ts = std::strtoull(next(), nullptr, 10);
        else if (a == "--tss-width") tss_width = std::strtof(next(), nullptr);
        else if (a == "--seed") seed = std::strtoull(next(), nullptr, 10);
        else if (a == "--device") device = std::atoi(next());
        else if (a == "--frames") frames = std::strtoul(next(), nullptr, 10);
        else if (a == "--fps") fps = std::strtof(next(), nullptr);
        else if (a == "--supersample") supersample = std::strtoul(next(), nullptr, 10);
        else if (a == "--filter") filter_radius = std::strtof(next(), nullptr);
        else if (a == "--deterministic") deterministic = 1;
        else if (a == "--math-mode") math_mode = std::atoi(next());
        else { usage(); return 2; }
    }
    if (genome.empty() || variations.empty() || out.empty() || !width || !height || !frames) { usage(); return 2; }

    auto die = [](const char* what) { std::fprintf(stderr, "rfk_render: %s: %s\n", what, rfk_last_error()); std::exit(1); };
    if (rfk_set_device(device) != RFK_OK) die("set_device");
    rfk_compiler* c = rfk_compiler_create(variations.c_str());
    if (!c) die("variations");
    if (!overlay.empty() && rfk_compiler_load_overlay(c, overlay.c_str()) != RFK_OK) die("overlay");
    rfk_flame* f = rfk_flame_load(genome.c_str(), c);
    if (!f) die("load_flame");
    if (deterministic || math_mode >= 0) {
        rfk_kernel_options o;
        rfk_flame_get_options(f, &o);
        if (deterministic) o.deterministic = 1;
        if (math_mode >= 0) o.math_mode = math_mode;
        if (rfk_flame_set_options(f, &o) != RFK_OK) die("set_options");
    }
    if (rfk_set_sim_parameters(particles, ts, 1024, seed) != RFK_OK) die("set_sim_parameters");

    std::vector<uint8_t> pixels((size_t)width * height * 4);
    rfk_frame_request req{};
    req.width = width; req.height = height; req.warmup_passes = warmup; req.drawing_passes = passes; req.tss_width = tss_width;
    req.target_binned = (uint64_t)quality * width * height * supersample * supersample; req.max_draw_calls = 0; req.scale_constant_exp = 4.0f;
    req.supersample = supersample; req.filter_radius = filter_radius;
    for (unsigned frame = 0; frame < frames; frame++) {
        if (frame) rfk_flame_rotate_xforms(f, 18.0f / fps);  // DEGREES_PER_SECOND * dt, main.cpp:224
        rfk_frame_stats st{};
        auto t0 = std::chrono::steady_clock::now();
        if (rfk_render_frame(f, &req, pixels.data(), nullptr, &st) != RFK_OK) die("render_frame");
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        char name[4096];
        std::snprintf(name, sizeof name, out.c_str(), frame);
        if (rfk_write_png(name, pixels.data(), width, height) != RFK_OK) die("write_png");
        std::printf("{\"frame\": %u, \"file\": \"%s\", \"iterations\": %llu, \"binned\": %llu, \"draw_calls\": %u, \"ms\": %.3f, \"ms_draw\": %.3f, \"ms_post\": %.3f, \"giter_per_s\": %.2f}\n",
                    frame, name, (unsigned long long)st.iterations, (unsigned long long)st.binned, st.draw_calls, ms, st.ms_draw, st.ms_post,
                    st.iterations / (ms * 1e6));
    }
    rfk_flame_destroy(f);
    rfk_compiler_destroy(c);
    return 0;
}
