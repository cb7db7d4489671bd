// rfk_render — headless command-line renderer over the C ABI (the reference only has an
// interactive screenshot button, src/main.cpp:590-593). Renders stills or an animation (per-frame rotation of
// src/main.cpp:383-395: 18 deg/s * rotation_frequency) to PNG files.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/refrakt_b200.h"

static void usage() {
    std::fprintf(stderr,
                 "usage: rfk_render --genome FILE.flam3 --variations variations.yaml [--overlay FILE.yaml] --out OUT.png\n"
                 "  [--width 1280] [--height 720] [--quality 2000 (samples/pixel)] [--passes 128] [--warmup 16]\n"
                 "  [--particles 2097152] [--temporal-samples 512] [--tss-width 0.02] [--seed 0] [--device 0]\n"
                 "  [--frames N --fps 60] (OUT.png takes a %%d / %%04d frame number) [--deterministic] [--math-mode 0|1|2]\n"
                 "  [--supersample 1] [--filter 1.0] (histogram at supersample x the image size, spatial filter radius in pixels;\n"
                 "   --quality is samples per histogram bin)\n"
                 "  [--exr OUT.exr] also write the float frame (OpenEXR, uncompressed 32-bit float RGBA; same %%d rule as --out)\n"
                 "  [--motion] evaluate the genome's <motion> elements at every frame's time (frame / fps)\n"
                 "  [--world N --rank R --comm-file PATH] one process per GPU (--device defaults to R): the N processes render every\n"
                 "   frame together (particle streams sharded, histograms reduce-scattered, rank 0 writes the PNG); rank 0 leaves the\n"
                 "   NCCL id in PATH, the others wait for it. [--frame-parallel]: instead, rank R renders the frames R, R+N, ... alone\n");
}

// The frame number goes into OUT.png through at most one %d / %0Nd conversion; anything else after a '%' (a stray %s or %n
// from the command line) would be undefined behaviour in printf, so the name is assembled by hand.
static bool frame_name(const std::string& pattern, unsigned frame, std::string& out) {
    out.clear();
    bool used = false;
    for (size_t i = 0; i < pattern.size(); i++) {
        if (pattern[i] != '%') { out += pattern[i]; continue; }
        if (i + 1 < pattern.size() && pattern[i + 1] == '%') { out += '%'; i++; continue; }
        size_t j = i + 1;
        bool zero = j < pattern.size() && pattern[j] == '0';
        unsigned width = 0;
        while (j < pattern.size() && pattern[j] >= '0' && pattern[j] <= '9') { width = width * 10 + (pattern[j] - '0'); j++; }
        if (j >= pattern.size() || (pattern[j] != 'd' && pattern[j] != 'u') || used || width > 32) return false;
        std::string digits = std::to_string(frame);
        if (digits.size() < width) digits.insert(0, width - digits.size(), zero ? '0' : ' ');
        out += digits;
        used = true;
        i = j;
    }
    return true;
}

static bool read_or_write_id(const std::string& path, int rank, uint8_t id[RFK_COMM_ID_BYTES]) {
    if (rank == 0) {
        if (rfk_comm_unique_id(id) != RFK_OK) return false;
        const std::string tmp = path + ".tmp";
        FILE* fh = std::fopen(tmp.c_str(), "wb");
        if (!fh) return false;
        const bool ok = std::fwrite(id, 1, RFK_COMM_ID_BYTES, fh) == RFK_COMM_ID_BYTES;
        std::fclose(fh);
        return ok && std::rename(tmp.c_str(), path.c_str()) == 0;
    }
    for (int tries = 0; tries < 6000; tries++) {  // up to ten minutes
        if (FILE* fh = std::fopen(path.c_str(), "rb")) {
            const size_t got = std::fread(id, 1, RFK_COMM_ID_BYTES, fh);
            std::fclose(fh);
            if (got == RFK_COMM_ID_BYTES) return true;
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(100));
    }
    return false;
}

int main(int argc, char** argv) {
    std::string genome, variations, overlay, out;
    unsigned width = 1280, height = 720, quality = 2000, passes = 128, warmup = 16, frames = 1;
    size_t particles = 2048 * 1024, ts = 512;
    float tss_width = 1.2f / 60.0f, fps = 60.0f;
    unsigned long long seed = 0;
    int device = -1, deterministic = 0, math_mode = -1, world = 1, rank = 0, frame_parallel = 0, motion = 0;
    std::string comm_file, exr;
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
        else if (a == "--world") world = std::atoi(next());
        else if (a == "--rank") rank = std::atoi(next());
        else if (a == "--comm-file") comm_file = next();
        else if (a == "--frame-parallel") frame_parallel = 1;
        else if (a == "--motion") motion = 1;
        else if (a == "--exr") exr = next();
        else if (a == "--deterministic") deterministic = 1;
        else if (a == "--math-mode") math_mode = std::atoi(next());
        else { usage(); return 2; }
    }
    if (genome.empty() || variations.empty() || out.empty() || !width || !height || !frames) { usage(); return 2; }
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && !frame_parallel && comm_file.empty())) { usage(); return 2; }
    if (device < 0) device = rank;
    std::string probe;
    if (!frame_name(out, 0, probe)) { std::fprintf(stderr, "rfk_render: --out takes at most one %%d / %%0Nd conversion\n"); return 2; }
    const bool sharded = world > 1 && !frame_parallel;

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
    // rank g seeds the particle slots [seed + g * P, seed + (g + 1) * P): disjoint JSF32 streams per GPU
    if (rfk_set_sim_parameters(particles, ts, 1024, seed + (sharded ? (unsigned long long)rank * particles : 0ull)) != RFK_OK) die("set_sim_parameters");
    if (sharded) {
        uint8_t id[RFK_COMM_ID_BYTES];
        if (!read_or_write_id(comm_file, rank, id)) { std::fprintf(stderr, "rfk_render: cannot exchange the NCCL id through %s\n", comm_file.c_str()); return 1; }
        if (rfk_comm_init(id, rank, world) != RFK_OK) die("comm_init");
    }

    std::vector<uint8_t> pixels((size_t)width * height * 4);
    std::vector<float> image;
    rfk_frame_request req{};
    req.width = width; req.height = height; req.warmup_passes = warmup; req.drawing_passes = passes; req.tss_width = tss_width;
    req.target_binned = (uint64_t)quality * width * height * supersample * supersample; req.max_draw_calls = 0; req.scale_constant_exp = 4.0f;
    req.supersample = supersample; req.filter_radius = filter_radius;
    unsigned rotated_to = 0;
    for (unsigned frame = 0; frame < frames; frame++) {
        if (frame_parallel && (int)(frame % (unsigned)world) != rank) continue;  // frames round-robin over the ranks, no exchange
        for (; rotated_to < frame; rotated_to++) rfk_flame_rotate_xforms(f, 18.0f / fps);  // DEGREES_PER_SECOND * dt, main.cpp:224; frame by frame, whoever renders it
        if (motion) rfk_flame_apply_motion(f, (float)frame / fps);  // the genome's <motion> elements at this frame's time
        std::string name;
        frame_name(out, frame, name);
        auto t0 = std::chrono::steady_clock::now();
        if (sharded) {
            rfk_sharded_request sr{};
            sr.frame = req; sr.want_rgba8 = 1; sr.want_image = 0;
            rfk_sharded_stats st{};
            if (rfk_render_frame_sharded(f, &sr, rank == 0 ? pixels.data() : nullptr, nullptr, &st) != RFK_OK) die("render_frame_sharded");
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (rank == 0) {
                if (rfk_write_png(name.c_str(), pixels.data(), width, height) != RFK_OK) die("write_png");
                std::printf("{\"frame\": %u, \"file\": \"%s\", \"n_gpus\": %d, \"p2p\": %u, \"iterations\": %llu, \"binned\": %llu, \"passes_per_rank\": %llu, \"ms\": %.3f, \"ms_draw\": %.3f, \"ms_reduce\": %.3f, \"ms_post\": %.3f, \"giter_per_s\": %.2f}\n",
                            frame, name.c_str(), world, st.p2p, (unsigned long long)st.iterations_global, (unsigned long long)st.binned_global, (unsigned long long)st.passes, ms,
                            st.ms_draw, st.ms_reduce, st.ms_post, st.iterations_global / (ms * 1e6));
            }
            continue;
        }
        rfk_frame_stats st{};
        if (!exr.empty() && image.empty()) image.resize((size_t)width * height * 4);
        if (rfk_render_frame(f, &req, pixels.data(), exr.empty() ? nullptr : image.data(), &st) != RFK_OK) die("render_frame");
        if (!exr.empty()) {
            std::string exr_name;
            if (!frame_name(exr, frame, exr_name) || rfk_write_exr(exr_name.c_str(), image.data(), width, height) != RFK_OK) die("write_exr");
        }
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (rfk_write_png(name.c_str(), pixels.data(), width, height) != RFK_OK) die("write_png");
        std::printf("{\"frame\": %u, \"file\": \"%s\", \"iterations\": %llu, \"binned\": %llu, \"draw_calls\": %u, \"ms\": %.3f, \"ms_draw\": %.3f, \"ms_post\": %.3f, \"giter_per_s\": %.2f}\n",
                    frame, name.c_str(), (unsigned long long)st.iterations, (unsigned long long)st.binned, st.draw_calls, ms, st.ms_draw, st.ms_post,
                    st.iterations / (ms * 1e6));
    }
    if (sharded) rfk_comm_destroy();
    rfk_flame_destroy(f);
    rfk_compiler_destroy(c);
    return 0;
}
