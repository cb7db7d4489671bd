// Drives refrakt's own `flame` interface (the calls of src/main.cpp:203-215, :403-412, :490-535, :590-593) with the B200
// binding of flame_b200.cpp behind it, and prints one JSON line. Working directory: a directory holding variations.yaml.
//   binding_demo <genome.flam3> <W> <H> <particles> <temporal samples> <draw passes> [out.png]
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <map>
#include <memory>
#include <optional>
#include <regex>
#include <set>
#include <string>
#include <vector>

#include "refrakt_b200.h"

#include "util.hpp"
#include "flame.hpp"
#include "variation_table.hpp"

namespace b200 { float* device_bins(std::size_t len); rfk_flame* handle_of(const flame* f); }

int main(int argc, char** argv) {
    if (argc < 7) { std::fprintf(stderr, "usage: binding_demo genome W H particles temporal_samples passes [out.png]\n"); return 2; }
    const std::size_t W = std::strtoul(argv[2], nullptr, 10), H = std::strtoul(argv[3], nullptr, 10);
    const std::size_t P = std::strtoul(argv[4], nullptr, 10), TS = std::strtoul(argv[5], nullptr, 10);
    const int passes = std::atoi(argv[6]);

    flame::set_sim_parameters(P, TS, 1024);                      // main.cpp:203
    flame_compiler variations{};                                 // main.cpp:205 (the reference's own table, for the UI)
    auto flame_def = flame::load_flame(argv[1], variations);     // main.cpp:206
    if (!flame_def) { std::printf("{\"loaded\": false}\n"); return 1; }

    flame::bin_t bins{W * H};                                    // main.cpp:214
    bins.zero_out();
    rfk_device_zero(b200::device_bins(W * H), W * H * 16);

    std::size_t binned = 0;
    bool warmed = false;
    if (flame_def->needs_warmup()) {                             // main.cpp:403-409
        flame_def->warmup(16, 1.2f / 60.0f);
        warmed = !flame_def->needs_warmup();
    }
    if (warmed) binned = flame_def->draw_to_bins(bins, W, passes);  // main.cpp:412

    int image_ok = 0;
    if (warmed && argc > 7) {                                    // main.cpp:490-535 + the screenshot of :590-593
        rfk_post_params p;
        rfk_flame_post_params(b200::handle_of(flame_def.get()), &p);
        uint8_t* rgba8 = static_cast<uint8_t*>(rfk_device_alloc(W * H * 4));
        std::vector<uint8_t> pixels(W * H * 4);
        if (rgba8 && rfk_density_tonemap(b200::device_bins(W * H), nullptr, rgba8, W, H, &p) == RFK_OK &&
            rfk_memcpy_to_host(pixels.data(), rgba8, pixels.size()) == RFK_OK && rfk_write_png(argv[7], pixels.data(), W, H) == RFK_OK)
            image_ok = 1;
        rfk_device_free(rgba8);
    }
    std::printf("{\"loaded\": true, \"xforms\": %zu, \"has_final\": %d, \"scale\": %.9g, \"gamma\": %.9g, \"variation_julian_known\": %d, "
                "\"warmed\": %d, \"binned\": %zu, \"image\": %d, \"error\": \"%s\"}\n",
                flame_def->xforms.size(), flame_def->final_xform ? 1 : 0, flame_def->scale, flame_def->gamma, variations.is_variation("julian") ? 1 : 0,
                warmed ? 1 : 0, binned, image_ok, warmed ? "" : rfk_last_error());
    return 0;
}
