// RGBA8 PNG writer (zlib deflate). Stands in for stb_image_write, which the reference uses for its
// screenshot button (src/main.cpp:590-593: stbi_write_png("screenshot.png", W, H, 4, pixels, 0)).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace rfk {

// rows top to bottom, 4 bytes per pixel, row stride = width * 4. Throws std::runtime_error on I/O failure.
void write_png_rgba8(const std::string& path, const std::uint8_t* rgba, std::size_t width, std::size_t height, int compression_level = 3);

// OpenEXR 2, scanline, uncompressed, four 32-bit FLOAT channels A B G R: the float frame (the reference keeps its frame as
// RGBA32F textures, main.cpp:218-219, and only ever saves 8 bits of it). rows top to bottom, 4 floats per pixel.
void write_exr_rgba32f(const std::string& path, const float* rgba, std::size_t width, std::size_t height);

}  // namespace rfk
