// TEST INFRASTRUCTURE — see softgl.hpp. Linked only into oracle/_ref/libref_host.so.
#include "softgl.hpp"

#include <dlfcn.h>
#include <sys/stat.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>

namespace softgl {
namespace {

const uint COMPUTE_SHADER = 0x91B9, VERTEX_SHADER = 0x8B31, FRAGMENT_SHADER = 0x8B30;

struct shader_object { uint type = 0; std::string source; };

struct stage_library {  // one transliterated shader (glsl_to_cpp.py + shader_rt_post.hpp)
    void* handle = nullptr;
    int (*bind)(int, void*) = nullptr;
    int (*location)(const char*) = nullptr;
    int (*set)(int, const void*, std::size_t) = nullptr;
    void (*dispatch)(unsigned, unsigned, unsigned) = nullptr;
    std::size_t (*private_bytes)() = nullptr;
    int (*varyings)(const char**, std::size_t*, std::size_t*, int*, int) = nullptr;
    void (*run_vertex)(int, void*, float*, float*) = nullptr;
    int (*run_fragment)(float, float, void*) = nullptr;
};

struct program_object {
    std::vector<uint> shaders;
    stage_library compute, vertex, fragment;
    std::vector<std::string> uniform_names;  // location -> name
    bool linked = false;
    std::string log;
};

// never destroyed: the reference's statics (flame::swap_buffer_, ...) release their GL objects during process exit
auto& g_buffers = *new std::map<uint, std::unique_ptr<std::vector<unsigned char>>>();
auto& g_shaders = *new std::map<uint, shader_object>();
auto& g_programs = *new std::map<uint, program_object>();
auto& g_bindings = *new std::map<uint, uint>();
uint g_next_name = 1, g_current = 0;
float* g_target = nullptr;
int g_target_w = 0, g_target_h = 0;

auto& g_sources = *new std::vector<std::string>();
auto& g_uniform_log = *new std::vector<uniform_record>();
auto& g_dispatch_log = *new std::vector<dispatch_record>();

std::string fnv1a_hex(const std::string& s) {
    std::uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    char b[32];
    std::snprintf(b, sizeof b, "%016llx", (unsigned long long)h);
    return b;
}

bool file_exists(const std::string& p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }

// builds (once) and loads the library of one shader source
bool load_stage(const std::string& source, uint type, stage_library& lib, std::string& log) {
    const char* dir_env = std::getenv("RFK_SOFTGL_DIR");
    const char* tool_env = std::getenv("RFK_SOFTGL_TRANSLATOR");
    if (!dir_env || !tool_env) { log = "soft GL: RFK_SOFTGL_DIR / RFK_SOFTGL_TRANSLATOR are not set"; return false; }
    const char* stage = type == COMPUTE_SHADER ? "compute" : type == VERTEX_SHADER ? "vertex" : "fragment";
    std::string base = std::string(dir_env) + "/" + stage + "_" + fnv1a_hex(source);
    if (!file_exists(base + ".so")) {
        { std::ofstream f(base + ".glsl", std::ios::binary); f << source; }
        std::string cmd = std::string("python3 '") + tool_env + "' '" + base + ".glsl' " + stage + " '" + base + ".so' 2> '" + base + ".log'";
        if (std::system(cmd.c_str()) != 0 || !file_exists(base + ".so")) {
            std::ifstream f(base + ".log");
            log.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
            if (log.empty()) log = "soft GL: translator failed: " + cmd;
            return false;
        }
    }
    lib.handle = ::dlopen((base + ".so").c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!lib.handle) { log = std::string("soft GL: dlopen: ") + ::dlerror(); return false; }
#define RFK_SYM(field, name) lib.field = reinterpret_cast<decltype(lib.field)>(::dlsym(lib.handle, name))
    RFK_SYM(bind, "rfk_sh_bind"); RFK_SYM(location, "rfk_sh_uniform_location"); RFK_SYM(set, "rfk_sh_set_uniform");
    RFK_SYM(dispatch, "rfk_sh_dispatch"); RFK_SYM(private_bytes, "rfk_sh_private_bytes"); RFK_SYM(varyings, "rfk_sh_varyings");
    RFK_SYM(run_vertex, "rfk_sh_run_vertex"); RFK_SYM(run_fragment, "rfk_sh_run_fragment");
#undef RFK_SYM
    if (!lib.bind || !lib.location || !lib.set || !lib.dispatch) { log = "soft GL: shader library lacks its entry points"; return false; }
    return true;
}

void bind_all(stage_library& lib) {
    for (auto& [index, name] : g_bindings) {
        auto it = g_buffers.find(name);
        if (it != g_buffers.end()) lib.bind((int)index, it->second->data());
    }
}

std::vector<unsigned char>* find_buffer(uint name) {
    auto it = g_buffers.find(name);
    return it == g_buffers.end() ? nullptr : it->second.get();
}

}  // namespace

// ------------------------------------------------------------------------------------------ buffers
void create_buffers(int n, uint* names) {
    for (int i = 0; i < n; i++) { names[i] = g_next_name++; g_buffers[names[i]] = std::make_unique<std::vector<unsigned char>>(); }
}
void buffer_storage(uint name, std::ptrdiff_t bytes, const void* data) {
    auto* b = find_buffer(name);
    if (!b) return;
    b->assign((std::size_t)bytes, 0);  // GL leaves it undefined; zero keeps runs reproducible
    if (data) std::memcpy(b->data(), data, (std::size_t)bytes);
}
void buffer_sub_data(uint name, std::ptrdiff_t offset, std::ptrdiff_t bytes, const void* data) {
    auto* b = find_buffer(name);
    if (b && offset >= 0 && (std::size_t)(offset + bytes) <= b->size()) std::memcpy(b->data() + offset, data, (std::size_t)bytes);
}
void get_buffer_sub_data(uint name, std::ptrdiff_t offset, std::ptrdiff_t bytes, void* out) {
    auto* b = find_buffer(name);
    if (b && offset >= 0 && (std::size_t)(offset + bytes) <= b->size()) std::memcpy(out, b->data() + offset, (std::size_t)bytes);
}
void clear_buffer(uint name) { if (auto* b = find_buffer(name)) std::fill(b->begin(), b->end(), 0); }
void delete_buffers(int n, const uint* names) { for (int i = 0; i < n; i++) g_buffers.erase(names[i]); }
void* map_buffer(uint name) { auto* b = find_buffer(name); return b ? b->data() : nullptr; }
void bind_buffer_base(uint index, uint name) { g_bindings[index] = name; }
std::size_t buffer_size(uint name) { auto* b = find_buffer(name); return b ? b->size() : 0; }
void* buffer_data(uint name) { auto* b = find_buffer(name); return b ? b->data() : nullptr; }
uint bound_buffer(uint index) { auto it = g_bindings.find(index); return it == g_bindings.end() ? 0 : it->second; }

// ------------------------------------------------------------------------------------------ programs
uint create_shader(uint type) { uint n = g_next_name++; g_shaders[n].type = type; return n; }
void shader_source(uint shader, const char* text) { g_shaders[shader].source = text; g_sources.emplace_back(text); }
uint create_program() { uint n = g_next_name++; g_programs[n]; return n; }
void attach_shader(uint program, uint shader) { g_programs[program].shaders.push_back(shader); }
void link_program(uint program) {
    auto& p = g_programs[program];
    p.linked = true;
    for (uint s : p.shaders) {
        auto& sh = g_shaders[s];
        stage_library& lib = sh.type == COMPUTE_SHADER ? p.compute : sh.type == VERTEX_SHADER ? p.vertex : p.fragment;
        if (!load_stage(sh.source, sh.type, lib, p.log)) { p.linked = false; return; }
    }
}
int link_status(uint program) { return g_programs[program].linked ? 1 : 0; }
std::string program_log(uint program) { return g_programs[program].log; }
void use_program(uint program) { g_current = program; }
int uniform_location(uint program, const char* name) {
    auto& p = g_programs[program];
    for (std::size_t i = 0; i < p.uniform_names.size(); i++) if (p.uniform_names[i] == name) return (int)i;
    p.uniform_names.emplace_back(name);
    return (int)p.uniform_names.size() - 1;
}
void set_uniform(int location, const void* data, std::size_t bytes) {
    auto it = g_programs.find(g_current);
    if (it == g_programs.end() || location < 0 || (std::size_t)location >= it->second.uniform_names.size()) return;
    const std::string& name = it->second.uniform_names[location];
    for (stage_library* lib : {&it->second.compute, &it->second.vertex, &it->second.fragment})
        if (lib->handle) lib->set(lib->location(name.c_str()), data, bytes);
    const unsigned char* p = static_cast<const unsigned char*>(data);
    g_uniform_log.push_back({g_current, name, std::vector<unsigned char>(p, p + bytes)});
}
void dispatch_compute(uint nx, uint ny, uint nz) {
    auto it = g_programs.find(g_current);
    if (it == g_programs.end() || !it->second.compute.handle) return;
    bind_all(it->second.compute);
    it->second.compute.dispatch(nx, ny, nz);
    g_dispatch_log.push_back({g_current, nx, ny, nz});
}

// ------------------------------------------------------------------------------------------ GL_POINTS, additive blend
// OpenGL 4.6 core, section 14.4 (points) with GL_PROGRAM_POINT_SIZE: the viewport transform of main.cpp:497 maps clip
// space to the W x H window; a point of size s produces one fragment for every pixel whose centre lies inside the
// s x s square centred on the point; gl_PointCoord (origin upper left) is s = 1/2 + (xf + 1/2 - xw) / size,
// t = 1/2 - (yf + 1/2 - yw) / size. Points whose z lies outside the clip volume are discarded (the vertex shader uses
// that for empty bins). A point whose CENTRE is outside the x/y clip planes is discarded by the letter of section 13.7;
// implementations with a guard band keep it. RFK_SOFTGL_CLIP_POINT_XY=1 selects the strict reading.
void set_color_target(float* rgba, int width, int height) { g_target = rgba; g_target_w = width; g_target_h = height; }
void draw_points(int first, int count) {
    auto it = g_programs.find(g_current);
    if (it == g_programs.end() || !it->second.vertex.handle || !it->second.fragment.handle || !g_target) return;
    auto& V = it->second.vertex;
    auto& F = it->second.fragment;
    bind_all(V);
    bind_all(F);
    const bool strict_xy = std::getenv("RFK_SOFTGL_CLIP_POINT_XY") && std::atoi(std::getenv("RFK_SOFTGL_CLIP_POINT_XY")) != 0;
    std::vector<unsigned char> vpriv(V.private_bytes() + 16), fpriv(F.private_bytes() + 16);
    const char* names[32]; std::size_t offs[32], sizes[32]; int outs[32];
    struct vary { std::string name; std::size_t off, bytes; int is_out; };
    std::vector<vary> vv, fv;
    int n = V.varyings(names, offs, sizes, outs, 32);
    for (int i = 0; i < n; i++) vv.push_back({names[i], offs[i], sizes[i], outs[i]});
    n = F.varyings(names, offs, sizes, outs, 32);
    for (int i = 0; i < n; i++) fv.push_back({names[i], offs[i], sizes[i], outs[i]});
    const vary* out_color = nullptr;
    for (auto& f : fv) if (f.is_out) out_color = &f;
    if (!out_color) return;
    const int W = g_target_w, H = g_target_h;
    for (int v = first; v < first + count; v++) {
        float pos[4], size = 1.0f;
        V.run_vertex(v, vpriv.data(), pos, &size);
        if (!(pos[3] > 0.0f) || pos[2] < -pos[3] || pos[2] > pos[3]) continue;
        if (strict_xy && (pos[0] < -pos[3] || pos[0] > pos[3] || pos[1] < -pos[3] || pos[1] > pos[3])) continue;
        const float xw = (pos[0] / pos[3] + 1.0f) * 0.5f * float(W), yw = (pos[1] / pos[3] + 1.0f) * 0.5f * float(H);
        const float half = size * 0.5f;
        int x0 = (int)std::ceil(xw - half - 0.5f), x1 = (int)std::floor(xw + half - 0.5f);
        int y0 = (int)std::ceil(yw - half - 0.5f), y1 = (int)std::floor(yw + half - 0.5f);
        for (int yf = std::max(y0, 0); yf <= std::min(y1, H - 1); yf++) {
            for (int xf = std::max(x0, 0); xf <= std::min(x1, W - 1); xf++) {
                const float cx = float(xf) + 0.5f, cy = float(yf) + 0.5f;
                if (!(cx > xw - half && cx < xw + half && cy > yw - half && cy < yw + half)) continue;  // pixel centre strictly inside the square
                std::memset(fpriv.data(), 0, fpriv.size());
                for (auto& f : fv)
                    if (!f.is_out)
                        for (auto& o : vv)
                            if (o.is_out && o.name == f.name) std::memcpy(fpriv.data() + f.off, vpriv.data() + o.off, std::min(f.bytes, o.bytes));
                const float s = 0.5f + (cx - xw) / size, t = 0.5f - (cy - yw) / size;
                if (!F.run_fragment(s, t, fpriv.data())) continue;
                float c[4];
                std::memcpy(c, fpriv.data() + out_color->off, 16);
                // window y grows upwards; row 0 of the target is the bottom row of the window, as glGetTextureImage returns it
                float* dst = g_target + 4 * ((std::size_t)yf * W + xf);
                for (int k = 0; k < 4; k++) dst[k] += c[k];  // glBlendFunc(GL_ONE, GL_ONE)
            }
        }
    }
}

std::vector<std::string>& shader_sources() { return g_sources; }
std::vector<uniform_record>& uniform_log() { return g_uniform_log; }
std::vector<dispatch_record>& dispatch_log() { return g_dispatch_log; }
void reset_logs() { g_sources.clear(); g_uniform_log.clear(); g_dispatch_log.clear(); }
}  // namespace softgl
