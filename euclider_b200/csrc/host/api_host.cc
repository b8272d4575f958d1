// C-ABI entry points of the scene front end (host only).  See include/euclider_b200.h.
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>

#include "error.h"
#include "euclider_b200.h"
#include "scene_parse.h"

namespace eucl {
namespace {
thread_local std::string g_last_error;
}
void set_last_error(const std::string& message) { g_last_error = message; }
int fail(int status, const std::string& message) {
    g_last_error = message;
    return status;
}
} // namespace eucl

struct EuclParsedScene {
    std::unique_ptr<eucl::ParsedScene> scene;
    bool dirty = true;
};

extern "C" {

const char* eucl_last_error(void) { return eucl::g_last_error.c_str(); }

const char* eucl_version(void) { return "euclider_b200 0.1 (sm_100a, f64)"; }

int eucl_scene_parse(const char* json_text, EuclParsedScene** out) {
    if (!json_text || !out) return eucl::fail(EUCL_ERR_INVALID_ARGUMENT, "eucl_scene_parse: null argument");
    *out = nullptr;
    std::unique_ptr<eucl::ParsedScene> scene;
    std::string error;
    int status = eucl::parse_scene(json_text, &scene, &error);
    if (status != EUCL_OK) return eucl::fail(status, error);
    auto* p = new EuclParsedScene();
    p->scene = std::move(scene);
    *out = p;
    return EUCL_OK;
}

int eucl_parsed_texture_count(const EuclParsedScene* p) { return p ? (int)p->scene->texture_paths.size() : 0; }

const char* eucl_parsed_texture_path(const EuclParsedScene* p, int slot) {
    if (!p || slot < 0 || slot >= (int)p->scene->texture_paths.size()) return nullptr;
    return p->scene->texture_paths[(size_t)slot].c_str();
}

int eucl_parsed_set_texture(EuclParsedScene* p, int slot, uint32_t width, uint32_t height, const uint8_t* rgba8) {
    if (!p || !rgba8 || slot < 0 || slot >= (int)p->scene->texture_paths.size() || width == 0 || height == 0)
        return eucl::fail(EUCL_ERR_INVALID_ARGUMENT, "eucl_parsed_set_texture: bad slot or empty image");
    auto& s = *p->scene;
    s.texture_pixels[(size_t)slot].assign(rgba8, rgba8 + (size_t)width * height * 4);
    s.textures[(size_t)slot].width = width;
    s.textures[(size_t)slot].height = height;
    s.texture_set[(size_t)slot] = true;
    p->dirty = true;
    return EUCL_OK;
}

const EuclFlatScene* eucl_parsed_flat(EuclParsedScene* p) {
    if (!p) return nullptr;
    if (p->dirty) {
        p->scene->refresh_flat();
        p->dirty = false;
    }
    return &p->scene->flat;
}

void eucl_parsed_destroy(EuclParsedScene* p) { delete p; }

int eucl_write_ppm(const char* path, uint32_t width, uint32_t height, const uint8_t* rgb) {
    if (!path || !rgb || width == 0 || height == 0) return eucl::fail(EUCL_ERR_INVALID_ARGUMENT, "eucl_write_ppm: bad argument");
    FILE* f = std::fopen(path, "wb");
    if (!f) return eucl::fail(EUCL_ERR_INVALID_ARGUMENT, std::string("eucl_write_ppm: cannot open ") + path);
    std::fprintf(f, "P6\n%u %u\n255\n", width, height);
    bool ok = true;
    for (uint32_t y = 0; y < height && ok; ++y) // row 0 of the buffer is the BOTTOM of the picture
        ok = std::fwrite(rgb + (size_t)(height - 1 - y) * width * 3, 1, (size_t)width * 3, f) == (size_t)width * 3;
    ok = std::fclose(f) == 0 && ok;
    return ok ? EUCL_OK : eucl::fail(EUCL_ERR_INVALID_ARGUMENT, std::string("eucl_write_ppm: short write to ") + path);
}

uint32_t eucl_band_rows_for_rank(const EuclRenderOpts* o) {
    if (!o || o->height == 0) return 0;
    uint32_t band = o->band_rows ? o->band_rows : o->height;
    uint32_t world = o->band_world ? o->band_world : 1;
    uint32_t n_bands = (o->height + band - 1) / band;
    uint32_t rows = 0;
    for (uint32_t b = o->band_rank; b < n_bands; b += world) {
        uint32_t y0 = b * band;
        uint32_t y1 = y0 + band < o->height ? y0 + band : o->height;
        rows += y1 - y0;
    }
    return rows;
}

} // extern "C"
