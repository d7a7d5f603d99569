// host_scene.cpp — the host-side pieces of the reference that sit either side of the GPU
// path and must be reproduced exactly because they are its INPUT or the
// step right after it (§8 f2): the one sprite, the palette, the default scene generator, the
// default light, the key map, the synthetic-scene recipe and the debug overlay.
// Plain C++ (no CUDA); exported through the same C ABI (include/par/par.h).
#include <cstdlib>

#include "par/par.h"

namespace {

// One texel of make_tile_floor(), /root/reference/src/sprites.hpp:73-364: a 20x40 billboard
// faking a 20^3 cube.  Rows 0-19 = top face (normal +y, depth 19-row); rows 20-39 = front
// face (normal -z, depth 0).  The top face carries a 12x12 checker of palette 2/3 on a
// palette-0 field; the front face is palette 2 framed by palette 1 (2 px left/right, 2 px
// bottom).
struct Texel {
    int color, depth;
    float nx, ny, nz;
};

Texel tile_floor_texel(int row, int col) {
    Texel t{};
    const bool top = row < PAR_SPRITE_W;
    if (top) {
        t.depth = PAR_SPRITE_W - 1 - row;
        t.ny = 1.f;
        const bool inside = row >= 4 && row <= 15 && col >= 4 && col <= 15;
        if (inside) {
            const int qr = (row - 4) / 6, qc = (col - 4) / 6;  // 2x2 quadrants of 6x6
            t.color = ((qr + qc) & 1) ? 3 : 2;
        }
    } else {
        t.nz = -1.f;
        const bool frame = col <= 1 || col >= PAR_SPRITE_W - 2 || row >= PAR_SPRITE_H - 2;
        t.color = frame ? 1 : 2;
    }
    return t;
}

struct SplitMix64 {
    uint64_t s;
    uint64_t next() {
        s += 0x9E3779B97F4A7C15ull;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
};

struct SceneWriter {
    par_aabb* out;
    int cap, n = 0;
    void cube(int x, int y, int z) {
        if (out && n < cap) {
            par_aabb& a = out[n];
            a.px = static_cast<int16_t>(x);
            a.py = static_cast<int16_t>(y);
            a.pz = static_cast<int16_t>(z);
            a.ex = a.ey = a.ez = 20;
            a.pad[0] = a.pad[1] = 0;
        }
        ++n;
    }
};

}  // namespace

extern "C" {

void par_sprite_tile_floor(par_sprite* out) {
    for (int row = 0; row < PAR_SPRITE_H; ++row)
        for (int col = 0; col < PAR_SPRITE_W; ++col) {
            const Texel t = tile_floor_texel(row, col);
            const int i = row * PAR_SPRITE_W + col;
            out->color[i] = t.color;
            out->depth[i] = t.depth;
            out->normal[i][0] = t.nx;
            out->normal[i][1] = t.ny;
            out->normal[i][2] = t.nz;
        }
}

void par_palette_default(par_color out[4]) {  // sprites.hpp:60-65
    const uint8_t grey[4] = {100, 140, 200, 240};
    for (int i = 0; i < 4; ++i) out[i] = par_color{grey[i], grey[i], grey[i], 0};
}

// alternative.cpp:519-599.  The generator is written against the DEFAULT view constants
// (480 x 320 x 320) whatever the rendered view is, so that a larger view shows the same
// scene (SURVEY.md §8c tier 1).  Every entity is a 20^3 cube drawing sprite 0 (quirk Q1).
int par_scene_default(par_aabb* out, int cap) {
    constexpr int kW = 480, kL = 320;
    SceneWriter w{out, cap};
    w.cube(kW / 2, 36, kL / 4);  // entity 0: the player
    // floor: a kW x kL lattice of cubes at 20-unit pitch, minus a hole around the centre
    for (int i = 0; i < kW; ++i)
        for (int j = 0; j < kL; ++j) {
            const int x = 20 * i, z = 20 * j;
            const bool hole = x >= kW / 2 - 40 && x < kW / 2 + 40 && z > kL / 2 - 40 && z < kL / 2 + 40;
            if (!hole) w.cube(x, 0, z);
        }
    // left wall block: 6 columns x 5 layers marching from z = kL towards the camera, with
    // the top-right 2x2 corner notched out
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < kL - 10; ++j)
            for (int k = 1; k < 6; ++k)
                if (!(i >= 4 && k >= 4)) w.cube(20 * i, 20 * k, kL - 20 * j);
    // right strip: two columns of cubes one layer up
    for (int i = 1; i <= 2; ++i)
        for (int j = 0; j < kL; ++j) w.cube(kW - 20 * i, 20, 20 * j);
    // back row
    for (int i = 1; i < 20; ++i) w.cube(kW - 40 - 20 * i, 20, kL - 60);
    return w.n;
}

void par_light_default(par_light* out) {  // alternative.cpp:624-626 with 480/320/320
    *out = par_light{480, 160, 80, 10};
}

void par_scene_synthetic(int width, int height, int length, uint64_t seed, int n,
                         par_aabb* out_aabbs, int n_lights, par_light* out_lights) {
    // The recipe is stated for the 4K / 8K views; for views too small for its ranges (width or length
    // <= 20, height <= 600) the modulus is clamped to 1 instead of dividing by zero or wrapping.
    const auto span = [](int v) { return static_cast<uint64_t>(v > 1 ? v : 1); };
    SplitMix64 rng{seed};
    SceneWriter w{out_aabbs, n};
    for (int e = 0; e < n; ++e) {
        const int x = static_cast<int>(rng.next() % span(width - 20));
        const int y = static_cast<int>(rng.next() % 200u);
        const int z = static_cast<int>(rng.next() % span(length - 20));
        w.cube(x, y, z);
    }
    for (int l = 0; l < n_lights; ++l) {
        par_light& lt = out_lights[l];
        lt.x = static_cast<int16_t>(rng.next() % span(width));
        lt.y = static_cast<int16_t>(40 + rng.next() % 400u);
        lt.z = static_cast<int16_t>(rng.next() % span(height - 600));
        lt.radius = 10;
    }
}

void par_apply_key(int key, par_aabb* player, par_light* light) {  // alternative.cpp:641-681
    constexpr int16_t kStep = 5;
    switch (key) {
        case 'L': player->px -= kStep; break;  // SDLK_LEFT
        case 'R': player->px += kStep; break;  // SDLK_RIGHT
        case 'U': player->pz += kStep; break;  // SDLK_UP
        case 'D': player->pz -= kStep; break;  // SDLK_DOWN
        case 'p': player->py -= kStep; break;  // SDLK_PAGEDOWN
        case 'P': player->py += kStep; break;  // SDLK_PAGEUP
        case 'a': light->z -= kStep; break;
        case 'k': light->z += kStep; break;
        case 'j': light->y -= kStep; break;
        case 'u': light->y += kStep; break;
        case 'h': light->x -= kStep; break;
        case 'o': light->x += kStep; break;
        default: break;
    }
}

// alternative.cpp:139-175 (Bresenham) + 762-772 (red line from the surface point under the
// cursor to light 0, per-pixel bounds check).  `length` is not needed: the projection uses
// only the view height.
void par_draw_overlay(int width, int height, const par_pixel* gbuf, const par_light* light,
                      int cursor_x, int cursor_y, par_color* frame) {
    par_draw_overlay_at(width, height, &gbuf[static_cast<size_t>(cursor_y) * width + cursor_x], light, cursor_x, frame);
}

void par_draw_overlay_at(int width, int height, const par_pixel* under_cursor, const par_light* light,
                         int cursor_x, par_color* frame) {
    const par_pixel& under = *under_cursor;
    int x = cursor_x, y = height - (under.y + under.z);
    const int x_end = light->x, y_end = height - (light->y + light->z);
    const int span_x = std::abs(x_end - x), span_y = -std::abs(y_end - y);
    const int dir_x = x < x_end ? 1 : -1, dir_y = y < y_end ? 1 : -1;
    int error = span_x + span_y;
    const par_color red{255, 0, 0, 255};
    while (true) {
        if (x >= 0 && y >= 0 && x < width && y < height) frame[static_cast<size_t>(y) * width + x] = red;
        if (x == x_end && y == y_end) break;
        const int twice = 2 * error;
        if (twice >= span_y) {
            if (x == x_end) break;
            error += span_y;
            x += dir_x;
        }
        if (twice <= span_x) {
            if (y == y_end) break;
            error += span_x;
            y += dir_y;
        }
    }
}

// FNV-1a-64 of a byte range: the per-frame hash of the 240-frame goldens, taken over the
// W*H*4 RGBA bytes a host hands to its frame sink (alternative.cpp:774-788).
uint64_t par_fnv1a64(const void* data, size_t bytes) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    uint64_t h = 1469598103934665603ull;
    for (size_t k = 0; k < bytes; k++) {
        h ^= p[k];
        h *= 1099511628211ull;
    }
    return h;
}

}  // extern "C"
