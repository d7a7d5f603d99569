/* Headless stand-in for <SDL2/SDL.h> — TEST INFRASTRUCTURE ONLY.
 *
 * Lets the UNMODIFIED reference translation unit (/root/reference/src/alternative.cpp)
 * compile and run without SDL2, a window or a keyboard, so that its frames can pin the
 * oracle (tier 0 / tier 1).  Only the calls the reference makes exist
 * (alternative.cpp:604-617, 629-686, 774-788, 816, 821-824).
 *
 * Behaviour is driven by environment variables:
 *   PAR_REF_FRAMES   frames to render before an Esc key-up is injected (default 1)
 *   PAR_REF_SCRIPT   "C" | "D" | unset — scripted key-downs per frame:
 *                    C = player keys 30xRIGHT 20xUP 50xLEFT 30xDOWN 30xPGUP 40xRIGHT
 *                        30xPGDN 9xUP (one per frame, frames 1..239; then idle)
 *                    D = C plus light key 'o' every frame >= 1
 *   PAR_REF_HASHES   file: one line "%03d %016llx\n" (frame, FNV-1a-64 of the RGBA bytes
 *                    handed to SDL_UnlockTexture) per frame
 *   PAR_REF_DUMP     file: raw RGBA bytes of the frames listed in PAR_REF_DUMP_FRAMES
 *                    (comma list, default "0"), concatenated
 *   PAR_REF_TIMES    file: one line "frame ns" per frame; ns = time from the end of
 *                    event pumping to SDL_LockTexture, i.e. alternative.cpp:689-772
 *   PAR_REF_DUMP_PRE / PAR_REF_DUMP_GBUF
 *                    files written by par_stub_dump_pre(), a hook that only the
 *                    sed-instrumented tier-1 build calls just before draw_line
 *                    (alternative.cpp:762): the shaded frame without the debug overlay and
 *                    the raw Pixel[] G-buffer, for the frames in PAR_REF_DUMP_FRAMES
 *   PAR_REF_SCENE    file read by par_stub_load_scene(), a second hook of the tier-1 build
 *                    (called once, just before the frame loop, alternative.cpp:628): replaces
 *                    the built-in scene and light by the file's — int32 n_boxes, int32
 *                    n_lights, n_boxes x 8 int16 (position xyz, extent xyz, 2 pad),
 *                    n_lights x 4 int16 (x y z radius) — through the reference's own
 *                    Entities::insert, so the REAL reference renders arbitrary scenes
 */
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

struct SDL_Window {};
struct SDL_Renderer {};
struct SDL_Texture {
    int w, h;
    unsigned char* pixels;
};
struct SDL_Rect {
    int x, y, w, h;
};

enum {
    SDL_INIT_VIDEO = 0x20,
    SDL_WINDOWPOS_UNDEFINED = 0x1FFF0000,
    SDL_RENDERER_SOFTWARE = 1,
    SDL_PIXELFORMAT_RGB888 = 0x16161804,
    SDL_TEXTUREACCESS_STREAMING = 1,
};
enum { SDL_KEYDOWN = 0x300, SDL_KEYUP = 0x301, SDL_MOUSEMOTION = 0x400 };
enum {
    SDLK_ESCAPE = 27,
    SDLK_a = 'a',
    SDLK_h = 'h',
    SDLK_j = 'j',
    SDLK_k = 'k',
    SDLK_o = 'o',
    SDLK_u = 'u',
    SDLK_RIGHT = 0x4000004F,
    SDLK_LEFT = 0x40000050,
    SDLK_DOWN = 0x40000051,
    SDLK_UP = 0x40000052,
    SDLK_PAGEUP = 0x4000004B,
    SDLK_PAGEDOWN = 0x4000004E,
};

struct SDL_Keysym {
    int sym;
};
struct SDL_KeyboardEvent {
    unsigned type;
    SDL_Keysym keysym;
};
union SDL_Event {
    unsigned type;
    SDL_KeyboardEvent key;
};

namespace par_stub {

struct State {
    int frames_wanted = 1;
    int frames_done = 0;
    int pumped_for_frame = -1;  // frame whose events have been queued
    int queue[4];
    int queue_len = 0, queue_pos = 0;
    char script = 0;
    FILE* hashes = nullptr;
    FILE* dump = nullptr;
    FILE* times = nullptr;
    FILE* dump_pre = nullptr;
    FILE* dump_gbuf = nullptr;
    bool dump_frame[4096] = {};
    std::chrono::steady_clock::time_point t0, frame_begin;
    bool init = false;
};

inline State& state() {
    static State s;
    if (!s.init) {
        s.init = true;
        s.t0 = std::chrono::steady_clock::now();
        if (char const* e = getenv("PAR_REF_FRAMES")) s.frames_wanted = atoi(e);
        if (char const* e = getenv("PAR_REF_SCRIPT")) s.script = e[0];
        if (char const* e = getenv("PAR_REF_HASHES")) s.hashes = fopen(e, "w");
        if (char const* e = getenv("PAR_REF_DUMP")) s.dump = fopen(e, "wb");
        if (char const* e = getenv("PAR_REF_TIMES")) s.times = fopen(e, "w");
        if (char const* e = getenv("PAR_REF_DUMP_PRE")) s.dump_pre = fopen(e, "wb");
        if (char const* e = getenv("PAR_REF_DUMP_GBUF")) s.dump_gbuf = fopen(e, "wb");
        char const* list = getenv("PAR_REF_DUMP_FRAMES");
        if (!list) list = "0";
        while (*list) {
            int f = atoi(list);
            if (f >= 0 && f < 4096) s.dump_frame[f] = true;
            while (*list && *list != ',') list++;
            if (*list == ',') list++;
        }
    }
    return s;
}

// Script C player key for frame f (1-based over the 239 scripted frames), 0 = none.
inline int script_c_key(int f) {
    static int const runs[8][2] = {{30, SDLK_RIGHT}, {20, SDLK_UP},   {50, SDLK_LEFT},
                                   {30, SDLK_DOWN},  {30, SDLK_PAGEUP}, {40, SDLK_RIGHT},
                                   {30, SDLK_PAGEDOWN}, {9, SDLK_UP}};
    int k = f - 1;
    if (k < 0) return 0;
    for (auto const& r : runs) {
        if (k < r[0]) return r[1];
        k -= r[0];
    }
    return 0;
}

}  // namespace par_stub

// Hook for the instrumented tier-1 build only (never called by the unmodified source).
inline void par_stub_dump_pre(void const* frame, size_t frame_bytes, void const* gbuf,
                              size_t gbuf_bytes) {
    auto& s = par_stub::state();
    if (s.frames_done >= 4096 || !s.dump_frame[s.frames_done]) return;
    if (s.dump_pre) {
        fwrite(frame, 1, frame_bytes, s.dump_pre);
        fflush(s.dump_pre);
    }
    if (s.dump_gbuf) {
        fwrite(gbuf, 1, gbuf_bytes, s.dump_gbuf);
        fflush(s.dump_gbuf);
    }
}

// Hook for the instrumented tier-1 build only: load scene + lights from PAR_REF_SCENE.
// Generic over the reference's types (Entities<N>*, std::vector<Light>&); entities go through
// the reference's own insert() (alternative.cpp:104-109), so sprites etc. are whatever it does.
template <class EntitiesPtr, class LightVector>
inline void par_stub_load_scene(EntitiesPtr p_entities, LightVector& lights) {
    char const* path = getenv("PAR_REF_SCENE");
    if (!path) return;
    FILE* f = fopen(path, "rb");
    int32_t hdr[2] = {0, 0};
    if (!f || fread(hdr, 4, 2, f) != 2) {
        fprintf(stderr, "par_stub_load_scene: cannot read %s\n", path);
        exit(3);
    }
    p_entities->aabbs.clear();
    p_entities->sprites.clear();
    p_entities->last_entity_index = 0;
    for (int i = 0; i < hdr[0]; i++) {
        int16_t v[8];
        if (fread(v, 2, 8, f) != 8) exit(3);
        typename std::remove_pointer<EntitiesPtr>::type::Entity e{};
        e.aabb.position.x = v[0];
        e.aabb.position.y = v[1];
        e.aabb.position.z = v[2];
        e.aabb.extent.x = v[3];
        e.aabb.extent.y = v[4];
        e.aabb.extent.z = v[5];
        p_entities->insert(e);
    }
    lights.clear();
    for (int i = 0; i < hdr[1]; i++) {
        int16_t v[4];
        if (fread(v, 2, 4, f) != 4) exit(3);
        typename LightVector::value_type l{};
        l.x = v[0];
        l.y = v[1];
        l.z = v[2];
        l.radius = v[3];
        lights.push_back(l);
    }
    fclose(f);
}

inline int SDL_InitSubSystem(unsigned) { return 0; }
inline SDL_Window* SDL_CreateWindow(char const*, int, int, int, int, unsigned) {
    static SDL_Window w;
    return &w;
}
inline SDL_Renderer* SDL_CreateRenderer(SDL_Window*, int, unsigned) {
    static SDL_Renderer r;
    return &r;
}
inline SDL_Texture* SDL_CreateTexture(SDL_Renderer*, unsigned, int, int w, int h) {
    auto* t = new SDL_Texture{w, h, static_cast<unsigned char*>(calloc((size_t)w * h, 4))};
    return t;
}

inline int SDL_PollEvent(SDL_Event* ev) {
    auto& s = par_stub::state();
    int f = s.frames_done;  // the frame about to be rendered
    if (s.pumped_for_frame != f) {
        s.pumped_for_frame = f;
        s.queue_len = s.queue_pos = 0;
        if (f >= s.frames_wanted) {
            s.queue[s.queue_len++] = -SDLK_ESCAPE;  // negative = key-up
        } else if (s.script == 'C' || s.script == 'D') {
            if (int k = par_stub::script_c_key(f)) s.queue[s.queue_len++] = k;
            if (s.script == 'D' && f >= 1) s.queue[s.queue_len++] = SDLK_o;
        }
    }
    if (s.queue_pos < s.queue_len) {
        int k = s.queue[s.queue_pos++];
        memset(ev, 0, sizeof *ev);
        if (k < 0) {
            ev->key.type = SDL_KEYUP;
            ev->key.keysym.sym = -k;
        } else {
            ev->key.type = SDL_KEYDOWN;
            ev->key.keysym.sym = k;
        }
        return 1;
    }
    s.frame_begin = std::chrono::steady_clock::now();
    return 0;
}

inline unsigned SDL_GetMouseState(int* x, int* y) {
    *x = 0;
    *y = 0;
    return 0;
}

inline int SDL_LockTexture(SDL_Texture* t, SDL_Rect const*, void** pixels, int* pitch) {
    auto& s = par_stub::state();
    if (s.times) {
        auto ns = std::chrono::duration_cast<std::chrono::nanoseconds>(
                      std::chrono::steady_clock::now() - s.frame_begin)
                      .count();
        fprintf(s.times, "%d %lld\n", s.frames_done, (long long)ns);
        fflush(s.times);
    }
    *pixels = t->pixels;
    *pitch = t->w * 4;
    return 0;
}

inline void SDL_UnlockTexture(SDL_Texture* t) {
    auto& s = par_stub::state();
    size_t n = (size_t)t->w * t->h * 4;
    if (s.hashes) {
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < n; i++) {
            h ^= t->pixels[i];
            h *= 1099511628211ull;
        }
        fprintf(s.hashes, "%03d %016llx\n", s.frames_done, (unsigned long long)h);
        fflush(s.hashes);
    }
    if (s.dump && s.frames_done < 4096 && s.dump_frame[s.frames_done]) {
        fwrite(t->pixels, 1, n, s.dump);
        fflush(s.dump);
    }
    s.frames_done++;
}

inline int SDL_RenderCopy(SDL_Renderer*, SDL_Texture*, SDL_Rect const*, SDL_Rect const*) {
    return 0;
}
inline void SDL_RenderPresent(SDL_Renderer*) {}
inline unsigned SDL_GetTicks() {
    auto& s = par_stub::state();
    return (unsigned)std::chrono::duration_cast<std::chrono::milliseconds>(
               std::chrono::steady_clock::now() - s.t0)
        .count();
}
inline void SDL_DestroyTexture(SDL_Texture*) {}
inline void SDL_DestroyWindow(SDL_Window*) {}
inline void SDL_DestroyRenderer(SDL_Renderer*) {}
inline void SDL_VideoQuit() {}
