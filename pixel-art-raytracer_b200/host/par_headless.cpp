// par_headless.cpp — headless stand-in for the reference's main loop
// (/root/reference/src/alternative.cpp:502-833) on top of the B200 render path.
// Builds the default scene with Entities::insert (519-599), applies a scripted key sequence
// (641-681) before each frame, renders with par::FrameRenderer (replaces 689-760), draws the
// debug overlay (762-772) and prints one "%03d %016llx" FNV-1a-64 line per frame — the same
// format the SDL stub of oracle/ uses for the real reference, so the two can be diffed.
//
//   par_headless [--view W H L] [--frames N] [--script C|D] [--device D] [--ppm out.ppm]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "par/reference_types.hpp"

namespace {

int script_c_key(int frame) {  // 239 player keys, frame 0 gets none
    static const int runs[8][2] = {{30, 'R'}, {20, 'U'}, {50, 'L'}, {30, 'D'},
                                   {30, 'P'}, {40, 'R'}, {30, 'p'}, {9, 'U'}};
    int k = frame - 1;
    if (k < 0) return 0;
    for (const auto& r : runs) {
        if (k < r[0]) return r[1];
        k -= r[0];
    }
    return 0;
}

unsigned long long fnv1a64(const unsigned char* p, size_t n) {
    unsigned long long h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

}  // namespace

int main(int argc, char** argv) {
    int W = 480, H = 320, L = 320, frames = 1, device = 0;
    char script = 0;
    const char* ppm = nullptr;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--view") && i + 3 < argc) {
            W = atoi(argv[++i]);
            H = atoi(argv[++i]);
            L = atoi(argv[++i]);
        } else if (!strcmp(argv[i], "--frames") && i + 1 < argc) {
            frames = atoi(argv[++i]);
        } else if (!strcmp(argv[i], "--script") && i + 1 < argc) {
            script = argv[++i][0];
        } else if (!strcmp(argv[i], "--device") && i + 1 < argc) {
            device = atoi(argv[++i]);
        } else if (!strcmp(argv[i], "--ppm") && i + 1 < argc) {
            ppm = argv[++i];
        } else {
            fprintf(stderr, "usage: %s [--view W H L] [--frames N] [--script C|D] [--device D] [--ppm f]\n", argv[0]);
            return 2;
        }
    }
    try {
        par::Entities<> entities;
        par::make_default_scene(entities);
        std::vector<par::Light> lights(1);
        par_light_default(reinterpret_cast<par_light*>(lights.data()));  // alternative.cpp:624-626

        par::FrameRenderer renderer(W, H, L, device);
        std::vector<par::Color> texture(static_cast<size_t>(W) * H);
        std::vector<par::Pixel> gbuf(static_cast<size_t>(W) * H);
        double gpu_ms = 0, wall_ms = 0;
        for (int f = 0; f < frames; f++) {
            if (script == 'C' || script == 'D') {
                par_aabb* player = reinterpret_cast<par_aabb*>(&entities.aabbs[0]);
                par_light* light = reinterpret_cast<par_light*>(&lights[0]);
                if (int k = script_c_key(f)) par_apply_key(k, player, light);
                if (script == 'D' && f >= 1) par_apply_key('o', player, light);
            }
            par_stats st{};
            auto t0 = std::chrono::steady_clock::now();
            renderer.render_frame(entities, lights, texture.data(), gbuf.data(), &st);
            wall_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            gpu_ms += st.ms_grid_build + st.ms_total;
            par_draw_overlay(W, H, reinterpret_cast<const par_pixel*>(gbuf.data()),
                             reinterpret_cast<const par_light*>(&lights[0]), 0, 0,
                             reinterpret_cast<par_color*>(texture.data()));
            printf("%03d %016llx\n", f, fnv1a64(reinterpret_cast<const unsigned char*>(texture.data()), texture.size() * 4));
        }
        fprintf(stderr, "%d frame(s) %dx%dx%d: %.3f ms/frame on the GPU (loader + kernels), %.3f ms/frame wall incl. "
                        "upload, G-buffer and frame readback\n", frames, W, H, L, gpu_ms / frames, wall_ms / frames);
        if (ppm) {
            FILE* fp = fopen(ppm, "wb");
            if (!fp) return 1;
            fprintf(fp, "P6\n%d %d\n255\n", W, H);
            for (const par::Color& c : texture) fwrite(&c, 1, 3, fp);
            fclose(fp);
        }
    } catch (const par::Error& e) {
        fprintf(stderr, "par error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
