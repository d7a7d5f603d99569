// par_headless.cpp — headless stand-in for the reference's main loop
// (/root/reference/src/alternative.cpp:502-833) on top of the B200 render path.
// Builds the default scene with Entities::insert (519-599), applies a scripted key sequence
// (641-681) before each frame, renders with par::FrameRenderer (replaces 689-760), draws the
// debug overlay (762-772) and prints one "%03d %016llx" FNV-1a-64 line per frame — the same
// format the SDL stub of oracle/ uses for the real reference, so the two can be diffed.
//
// Default: frames are pipelined (FrameRenderer::submit_frame / wait_frame, two in flight): while
// the GPU renders frame k the host draws the overlay of frame k-1 and hashes it; the overlay's
// record under the cursor comes from the cursor probe, not from a G-buffer readback.
// --sync: the blocking render_frame with the whole G-buffer, the reference's call shape.
//
// Frames after the first send only what moved (FrameRenderer::submit_frame_moved: the scripted keys move
// entity 0, alternative.cpp:641-660); --full-upload re-sends the whole scene every frame instead, as the
// reference re-bins it every frame (alternative.cpp:689-693).
// Frame sink (the step after the path: alternative.cpp:774-788 hands the same bytes to the display, and the
// reference's README shows them as gif.gif): --ppm-seq DIR / --png-seq DIR write every finished frame as
// DIR/frame_NNN.ppm / .png, --gif FILE writes the whole sequence as one animated GIF (include/par/frame_sink.hpp);
// --pitch BYTES renders into host frames with that row pitch, the locked-texture contract of
// alternative.cpp:774-783.
//
//   par_headless [--view W H L] [--frames N] [--script C|D] [--device D] [--ppm out.ppm] [--ppm-seq dir]
//                [--png-seq dir] [--gif out.gif] [--pitch bytes] [--sync] [--full-upload] [--no-hash]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "par/frame_sink.hpp"
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
    const char* ppm_seq = nullptr;
    const char* png_seq = nullptr;
    const char* gif_path = nullptr;
    size_t pitch = 0;
    bool sync = false, hash = true, full_upload = false;
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
        } else if (!strcmp(argv[i], "--ppm-seq") && i + 1 < argc) {
            ppm_seq = argv[++i];
        } else if (!strcmp(argv[i], "--png-seq") && i + 1 < argc) {
            png_seq = argv[++i];
        } else if (!strcmp(argv[i], "--gif") && i + 1 < argc) {
            gif_path = argv[++i];
        } else if (!strcmp(argv[i], "--pitch") && i + 1 < argc) {
            pitch = static_cast<size_t>(atoll(argv[++i]));
        } else if (!strcmp(argv[i], "--full-upload")) {
            full_upload = true;
        } else if (!strcmp(argv[i], "--sync")) {
            sync = true;
        } else if (!strcmp(argv[i], "--no-hash")) {  // frame rate of the loop without the checker's FNV pass
            hash = false;
        } else {
            fprintf(stderr, "usage: %s [--view W H L] [--frames N] [--script C|D] [--device D] [--ppm f] [--ppm-seq dir] "
                            "[--png-seq dir] [--gif f] [--pitch bytes] [--sync] [--full-upload] [--no-hash]\n", argv[0]);
            return 2;
        }
    }
    try {
        par::Entities<> entities;
        par::make_default_scene(entities);
        std::vector<par::Light> lights(1);
        par_light_default(reinterpret_cast<par_light*>(lights.data()));  // alternative.cpp:624-626

        par::FrameRenderer renderer(W, H, L, device);
        const size_t px = static_cast<size_t>(W) * H;
        const size_t row_bytes = static_cast<size_t>(W) * sizeof(par::Color);
        if (pitch && (pitch < row_bytes || pitch % 4)) {
            fprintf(stderr, "--pitch must be a multiple of 4 and at least %zu\n", row_bytes);
            return 2;
        }
        const size_t stride = pitch ? pitch : row_bytes;  // bytes between rows of a host frame
        if (pitch) renderer.set_output_pitch(pitch);
        auto alloc_texture = [&]() {
            void* p = par_alloc_host(stride * H);
            if (!p) throw par::Error(PAR_ERR_OUT_OF_MEMORY, par_last_error());
            memset(p, 0, stride * H);
            return static_cast<par::Color*>(p);
        };
        par::Color* texture[2] = {alloc_texture(), alloc_texture()};
        par::Color* last = texture[0];
        std::vector<par::Color> packed(pitch ? px : 0);  // pitched frames are packed for overlay / hash / PPM
        auto write_ppm = [&](const char* path, const par::Color* frame) {  // (frames reach the sink packed)
            return par::sink::write_ppm(path, frame, W, H, row_bytes);
        };
        std::unique_ptr<par::sink::GifWriter> gif;
        if (gif_path) {
            gif.reset(new par::sink::GifWriter(gif_path, W, H));
            if (!gif->ok()) throw par::Error(PAR_ERR_INVALID_ARG, "cannot write the --gif file");
        }
        auto apply_script = [&](int f) {
            if (script == 'C' || script == 'D') {
                par_aabb* player = reinterpret_cast<par_aabb*>(&entities.aabbs[0]);
                par_light* light = reinterpret_cast<par_light*>(&lights[0]);
                if (int k = script_c_key(f)) par_apply_key(k, player, light);
                if (script == 'D' && f >= 1) par_apply_key('o', player, light);
            }
        };
        auto finish = [&](int f, par::Color* tex, const par::Pixel& under, const par::Light& light) {
            if (pitch) {  // the consumer's view of a pitched frame: its W * 4 bytes of every row
                for (int j = 0; j < H; j++)
                    memcpy(&packed[static_cast<size_t>(j) * W], reinterpret_cast<const char*>(tex) + j * stride, row_bytes);
                tex = packed.data();
            }
            par_draw_overlay_at(W, H, reinterpret_cast<const par_pixel*>(&under),
                                reinterpret_cast<const par_light*>(&light), 0, reinterpret_cast<par_color*>(tex));
            if (hash) printf("%03d %016llx\n", f, fnv1a64(reinterpret_cast<const unsigned char*>(tex), px * 4));
            if (ppm_seq) {
                char name[32];
                snprintf(name, sizeof name, "/frame_%03d.ppm", f);
                if (!write_ppm((std::string(ppm_seq) + name).c_str(), tex)) throw par::Error(PAR_ERR_INVALID_ARG, "cannot write into the --ppm-seq directory");
            }
            if (png_seq) {
                char name[32];
                snprintf(name, sizeof name, "/frame_%03d.png", f);
                if (!par::sink::write_png((std::string(png_seq) + name).c_str(), tex, W, H, row_bytes))
                    throw par::Error(PAR_ERR_INVALID_ARG, "cannot write into the --png-seq directory");
            }
            if (gif && !gif->add_frame(tex, row_bytes)) throw par::Error(PAR_ERR_INVALID_ARG, "cannot write the --gif file");
            last = tex;
        };
        double gpu_ms = 0;
        const auto t0 = std::chrono::steady_clock::now();
        if (sync) {
            std::vector<par::Pixel> gbuf(px);
            for (int f = 0; f < frames; f++) {
                apply_script(f);
                par_stats st{};
                renderer.render_frame(entities, lights, texture[0], gbuf.data(), &st);
                gpu_ms += st.ms_grid_build + st.ms_total;
                finish(f, texture[0], gbuf[0], lights[0]);  // cursor at (0, 0), like the headless reference
            }
        } else {
            renderer.set_cursor(0, 0);
            par::Light light_of[2];
            for (int f = 0; f <= frames; f++) {
                if (f < frames) {
                    apply_script(f);
                    light_of[f & 1] = lights[0];
                    if (full_upload || f == 0)
                        renderer.submit_frame(entities, lights, texture[f & 1]);
                    else  // the scripted keys move entity 0 only (alternative.cpp:641-660)
                        renderer.submit_frame_moved(entities, 0, 1, lights, texture[f & 1]);
                }
                if (f >= 1) {
                    par_stats st{};
                    renderer.wait_frame(&st);
                    gpu_ms += st.ms_total;
                    finish(f - 1, texture[(f - 1) & 1], renderer.cursor_pixel(), light_of[(f - 1) & 1]);
                }
            }
        }
        const double wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "%d frame(s) %dx%dx%d, %s: %.3f ms/frame wall (scene update, upload, render, readback, overlay%s) "
                        "= %.1f frames/s; %.3f ms/frame %s\n", frames, W, H, L,
                sync ? "blocking render_frame + G-buffer"
                     : full_upload ? "pipelined submit_frame/wait_frame (whole scene uploaded per frame) + cursor probe"
                                   : "pipelined submit_frame_moved/wait_frame (16-byte scene update per frame) + cursor probe",
                wall_ms / frames, hash ? ", FNV hash" : "", 1e3 * frames / wall_ms, gpu_ms / frames,
                sync ? "on the GPU (loader + kernels)" : "submit -> frame on the host (latency)");
        if (ppm && !write_ppm(ppm, last)) return 1;
        if (gif && !gif->close()) return 1;
        par_free_host(texture[0]);
        par_free_host(texture[1]);
    } catch (const par::Error& e) {
        fprintf(stderr, "par error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
