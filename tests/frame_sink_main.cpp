// Harness of include/par/frame_sink.hpp (PPM / PNG / animated GIF writers of the headless host).
// TEST INFRASTRUCTURE: built with g++ and run by tests/test_frame_sink.py, which re-creates the same synthetic
// frames with numpy and decodes the files with an independent decoder.
//
//   frame_sink_main <dir> <W> <H> <frames> <pitch bytes> <colours>
// pixel (x, y) of frame f has palette index (7x + 13y + 31f + (x*y) % 11) % colours, palette entry i is
// (37i % 256, 91i % 256, 53i % 256, 255); rows lie `pitch` bytes apart with 0xEE in the padding.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "par/frame_sink.hpp"

int main(int argc, char** argv) {
    if (argc < 7) return 2;
    const std::string dir = argv[1];
    const int W = atoi(argv[2]), H = atoi(argv[3]), frames = atoi(argv[4]);
    const size_t pitch = static_cast<size_t>(atoll(argv[5]));
    const int colours = atoi(argv[6]);
    if (pitch < static_cast<size_t>(W) * 4) return 2;
    par::sink::GifWriter gif((dir + "/seq.gif").c_str(), W, H, 4);
    std::vector<unsigned char> frame(pitch * static_cast<size_t>(H));
    for (int f = 0; f < frames; f++) {
        std::fill(frame.begin(), frame.end(), static_cast<unsigned char>(0xEE));
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                const int i = (7 * x + 13 * y + 31 * f + (x * y) % 11) % colours;
                unsigned char* p = &frame[static_cast<size_t>(y) * pitch + 4 * static_cast<size_t>(x)];
                p[0] = static_cast<unsigned char>(37 * i % 256);
                p[1] = static_cast<unsigned char>(91 * i % 256);
                p[2] = static_cast<unsigned char>(53 * i % 256);
                p[3] = 255;
            }
        char name[64];
        snprintf(name, sizeof name, "/frame_%03d", f);
        if (!par::sink::write_ppm((dir + name + ".ppm").c_str(), frame.data(), W, H, pitch)) return 1;
        if (!par::sink::write_png((dir + name + ".png").c_str(), frame.data(), W, H, pitch)) return 1;
        if (!gif.add_frame(frame.data(), pitch)) return 1;
        printf("%d %d\n", f, gif.last_frame_exact() ? 1 : 0);
    }
    return gif.close() && gif.frames() == frames ? 0 : 1;
}
