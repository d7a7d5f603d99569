// Property test of the shadow-walk restructuring used by k_tile / k_walk on the CPU.
// TEST INFRASTRUCTURE: built and run by tests/test_walk_property.py.
//
// Reference walk (alternative.cpp:399-476, oracle light_visible): per step the six partial advances
// (x, y, z, xy, xz, yz) from the last full position and then the full advance are probed; bins whose
// flat index equals the start bin's are skipped.  Visibility is an OR over the probed bins, so only the
// SET of probed bins matters.  The GPU enumerates, per step, the non-empty subsets of the axes whose
// integer coordinate changed.  Claim: both enumerations visit the same set of flat bin indices, for
// every start / end bin — including ends outside the grid, negative coordinates and flat-index
// aliasing (quirk Q18: out-of-range coordinates wrap into other bins' indices).
//
//   walk_property <trials> <seed>      prints "walks W steps S violations V"
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

static uint64_t s_state;
static inline uint64_t rnd() {
    uint64_t z = (s_state += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline int ri(int lo, int hi) { return lo + (int)(rnd() % (uint64_t)(hi - lo + 1)); }

int main(int argc, char** argv) {
    const long trials = argc > 1 ? atol(argv[1]) : 100000;
    s_state = argc > 2 ? strtoull(argv[2], nullptr, 10) : 1;
    long steps_total = 0, violations = 0;
    std::vector<int> ref, gpu;
    for (long t = 0; t < trials; t++) {
        const int HH = ri(1, 60), HL = ri(1, 60), HW = ri(1, 100);
        const int sx = ri(0, HW - 1), sy = ri(0, HH - 1), sz = ri(0, HL - 1);
        const int slack = (rnd() & 3) ? 0 : 40;  // sometimes the light bin is far outside the grid
        const int lx = ri(-slack, HW - 1 + slack), ly = ri(-slack, HH - 1 + slack), lz = ri(-slack, HL - 1 + slack);
        auto flat = [&](int x, int y, int z) { return x * HH * HL + y * HL + z; };  // alternative.cpp:180-182
        const float dx = (float)lx - (float)sx, dy = (float)ly - (float)sy, dz = (float)lz - (float)sz;
        const float big = fmaxf(fmaxf(fabsf(dx), fabsf(dy)), fabsf(dz));
        const int steps = (int)big;
        if (steps == 0) continue;
        const float stx = dx / big, sty = dy / big, stz = dz / big;
        const int start = flat(sx, sy, sz);
        ref.clear();
        gpu.clear();
        {  // reference order: x, y, z, xy, xz, yz, xyz
            static const int adv[7] = {1, 2, 4, 3, 5, 6, 7};
            float px = (float)sx, py = (float)sy, pz = (float)sz;
            for (int s = 0; s < steps; s++) {
                const float nx = px + stx, ny = py + sty, nz = pz + stz;
                for (int p = 0; p < 7; p++) {
                    const float cx = (adv[p] & 1) ? nx : px, cy = (adv[p] & 2) ? ny : py, cz = (adv[p] & 4) ? nz : pz;
                    const int f = flat((int)cx, (int)cy, (int)cz);
                    if (f != start) ref.push_back(f);
                }
                px = nx;
                py = ny;
                pz = nz;
            }
        }
        {  // GPU enumeration: non-empty subsets of the changed axes, per step
            float px = (float)sx, py = (float)sy, pz = (float)sz;
            int x0 = sx, y0 = sy, z0 = sz;
            for (int s = 0; s < steps; s++) {
                px = px + stx;
                py = py + sty;
                pz = pz + stz;
                const int x1 = (int)px, y1 = (int)py, z1 = (int)pz;
                const int changed = (x1 != x0) | (y1 != y0) << 1 | (z1 != z0) << 2;
                for (int sub = changed; sub; sub = (sub - 1) & changed) {
                    const int f = flat((sub & 1) ? x1 : x0, (sub & 2) ? y1 : y0, (sub & 4) ? z1 : z0);
                    if (f != start) gpu.push_back(f);
                }
                x0 = x1;
                y0 = y1;
                z0 = z1;
            }
        }
        steps_total += steps;
        std::sort(ref.begin(), ref.end());
        ref.erase(std::unique(ref.begin(), ref.end()), ref.end());
        std::sort(gpu.begin(), gpu.end());
        gpu.erase(std::unique(gpu.begin(), gpu.end()), gpu.end());
        if (ref != gpu) {
            if (violations < 5)
                fprintf(stderr, "VIOLATION grid %dx%dx%d start (%d %d %d) end (%d %d %d): %zu vs %zu bins\n", HW, HH, HL, sx,
                        sy, sz, lx, ly, lz, ref.size(), gpu.size());
            violations++;
        }
    }
    printf("walks %ld steps %ld violations %ld\n", trials, steps_total, violations);
    return violations ? 1 : 0;
}
