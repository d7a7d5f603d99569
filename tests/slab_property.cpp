// Property test of the three slab-test variants of k_tile (pixel-art-raytracer_b200/csrc/tile.cu)
// on the CPU.  TEST INFRASTRUCTURE: built and run by tests/test_slab_variants_property.py, which
// extracts std_min/std_max (par_device.cuh) and slab_hit_exact / slab_hit_fast / slab_hit_near_far
// (tile.cu) verbatim into slab_host.h.
//
// Claims under test, against the oracle's predicate (orc_slab_hit_point: ray set-up of
// alternative.cpp:712-722 + AABB::intersect 40-83):
//   1. slab_hit_exact equals it for EVERY input, zero direction components (inf / NaN, Q13) included;
//   2. slab_hit_fast (fminf/fmaxf) equals it whenever no direction component is zero;
//   3. slab_hit_near_far, given the box as (near, far) corners for the ray's sign octant, equals it
//      whenever no direction component is zero.
//
//   slab_property <trials> <seed>      prints "trials T zero_dir Z hits H violations V"
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define __device__
#define __forceinline__ inline
struct float4 {
    float x, y, z, w;
};
namespace par {
#include "slab_host.h"
}

extern "C" {
typedef struct {
    int16_t px, py, pz, ex, ey, ez, pad[2];
} orc_aabb;
typedef struct {
    int16_t x, y, z, radius;
} orc_light;
int orc_slab_hit_point(const orc_aabb* box, int ox, int oy, int oz, const orc_light* lt);
}

static uint64_t s_state;
static inline uint64_t rnd() {
    uint64_t z = (s_state += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline int ri(int lo, int hi) { return lo + (int)(rnd() % (uint64_t)(hi - lo + 1)); }

int main(int argc, char** argv) {
    const long trials = argc > 1 ? atol(argv[1]) : 1000000;
    s_state = argc > 2 ? strtoull(argv[2], nullptr, 10) : 1;
    long zero_dir = 0, hits = 0, violations = 0;
    for (long t = 0; t < trials; t++) {
        int o[3], L[3];
        orc_aabb b;
        const int span = (rnd() & 3) ? 120 : 3000;  // mostly close quarters: grazing hits, shared coordinates
        for (int a = 0; a < 3; a++) o[a] = ri(-500, 4000);
        b.px = (int16_t)(o[0] + ri(-span, span));
        b.py = (int16_t)(o[1] + ri(-span, span));
        b.pz = (int16_t)(o[2] + ri(-span, span));
        b.ex = (int16_t)ri(0, 20);
        b.ey = (int16_t)ri(0, 20);
        b.ez = (int16_t)ri(0, 20);
        b.pad[0] = b.pad[1] = 0;
        const bool aimed = (rnd() & 1) != 0;  // the light on a line through the box (jittered): hits and grazes
        const int k = ri(1, 3), jit = ri(0, 3);
        for (int a = 0; a < 3; a++) {
            const int p = (&b.px)[a] + ri(0, (&b.ex)[a]);
            L[a] = aimed ? p + (k - 1) * (p - o[a]) + ri(-jit, jit) : o[a] + ri(-span, span);
            if (L[a] < -32768 || L[a] > 32767) L[a] = o[a] + 1;
            if ((rnd() & 7) == 0) L[a] = o[a];                       // zero direction component (Q13)
            if ((rnd() & 7) == 0) L[a] = (&b.px)[a] + ((rnd() & 1) ? (&b.ex)[a] : 0);  // light on a box plane
        }
        if (L[0] == o[0] && L[1] == o[1] && L[2] == o[2]) continue;  // light inside the pixel: len = 0 (not reachable: y+z)
        orc_light lt = {(int16_t)L[0], (int16_t)L[1], (int16_t)L[2], 10};
        const int want = orc_slab_hit_point(&b, o[0], o[1], o[2], &lt);
        hits += want;

        // the product's ray set-up (tile.cu, phase F): same operations as the reference
        float tx = (float)(L[0] - o[0]), ty = (float)(L[1] - o[1]), tz = (float)(L[2] - o[2]);
        const float len = fabsf(tx) + fabsf(ty) + fabsf(tz);
        tx = tx / len;
        ty = ty / len;
        tz = tz / len;
        const float ix = 1.f / tx, iy = 1.f / ty, iz = 1.f / tz;
        const float ox = (float)o[0], oy = (float)o[1], oz = (float)o[2];
        const float4 lo = {(float)b.px, (float)b.py, (float)b.pz, 0.f};
        const float4 hi = {(float)(b.px + b.ex), (float)(b.py + b.ey), (float)(b.pz + b.ez), 0.f};
        bool bad = par::slab_hit_exact(lo, hi, ox, oy, oz, ix, iy, iz) != (want != 0);
        const bool any_zero = L[0] == o[0] || L[1] == o[1] || L[2] == o[2];
        zero_dir += any_zero;
        if (!any_zero) {
            bad = bad || par::slab_hit_fast(lo, hi, ox, oy, oz, ix, iy, iz) != (want != 0);
            const bool nx = L[0] < o[0], ny = L[1] < o[1], nz = L[2] < o[2];  // negative component: near corner = hi
            const float4 nr = {nx ? hi.x : lo.x, ny ? hi.y : lo.y, nz ? hi.z : lo.z, 0.f};
            const float4 fr = {nx ? lo.x : hi.x, ny ? lo.y : hi.y, nz ? lo.z : hi.z, 0.f};
            bad = bad || par::slab_hit_near_far(nr, fr, ox, oy, oz, ix, iy, iz) != (want != 0);
        }
        if (bad) {
            if (violations < 5)
                fprintf(stderr, "VIOLATION box (%d %d %d)+(%d %d %d) origin (%d %d %d) light (%d %d %d) want %d\n", b.px, b.py,
                        b.pz, b.ex, b.ey, b.ez, o[0], o[1], o[2], L[0], L[1], L[2], want);
            violations++;
        }
    }
    printf("trials %ld zero_dir %ld hits %ld violations %ld\n", trials, zero_dir, hits, violations);
    return violations ? 1 : 0;
}
