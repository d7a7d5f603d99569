// Property test of the exact-output shaft cull (pixel-art-raytracer_b200/csrc/shaft.cuh) on the CPU.
// TEST INFRASTRUCTURE: built and run by tests/test_shaft_cull_property.py.
//
// Claim under test: if shaft_may_hit(box, light, origin bounds) is false, then the reference's
// occlusion predicate (oracle: orc_slab_hit_point = ray set-up of alternative.cpp:712-722 +
// AABB::intersect, 40-83) is false for EVERY integer ray origin inside the bounds — so dropping
// the box cannot change any pixel.  The device evaluates the cull with approximate reciprocals
// (__fdividef, <= 2 ulp); here RCP_SKEW scales every reciprocal by (1 + k * 2^-23), k in
// {-4, 0, +4}, to cover that.  Configurations are adversarial: the light is placed on (or within
// a few units of) a line through a group origin and a point of the box, so the shaft grazes it.
//
//   shaft_property <trials> <seed>      prints "culled C kept K violations V"
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define __device__
#define __forceinline__ inline
static float g_rcp_skew = 1.f;
static inline float __fdividef(float a, float b) { return (a / b) * g_rcp_skew; }
#include "shaft_host.h"  // shaft.cuh without its CUDA include (written by the test)

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
static inline uint64_t rnd() {  // splitmix64
    uint64_t z = (s_state += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline int ri(int lo, int hi) { return lo + (int)(rnd() % (uint64_t)(hi - lo + 1)); }  // inclusive

int main(int argc, char** argv) {
    const long trials = argc > 1 ? atol(argv[1]) : 200000;
    s_state = argc > 2 ? strtoull(argv[2], nullptr, 10) : 1;
    long culled = 0, kept = 0, violations = 0;
    for (long t = 0; t < trials; t++) {
        // group origin bounds: up to a 40-column tile, a z-group of up to 40 and the y range that goes with it
        int ol[3], oh[3];
        ol[0] = ri(-40, 2000);
        oh[0] = ol[0] + ri(0, 39);
        ol[1] = ri(-300, 900);
        oh[1] = ol[1] + ri(0, 79);
        ol[2] = ri(-80, 1200);
        oh[2] = ol[2] + ri(0, 39);
        orc_aabb b;
        const int mode = (int)(rnd() % 4);
        const int reach = mode == 0 ? 60 : mode == 1 ? 300 : 1500;  // near boxes matter most
        b.px = (int16_t)(ol[0] + ri(-reach, reach));
        b.py = (int16_t)(ol[1] + ri(-reach, reach));
        b.pz = (int16_t)(ol[2] + ri(-reach, reach));
        b.ex = (int16_t)ri(0, 20);
        b.ey = (int16_t)ri(0, 20);
        b.ez = (int16_t)ri(0, 20);
        b.pad[0] = b.pad[1] = 0;
        // light: on a line from a group origin through a point of the box (grazing), jittered
        int o[3], p[3];
        for (int a = 0; a < 3; a++) o[a] = ri(ol[a], oh[a]);
        p[0] = b.px + ri(0, b.ex);
        p[1] = b.py + ri(0, b.ey);
        p[2] = b.pz + ri(0, b.ez);
        orc_light lt;
        const int k = ri(1, 4), jit = (int)(rnd() % 3) == 0 ? 0 : ri(1, 6);
        int Lc[3];
        for (int a = 0; a < 3; a++) Lc[a] = p[a] + (k - 1) * (p[a] - o[a]) + ri(-jit, jit);
        if ((rnd() & 7) == 0)  // sometimes anywhere
            for (int a = 0; a < 3; a++) Lc[a] = ol[a] + ri(-1500, 1500);
        bool in_range = true;
        for (int a = 0; a < 3; a++) in_range = in_range && Lc[a] >= -32768 && Lc[a] <= 32767;
        if (!in_range) continue;
        lt.x = (int16_t)Lc[0];
        lt.y = (int16_t)Lc[1];
        lt.z = (int16_t)Lc[2];
        lt.radius = 10;

        const float lo[3] = {(float)b.px, (float)b.py, (float)b.pz};
        const float hi[3] = {(float)(b.px + b.ex), (float)(b.py + b.ey), (float)(b.pz + b.ez)};
        const float Lf[3] = {(float)lt.x, (float)lt.y, (float)lt.z};
        const float olf[3] = {(float)ol[0], (float)ol[1], (float)ol[2]};
        const float ohf[3] = {(float)oh[0], (float)oh[1], (float)oh[2]};
        bool may = false;
        for (int s = -1; s <= 1; s++) {  // any evaluation the device could make
            g_rcp_skew = 1.f + (float)(4 * s) * 1.1920929e-7f;
            may = may || par::shaft_may_hit(lo, hi, Lf, olf, ohf);
        }
        if (may) {
            kept++;
            continue;
        }
        culled++;
        // every corner, every edge midpoint-ish and random interior origins must miss in the reference
        for (int c = 0; c < 8 + 40; c++) {
            int q[3];
            for (int a = 0; a < 3; a++)
                q[a] = c < 8 ? ((c >> a) & 1 ? oh[a] : ol[a]) : ri(ol[a], oh[a]);
            if (orc_slab_hit_point(&b, q[0], q[1], q[2], &lt)) {
                if (violations < 5)
                    fprintf(stderr, "VIOLATION box (%d %d %d)+(%d %d %d) light (%d %d %d) origin (%d %d %d) bounds [%d %d %d]-[%d %d %d]\n",
                            b.px, b.py, b.pz, b.ex, b.ey, b.ez, lt.x, lt.y, lt.z, q[0], q[1], q[2], ol[0], ol[1],
                            ol[2], oh[0], oh[1], oh[2]);
                violations++;
                break;
            }
        }
    }
    printf("culled %ld kept %ld violations %ld\n", culled, kept, violations);
    return violations ? 1 : 0;
}
