// shaft.cuh — exact-output cull of candidate occluders for a whole group of shadow rays
// (tile.cu; compiled for the host by tests/test_shaft_cull_property.py).
#pragma once
#include "par_device.cuh"

namespace par {

// Exact-output "shaft" cull of a candidate box for a whole pixel group: all ray origins of the
// group lie in the integer box [ol, oh] (per axis), all rays go through the light L.  In real
// arithmetic a pixel's line hits the box iff the per-axis parameter intervals
// [min, max]{(lo_a - o_a)/(L_a - o_a), (hi_a - o_a)/(L_a - o_a)} intersect (the reference's
// slab test up to the positive scale |L - o|_1).  Each endpoint is monotonic in o_a, so its hull
// over the group is attained at the interval ends; if the hulls of the three axes do not
// intersect — with a 1e-4 relative margin, three orders of magnitude above the fp32 error of
// the reference's formula — no pixel of the group can pass the reference's test and the box is
// dropped.  Axes on which some pixel may have a zero direction component (0 in [L-oh, L-ol])
// impose no constraint, which also covers every NaN/inf case (quirk Q13) conservatively.
// The part that does not depend on the box, once per (group, light): the reciprocals of the light's distance to
// both ends of the origin interval, and the axes that impose no constraint.
__device__ __forceinline__ unsigned shaft_prepare(const float L[3], const float ol[3], const float oh[3], float rl[3],
                                                  float rh[3]) {
    unsigned free_axes = 0u;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float dl = L[a] - oh[a], dh = L[a] - ol[a];
        if (dl <= 0.f && dh >= 0.f) {
            free_axes |= 1u << a;
            rl[a] = rh[a] = 0.f;
        } else {  // approximate reciprocals (2 ulp) are plenty under the 1e-4 margin
            rl[a] = __fdividef(1.f, dl);
            rh[a] = __fdividef(1.f, dh);
        }
    }
    return free_axes;
}

__device__ __forceinline__ bool shaft_may_hit_prepared(const float lo[3], const float hi[3], const float ol[3],
                                                       const float oh[3], const float rl[3], const float rh[3],
                                                       unsigned free_axes) {
    float smin = -INFINITY, smax = INFINITY;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if ((free_axes >> a) & 1u) continue;
        const float v0 = (lo[a] - ol[a]) * rh[a], v1 = (lo[a] - oh[a]) * rl[a];
        const float v2 = (hi[a] - ol[a]) * rh[a], v3 = (hi[a] - oh[a]) * rl[a];
        smin = fmaxf(smin, fminf(fminf(v0, v1), fminf(v2, v3)));
        smax = fminf(smax, fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)));
    }
    return !(smin > smax + 1e-4f * (1.f + fabsf(smin) + fabsf(smax)));
}

__device__ __forceinline__ bool shaft_may_hit(const float lo[3], const float hi[3], const float L[3],
                                              const float ol[3], const float oh[3]) {
    float rl[3], rh[3];
    const unsigned free_axes = shaft_prepare(L, ol, oh, rl, rh);
    return shaft_may_hit_prepared(lo, hi, ol, oh, rl, rh, free_axes);
}

}  // namespace par
