// shade.cu — fused shading + shadow rays + RGBA8 quantise/pack: replaces the shading loop
// (/root/reference/src/alternative.cpp:702-760), trace_hash_for_light (399-500),
// AABB::intersect (40-83), Vector::normalize (sprites.hpp:28-35) and Color::operator*
// (sprites.hpp:8-16), generalised to N lights:
//     acc = sum over visible lights of max(0, n . t_l);   out = color * min(1, acc + ambient)
//
// What makes it fast without changing a bit of the output:
//  * The reference walks the grid once per PIXEL; but the sequence of probed bins depends only
//    on (start bin, light bin), and every hit pixel of a 40x40 screen tile starts in bin
//    (tile x, tile y, z/40) (quirk Q11).  One CTA owns a tile, groups its pixels by z/40 and
//    walks once per (group, light): the fp32 position chain is accumulated sequentially
//    exactly as the reference does (quirk Q15), the 7 probes of a step collapse to the
//    distinct bins among them (probing a bin twice cannot change an OR), and the occupied
//    bins' boxes are gathered into shared memory.
//  * The per-pixel work is then a loop of slab tests over that shared list, with the
//    reference's unbounded-line semantics (Q14), argument-order-exact min/max (Q13), the
//    start-bin skip (Q16) applied at gather time and the self-entity skip (Q17) per lane.
//  * A pixel whose Lambert term is 0 for a light never issues the shadow query: visible or
//    not, it adds +0 (quirk Q19).
//  * Four neighbouring lanes merge their RGBA8 pixels with shuffles into one 16-byte store.
// All fp32 arithmetic is IEEE round-to-nearest with no FMA contraction (-fmad=false).
#include "par_kernels.cuh"

namespace par {

constexpr int kListCap = 1024;  // boxes per shared-memory window (2 x float4 each = 32 KB)
constexpr int kNoGroup = 0x7fffffff;

// alternative.cpp:40-83 on a box given as float lo/hi corners.  (float)(int - int) of
// 16-bit operands equals the float difference exactly, so the int subtract + convert of the
// reference is one FADD here.
__device__ __forceinline__ bool slab_hit(const float4 lo, const float4 hi, float ox, float oy,
                                         float oz, float ix, float iy, float iz) {
    float x1 = (lo.x - ox) * ix, x2 = (hi.x - ox) * ix;
    float tmin = std_min(x1, x2);
    float tmax = std_max(x1, x2);
    float y1 = (lo.y - oy) * iy, y2 = (hi.y - oy) * iy;
    tmin = std_max(tmin, std_min(y1, y2));
    tmax = std_min(tmax, std_max(y1, y2));
    float z1 = (lo.z - oz) * iz, z2 = (hi.z - oz) * iz;
    tmin = std_max(tmin, std_min(z1, z2));
    tmax = std_min(tmax, std_max(z1, z2));
    return tmax >= tmin;
}

// Block-wide exclusive scan of one int per thread (blockDim = kTileThreads = 10 warps).
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int* s_total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        int x = threadIdx.x < kTileThreads / 32 ? s_warp[threadIdx.x] : 0;
        int xi = x;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, xi, o);
            if (lane >= o) xi += t;
        }
        if (threadIdx.x < kTileThreads / 32) s_warp[threadIdx.x] = xi - x;
        if (threadIdx.x == kTileThreads / 32 - 1) *s_total = xi;
    }
    __syncthreads();
    return s_warp[w] + incl - v;
}

__global__ void __launch_bounds__(kTileThreads, 3)
k_shade(const __grid_constant__ ShadeParams p) {
    __shared__ float4 s_lo[kListCap];  // box min corner; .w carries the entity index bits
    __shared__ float4 s_hi[kListCap];  // box max corner
    __shared__ int s_warp[kTileThreads / 32];
    __shared__ int s_total;
    __shared__ int s_group;
    __shared__ float s_seed[3];

    const ViewDims& d = p.d;
    const int tid = threadIdx.x;
    const int bx = blockIdx.x % d.HW;
    const int ty = p.tile_row_first + blockIdx.x / d.HW;
    const int col = tid % kBin, rsub = tid / kBin;
    const int i = bx * kBin + col;
    const int ra = max(ty * kBin, d.row0), rb = min(ty * kBin + kBin, d.row1);

    // Per-thread pixel state: 5 pixels of one column, 8 rows apart.
    int gz[kTileRowsPerThread];     // start bin z of the pixel = world z / 40, kNoGroup for none
    float acc[kTileRowsPerThread];  // running sum of Lambert terms of the visible lights
#pragma unroll
    for (int m = 0; m < kTileRowsPerThread; m++) {
        const int j = ty * kBin + rsub + 8 * m;
        gz[m] = kNoGroup;
        acc[m] = 0.f;
        if (j >= ra && j < rb) {
            const int4 g = p.gbuf[(size_t)j * d.W + i];
            if (g.w >= 0) gz[m] = g.z / kBin;  // ray_bin_z, alternative.cpp:727
        }
    }

    int last_group = -0x7fffffff - 1;
    for (;;) {
        // ---- next group: the smallest start-bin z not yet processed in this tile ----
        if (tid == 0) s_group = kNoGroup;
        __syncthreads();
        int mine = kNoGroup;
#pragma unroll
        for (int m = 0; m < kTileRowsPerThread; m++)
            if (gz[m] > last_group) mine = min(mine, gz[m]);
        mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, 16));
        mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, 8));
        mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, 4));
        mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, 2));
        mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, 1));
        if ((tid & 31) == 0 && mine != kNoGroup) atomicMin(&s_group, mine);
        __syncthreads();
        const int group = s_group;
        if (group == kNoGroup) break;
        last_group = group;

        // start bin of every pixel of the group (alternative.cpp:724-727, quirk Q11)
        const int start = flat_bin(d, bx, ty, group);

        for (int l = 0; l < p.n_lights; l++) {
            const short4 lt = p.lights[l];
            // light bin, alternative.cpp:729-732 ('/' truncates toward zero)
            const int lbx = lt.x / kBin, lby = (d.H - lt.y - lt.z) / kBin, lbz = lt.z / kBin;
            // walk set-up, alternative.cpp:406-430
            const float dx = (float)lbx - (float)bx, dy = (float)lby - (float)ty,
                        dz = (float)lbz - (float)group;
            const float big = fmaxf(fmaxf(fabsf(dx), fabsf(dy)), fabsf(dz));
            const int steps = (int)big;  // 0 when big < 1 (and the NaN steps are never used)
            const float stx = dx / big, sty = dy / big, stz = dz / big;

            unsigned shadowed = 0;  // bit m: pixel m already found an occluder for this light
            if (tid == 0) {
                s_seed[0] = (float)bx;
                s_seed[1] = (float)ty;
                s_seed[2] = (float)group;
            }
            // chunks of blockDim steps (one step per thread); almost always exactly one
            for (int chunk0 = 0; chunk0 == 0 || chunk0 < steps; chunk0 += kTileThreads) {
                __syncthreads();
                // ---- P1: this thread's step: positions, distinct probed bins, entry counts ----
                const int k = chunk0 + tid;
                unsigned probe[7];
                int n_probe = 0, my_entries = 0;
                float qx = 0.f, qy = 0.f, qz = 0.f;
                if (k < steps) {
                    // sequential fp32 accumulation from the chunk seed (quirk Q15)
                    float px = s_seed[0], py = s_seed[1], pz = s_seed[2];
                    for (int s = 0; s < tid; s++) {
                        px = px + stx;
                        py = py + sty;
                        pz = pz + stz;
                    }
                    qx = px + stx;
                    qy = py + sty;
                    qz = pz + stz;
                    const int x0 = (int)px, y0 = (int)py, z0 = (int)pz;
                    const int x1 = (int)qx, y1 = (int)qy, z1 = (int)qz;
                    const int cx = x1 != x0, cy = y1 != y0, cz = z1 != z0;
                    // The 7 probes of the step are the bins {x0|x1} x {y0|y1} x {z0|z1} minus
                    // "all old"; the all-old bin was the previous step's last probe (or the
                    // start bin, which is skipped anyway: quirk Q16).
#pragma unroll
                    for (int mask = 1; mask < 8; mask++) {
                        if (((mask & 1) && !cx) || ((mask & 2) && !cy) || ((mask & 4) && !cz))
                            continue;  // same bin as the probe with that bit cleared
                        const int f = flat_bin(d, (mask & 1) ? x1 : x0, (mask & 2) ? y1 : y0,
                                               (mask & 4) ? z1 : z0);
                        if (f == start || f < 0 || f >= d.V) continue;  // Q16 / Q18
                        const int c = p.cnt[f] & (kSlots - 1);
                        if (c) {
                            probe[n_probe++] = (unsigned)f << 3 | (unsigned)c;
                            my_entries += c;
                        }
                    }
                }
                const int my_off = block_exclusive_scan(my_entries, s_warp, &s_total);
                const int total = s_total;
                if (tid == kTileThreads - 1 && k < steps) {  // seed of the next chunk
                    s_seed[0] = qx;
                    s_seed[1] = qy;
                    s_seed[2] = qz;
                }
                const bool last_chunk = chunk0 + kTileThreads >= steps;

                // ---- windows of at most kListCap boxes (almost always exactly one) ----
                for (int w0 = 0; w0 == 0 || w0 < total; w0 += kListCap) {
                    if (w0) __syncthreads();  // previous window fully consumed
                    // P2: gather this thread's boxes that fall into the window
                    int pos = my_off;
                    for (int q = 0; q < n_probe; q++) {
                        const int f = probe[q] >> 3, c = probe[q] & 7;
                        for (int s = 0; s < c; s++, pos++) {
                            if (pos < w0 || pos >= w0 + kListCap) continue;
                            const int ent = p.ids[f * kSlots + s];
                            const Box b = unpack_box(p.boxes[ent]);
                            s_lo[pos - w0] = make_float4((float)b.px, (float)b.py, (float)b.pz,
                                                         __int_as_float(ent));
                            s_hi[pos - w0] = make_float4((float)(b.px + b.ex), (float)(b.py + b.ey),
                                                         (float)(b.pz + b.ez), 0.f);
                        }
                    }
                    __syncthreads();
                    const int n = min(total - w0, kListCap);
                    const bool last_window = last_chunk && (w0 + kListCap >= total);

                    // ---- P3: per-pixel shading against the window ----
#pragma unroll
                    for (int m = 0; m < kTileRowsPerThread; m++) {
                        if (gz[m] != group) continue;
                        const int j = ty * kBin + rsub + 8 * m;
                        const int4 g = p.gbuf[(size_t)j * d.W + i];
                        const float* nrm = p.atlas_normal + ((g.w >> 10) * kTexels + (g.w & 1023)) * 3;
                        const float nx = __ldg(nrm), ny = __ldg(nrm + 1), nz = __ldg(nrm + 2);
                        // towards_light, L1-normalised (alternative.cpp:711-715, sprites.hpp:28-35)
                        float tx = (float)(lt.x - i), tyv = (float)(lt.y - g.y), tz = (float)(lt.z - g.z);
                        const float len = fabsf(tx) + fabsf(tyv) + fabsf(tz);
                        tx = tx / len;
                        tyv = tyv / len;
                        tz = tz / len;
                        // alternative.cpp:745-747
                        const float lam = std_max(0.f, nx * tx + ny * tyv + nz * tz);
                        if (!(lam > 0.f)) continue;  // quirk Q19: adds +0 whether visible or not
                        if (!(shadowed >> m & 1) && n > 0) {
                            // Ray, alternative.cpp:717-722
                            const float ix = 1.f / tx, iy = 1.f / tyv, iz = 1.f / tz;
                            const float ox = (float)(short)i, oy = (float)(short)g.y,
                                        oz = (float)(short)g.z;
                            bool hit = false;
                            for (int e = 0; e < n; e++) {
                                const float4 lo = s_lo[e];
                                if (__float_as_int(lo.w) == g.x) continue;  // quirk Q17
                                if (slab_hit(lo, s_hi[e], ox, oy, oz, ix, iy, iz)) {
                                    hit = true;
                                    break;
                                }
                            }
                            if (hit) shadowed |= 1u << m;
                        }
                        if (last_window && !(shadowed >> m & 1)) acc[m] = acc[m] + lam;
                    }
                }
            }
        }
    }

    // ---- quantise + pack + 16-byte stores (alternative.cpp:735/757-758, sprites.hpp:8-16) ----
#pragma unroll
    for (int m = 0; m < kTileRowsPerThread; m++) {
        const int j = ty * kBin + rsub + 8 * m;
        const bool valid = j >= ra && j < rb;  // uniform across the 4 lanes of a quad
        unsigned rgba = 0;
        if (valid) {
            const int4 g = p.gbuf[(size_t)j * d.W + i];
            uchar4 c = make_uchar4(127, 127, 127, 0);  // miss colour, alternative.cpp:281
            if (g.w >= 0) c = p.palette[p.atlas_color[(g.w >> 10) * kTexels + (g.w & 1023)]];
            const float f = std_min(1.f, acc[m] + p.ambient);
            const unsigned r = (unsigned char)((float)c.x * f);
            const unsigned gg = (unsigned char)((float)c.y * f);
            const unsigned b = (unsigned char)((float)c.z * f);
            rgba = r | gg << 8 | b << 16 | (unsigned)c.w << 24;
        }
        // lanes 4q..4q+3 hold 4 consecutive pixels of one row (40 and 32 are multiples of 4)
        const unsigned v1 = __shfl_down_sync(0xffffffffu, rgba, 1);
        const unsigned v2 = __shfl_down_sync(0xffffffffu, rgba, 2);
        const unsigned v3 = __shfl_down_sync(0xffffffffu, rgba, 3);
        if (valid && (tid & 3) == 0)
            *reinterpret_cast<uint4*>(&p.out[(size_t)j * d.W + i]) = make_uint4(rgba, v1, v2, v3);
    }
}

cudaError_t launch_shade(const ShadeParams& p, cudaStream_t s) {
    const ViewDims& d = p.d;
    int tile_rows = (d.row1 + kBin - 1) / kBin - d.row0 / kBin;
    if (tile_rows <= 0) return cudaSuccess;
    k_shade<<<tile_rows * d.HW, kTileThreads, 0, s>>>(p);
    return cudaGetLastError();
}

}  // namespace par
