// walks.cu — the grid-walk half of trace_hash_for_light (/root/reference/src/alternative.cpp:399-476)
// as its own throughput kernel: ONE THREAD per (tile, z-group, light).
//
// The sequence of bins a shadow ray probes depends only on its start bin and the light's bin,
// and every hit pixel of a 40x40 screen tile starts in bin (tile x, tile y, z/40) (quirk Q11).
// So the walk is done once per (tile, group, light) — here, with hundreds of thousands of
// independent threads so that its dependent chains (sequential fp32 position accumulation, quirk Q15;
// occupancy -> entity -> box loads) are hidden by thread-level parallelism instead of stalling
// a whole shading CTA.  The result — the de-duplicated boxes that can shadow some pixel of the
// group (shaft cull, see shade.cu) — goes to a pool in HBM; the shade kernel fetches the lists.
// Anything that does not fit the fixed per-thread buffers is marked "not available" and the shade
// kernel walks that case itself, so no scene is ever rejected.
#include "par_kernels.cuh"
#include "shaft.cuh"

namespace par {

constexpr int kWalkThreads = 128;
constexpr int kWalkOccCap = 32;  // occupied bins buffered between two gathers of one walk

// Shared memory: per-thread lists, entry-major so that a warp's accesses never conflict.
struct WalkSmem {
    unsigned occ[kWalkOccCap][kWalkThreads];  // count << 25 | flat bin
    int kept[kWalkListCap][kWalkThreads];     // entity indices of the boxes that survive the cull
};

// One THREAD per (tile, z-group, light); the 32 lanes of a warp are 32 neighbouring tiles of one
// tile row (same group rank, same light), whose walks have nearly the same length and direction.
// The sequential fp32 position chain (quirk Q15) costs one FADD per axis per step and lane — the
// warp advances 32 walks at once.
__global__ void __launch_bounds__(kWalkThreads)
k_walk(const __grid_constant__ WalkParams p) {
    __shared__ WalkSmem s;
    const ViewDims& d = p.d;
    const int tid = threadIdx.x, lane = tid & 31;

    // thread -> (tile row, group rank k, light l, tile column bx), bx fastest
    const int hw_pad = (d.HW + 31) & ~31;  // whole warps per tile row
    long long g = (long long)blockIdx.x * kWalkThreads + tid;
    const int bx = (int)(g % hw_pad);
    g /= hw_pad;
    const int l = (int)(g % p.n_lights);
    g /= p.n_lights;
    const int k = (int)(g % kMaxGroups);
    int first_row, n_rows;
    owned_tile_rows(d, first_row, n_rows);
    const bool in_grid = bx < d.HW && g / kMaxGroups < n_rows;  // the launch is rounded up to whole CTAs
    const int ty = p.tile_row_first + (int)(g / kMaxGroups) * max(d.stripe_n, 1);
    const int tile = in_grid ? ty * d.HW + bx : 0;
    const bool active = in_grid && k < p.tile_ngroups[tile];

    int n_occ = 0, n_kept = 0;
    bool overflow = false;
    if (active) {
        const GroupMeta gm = p.groups[(size_t)tile * kMaxGroups + k];
        const int group = gm.z;
        const short4 lt = p.lights[l];
        // light bin, alternative.cpp:729-732 ('/' truncates toward zero); walk set-up, 406-430
        const int lbx = lt.x / kBin, lby = (d.H - lt.y - lt.z) / kBin, lbz = lt.z / kBin;
        const int start = flat_bin(d, bx, ty, group);
        const float dx = (float)lbx - (float)bx, dy = (float)lby - (float)ty, dz = (float)lbz - (float)group;
        const float big = fmaxf(fmaxf(fabsf(dx), fabsf(dy)), fabsf(dz));
        const int steps = (int)big;  // 0 when big < 1 (then the NaN step is never used)
        const float sx = dx / big, sy = dy / big, sz = dz / big;
        const int sxy = d.HH * d.HL;

        float org_lo[3], org_hi[3];
        bool can_cull = !(p.debug_flags & 1);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            org_lo[a] = (float)gm.omin[a];
            org_hi[a] = (float)gm.omax[a];
            can_cull = can_cull && gm.omin[a] >= -32768 && gm.omax[a] <= 32767;  // origins are cast to short
        }
        const float lp[3] = {(float)lt.x, (float)lt.y, (float)lt.z};

        // ---- walk (sequential fp32 accumulation from the start bin, quirk Q15) in stretches that
        //      fill the occupied-bin buffer, each followed by a gather of the buffered bins ----
        float px = (float)bx, py = (float)ty, pz = (float)group;
        int x0 = bx, y0 = ty, z0 = group;
        int kk = 0;
        while (kk < steps && !overflow) {
            n_occ = 0;
            for (; kk < steps && n_occ + 7 <= kWalkOccCap; kk++) {
                px = px + sx;
                py = py + sy;
                pz = pz + sz;
                const int x1 = (int)px, y1 = (int)py, z1 = (int)pz;
                // distinct bins among the step's 7 probes = non-empty subsets of the changed axes
                // (see shade.cu); start-bin skip Q16, flat-index bounds Q18
                const int changed = (x1 != x0) | (y1 != y0) << 1 | (z1 != z0) << 2;
                const int fx0 = x0 * sxy, fx1 = x1 * sxy, fy0 = y0 * d.HL, fy1 = y1 * d.HL;
                for (int sub = changed; sub; sub = (sub - 1) & changed) {
                    const int f = ((sub & 1) ? fx1 : fx0) + ((sub & 2) ? fy1 : fy0) + ((sub & 4) ? z1 : z0);
                    if (f == start || (unsigned)f >= (unsigned)d.V) continue;
                    const unsigned c = (__ldg(&p.occ4[f >> 3]) >> ((f & 7) * 4)) & 7;
                    if (c) s.occ[n_occ++][tid] = c << 25 | (unsigned)f;
                }
                x0 = x1;
                y0 = y1;
                z0 = z1;
            }
            // gather: entity -> box -> shaft cull -> de-duplicate against the kept list
            for (int i = 0; i < n_occ && !overflow; i++) {
                const unsigned desc = s.occ[i][tid];
                const int c = (desc >> 25) & 7;
                const int* slot = p.ids + (size_t)(desc & 0x1ffffffu) * kSlots;
                for (int j = 0; j < c; j++) {
                    const int ent = slot[j];
                    const Box b = unpack_box(p.boxes[ent]);
                    if (can_cull) {
                        const float blo[3] = {(float)b.px, (float)b.py, (float)b.pz};
                        const float bhi[3] = {(float)(b.px + b.ex), (float)(b.py + b.ey), (float)(b.pz + b.ez)};
                        if (!shaft_may_hit(blo, bhi, lp, org_lo, org_hi)) continue;
                    }
                    bool dup = false;
                    for (int q = 0; q < min(n_kept, kWalkListCap); q++) dup = dup || s.kept[q][tid] == ent;
                    if (dup) continue;
                    if (n_kept < kWalkListCap) s.kept[n_kept][tid] = ent;
                    n_kept++;
                }
                overflow = n_kept > kWalkListCap;
            }
        }
    }

    // ---- publish: one pool reservation per warp, then every lane copies its records ----
    const int mine = active && !overflow ? n_kept : 0;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 0 && total) base = atomicAdd(p.pool_cursor, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    const bool exhausted = base + total > p.pool_cap;
    if (active) {
        const int off = base + incl - mine;
        if (!overflow && !exhausted)
            for (int i = 0; i < n_kept; i++) {
                const int ent = s.kept[i][tid];
                const int4 rec = p.boxes[ent];
                p.pool[off + i] = make_int4(rec.x, rec.y, rec.z, ent);
            }
        p.table[((size_t)tile * kMaxGroups + k) * p.n_lights + l] =
            make_int2(off, overflow || exhausted ? -1 : n_kept);
    }
}

cudaError_t launch_walks(const WalkParams& p, cudaStream_t st) {
    int first, tile_rows;
    owned_tile_rows(p.d, first, tile_rows);
    if (tile_rows <= 0 || p.n_lights <= 0) return cudaSuccess;
    const long long hw_pad = (p.d.HW + 31) & ~31;
    const long long threads = (long long)tile_rows * kMaxGroups * p.n_lights * hw_pad;
    k_walk<<<(unsigned)((threads + kWalkThreads - 1) / kWalkThreads), kWalkThreads, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace par
