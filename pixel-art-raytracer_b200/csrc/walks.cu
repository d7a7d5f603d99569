// walks.cu — the grid-walk half of trace_hash_for_light (/root/reference/src/alternative.cpp:399-476)
// as its own throughput kernel: ONE WARP per (tile, z-group, light).
//
// The sequence of bins a shadow ray probes depends only on its start bin and the light's bin,
// and every hit pixel of a 40x40 screen tile starts in bin (tile x, tile y, z/40) (quirk Q11).
// So the walk is done once per (tile, group, light) — here, with thousands of independent warps
// in flight so that its dependent chains (sequential fp32 position accumulation, quirk Q15;
// occupancy -> entity -> box loads) are hidden by thread-level parallelism instead of stalling
// a whole shading CTA.  The result — the de-duplicated boxes that can shadow some pixel of the
// group (shaft cull, see shade.cu) — goes to a pool in HBM; the shade kernel fetches the lists.
// Anything that does not fit the fixed per-warp buffers is marked "not available" and the shade
// kernel walks that case itself, so no scene is ever rejected.
#include "par_kernels.cuh"
#include "shaft.cuh"

namespace par {

constexpr int kWalkWarps = 8;      // warps per CTA
constexpr int kWalkOccCap = 128;   // occupied bins one walk may find
constexpr int kWalkHash = 512;     // de-duplication set of one walk
constexpr int kWalkMaxSlots = 384;  // candidate slots (boxes before de-duplication) one walk may gather
constexpr unsigned kWalkEmpty = 0xffffffffu;

struct WalkWarpSmem {
    unsigned occ[kWalkOccCap];  // count << 25 | flat bin
    unsigned hash[kWalkHash];
    int4 kept[kWalkListCap];
    int n_occ;
    int n_kept;
    int pool_base;
    int pad;
};

__global__ void __launch_bounds__(kWalkWarps * 32)
k_walk(const __grid_constant__ WalkParams p) {
    __shared__ WalkWarpSmem smem[kWalkWarps];
    const ViewDims& d = p.d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WalkWarpSmem& s = smem[warp];

    const int l = blockIdx.x % p.n_lights;
    const int t_local = blockIdx.x / p.n_lights;
    const int bx = t_local % d.HW;
    const int ty = p.tile_row_first + (t_local / d.HW) * max(d.stripe_n, 1);
    const int tile = ty * d.HW + bx;
    const int n_groups = p.tile_ngroups[tile];
    const short4 lt = p.lights[l];
    // light bin, alternative.cpp:729-732 ('/' truncates toward zero)
    const int lbx = lt.x / kBin, lby = (d.H - lt.y - lt.z) / kBin, lbz = lt.z / kBin;
    const float lp[3] = {(float)lt.x, (float)lt.y, (float)lt.z};
    const int sxy = d.HH * d.HL;

    for (int k = warp; k < n_groups; k += kWalkWarps) {
        const GroupMeta gm = p.groups[(size_t)tile * kMaxGroups + k];
        const int group = gm.z;
        int2* slot = &p.table[((size_t)tile * kMaxGroups + k) * p.n_lights + l];
        // walk set-up, alternative.cpp:406-430
        const int start = flat_bin(d, bx, ty, group);
        const float dx = (float)lbx - (float)bx, dy = (float)lby - (float)ty, dz = (float)lbz - (float)group;
        const float big = fmaxf(fmaxf(fabsf(dx), fabsf(dy)), fabsf(dz));
        const int steps = (int)big;  // 0 when big < 1 (then the NaN step is never used)
        const float sx = dx / big, sy = dy / big, sz = dz / big;

        if (lane == 0) {
            s.n_occ = 0;
            s.n_kept = 0;
        }
        for (int i = lane; i < kWalkHash; i += 32) s.hash[i] = kWalkEmpty;
        __syncwarp();

        // ---- walk: lane handles steps [k0, k1) ----
        const int run = (steps + 31) / 32;
        const int k0 = min(lane * run, steps), k1 = min(k0 + run, steps);
        if (k0 < k1) {
            // sequential fp32 accumulation from the start bin (quirk Q15)
            float px = (float)bx, py = (float)ty, pz = (float)group;
            int kk = 0;
            for (; kk + 8 <= k0; kk += 8) {
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    px = px + sx;
                    py = py + sy;
                    pz = pz + sz;
                }
            }
            for (; kk < k0; kk++) {
                px = px + sx;
                py = py + sy;
                pz = pz + sz;
            }
            int x0 = (int)px, y0 = (int)py, z0 = (int)pz;
            for (; kk < k1; kk++) {
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
                    if (!c) continue;
                    const int o = atomicAdd(&s.n_occ, 1);
                    if (o < kWalkOccCap) s.occ[o] = c << 25 | (unsigned)f;
                }
                x0 = x1;
                y0 = y1;
                z0 = z1;
            }
        }
        __syncwarp();
        const int n_occ = s.n_occ;
        bool overflow = n_occ > kWalkOccCap;
        {   // the de-duplication set must stay sparse: bound the candidate slots of the walk
            int slots = 0;
            for (int i = lane; i < min(n_occ, kWalkOccCap); i += 32) slots += (s.occ[i] >> 25) & 7;
#pragma unroll
            for (int o = 16; o; o >>= 1) slots += __shfl_xor_sync(0xffffffffu, slots, o);
            overflow = overflow || slots > kWalkMaxSlots;
        }

        // ---- gather: expand bins into (bin, slot) lanes; entity -> de-dup -> box -> shaft cull ----
        float org_lo[3], org_hi[3];
        bool can_cull = !(p.debug_flags & 1);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            org_lo[a] = (float)gm.omin[a];
            org_hi[a] = (float)gm.omax[a];
            can_cull = can_cull && gm.omin[a] >= -32768 && gm.omax[a] <= 32767;  // origins are cast to short
        }
        for (int ob = 0; ob < n_occ && !overflow; ob += 32) {
            const unsigned mine = ob + lane < n_occ ? s.occ[ob + lane] : 0u;
            int incl = (mine >> 25) & 7;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            for (int t0 = 0; t0 < total; t0 += 32) {
                const int t = t0 + lane;
                int src = 0;  // first lane whose inclusive prefix exceeds t
#pragma unroll
                for (int step = 16; step; step >>= 1) {
                    const int v = __shfl_sync(0xffffffffu, incl, src + step - 1);
                    if (v <= t) src += step;
                }
                src = min(src, 31);
                const unsigned desc = __shfl_sync(0xffffffffu, mine, src);
                const int slot_i = t - (__shfl_sync(0xffffffffu, incl, src) - (int)((desc >> 25) & 7));
                if (t >= total) continue;
                const int ent = p.ids[(desc & 0x1ffffffu) * kSlots + slot_i];
                unsigned h = ((unsigned)ent * 2654435761u) >> 23;
                bool fresh_key;
                for (;;) {
                    const unsigned old = atomicCAS(&s.hash[h], kWalkEmpty, (unsigned)ent);
                    if (old == kWalkEmpty || old == (unsigned)ent) {
                        fresh_key = old == kWalkEmpty;
                        break;
                    }
                    h = (h + 1) & (kWalkHash - 1);
                }
                if (!fresh_key) continue;
                const int4 rec = p.boxes[ent];
                const Box b = unpack_box(rec);
                if (can_cull) {
                    const float blo[3] = {(float)b.px, (float)b.py, (float)b.pz};
                    const float bhi[3] = {(float)(b.px + b.ex), (float)(b.py + b.ey), (float)(b.pz + b.ez)};
                    if (!shaft_may_hit(blo, bhi, lp, org_lo, org_hi)) continue;
                }
                const int at = atomicAdd(&s.n_kept, 1);
                if (at < kWalkListCap) s.kept[at] = make_int4(rec.x, rec.y, rec.z, ent);
            }
            __syncwarp();
            overflow = overflow || s.n_kept > kWalkListCap;  // no point in going on
        }
        __syncwarp();
        const int n_kept = s.n_kept;
        overflow = overflow || n_kept > kWalkListCap;

        // ---- publish: reserve pool space, copy, write the table entry ----
        if (lane == 0) s.pool_base = overflow || n_kept == 0 ? 0 : atomicAdd(p.pool_cursor, n_kept);
        __syncwarp();
        const int base = s.pool_base;
        if (!overflow && n_kept > 0 && base + n_kept > p.pool_cap) overflow = true;  // pool exhausted
        if (!overflow)
            for (int i = lane; i < n_kept; i += 32) p.pool[base + i] = s.kept[i];
        if (lane == 0) *slot = make_int2(base, overflow ? -1 : n_kept);
        __syncwarp();
    }
}

cudaError_t launch_walks(const WalkParams& p, cudaStream_t st) {
    int first, tile_rows;
    owned_tile_rows(p.d, first, tile_rows);
    if (tile_rows <= 0 || p.n_lights <= 0) return cudaSuccess;
    k_walk<<<tile_rows * p.d.HW * p.n_lights, kWalkWarps * 32, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace par
