// shade.cu — fused shading + shadow rays + RGBA8 quantise/pack: replaces the shading loop
// (/root/reference/src/alternative.cpp:702-760), trace_hash_for_light (399-500),
// AABB::intersect (40-83), Vector::normalize (sprites.hpp:28-35) and Color::operator*
// (sprites.hpp:8-16), generalised to N lights:
//     acc = sum over visible lights of max(0, n . t_l);   out = color * min(1, acc + ambient)
//
// What makes it fast without changing a bit of the output:
//  * The reference walks the grid once per PIXEL and per light; but the sequence of probed
//    bins depends only on (start bin, light bin), and every hit pixel of a 40x40 screen tile
//    starts in bin (tile x, tile y, z/40) (quirk Q11).  One CTA owns a tile, groups its pixels
//    by z/40 ("group"), compacts each group into a dense pixel list, and walks ONCE per
//    (group, light) — for a whole batch of lights at a time:
//      walk    one thread per run of steps: the fp32 position chain is accumulated sequentially
//              exactly as the reference does (quirk Q15); the 7 probes of a step collapse to the
//              distinct bins among them; their 4-bit counts are fetched in batches of independent
//              loads and occupied bins are appended to a shared list;
//      gather  a warp scan expands the occupied bins into dense (bin, slot) lanes: entity ->
//              de-duplicate per light (shared-memory hash set; testing a box twice cannot change
//              an OR) -> box -> shaft cull (shaft.cuh: boxes no ray of the group can hit) -> float
//              corners in the light's segment of a shared box list;
//      shade   one lane per pixel of the group: for every light of the batch, the Lambert term
//              and — only when it is > 0 (quirk Q19) — a loop of slab tests over that light's
//              boxes: unbounded-line semantics (Q14), the self-entity skip (Q17), the start-bin
//              skip (Q16) already applied in the walk, and std::min/std::max NaN semantics (Q13)
//              reproduced exactly: when the light lies strictly outside the group's origin bounds
//              no NaN can arise and the boxes are stored as (near, far) corners (6 FADD + 6 FMUL +
//              2 FMNMX3 per box); otherwise warps without a zero/NaN direction component use
//              FMNMX and the rest the literal ternaries.
//    Batches that do not fit the shared lists are split (fewer lights, then fewer steps of
//    one light); the shadow state of a split light is carried between rounds.
//  * Finished RGBA8 pixels are staged in shared memory and leave as 16-byte stores.
// All fp32 arithmetic is IEEE round-to-nearest with no FMA contraction (-fmad=false).
#include "par_kernels.cuh"
#include "shaft.cuh"

namespace par {

#ifndef PAR_SHADE_THREADS
#define PAR_SHADE_THREADS 160
#endif
#ifndef PAR_SHADE_LIST_CAP
#define PAR_SHADE_LIST_CAP 512
#endif
#ifndef PAR_SHADE_MIN_CTAS
#define PAR_SHADE_MIN_CTAS 6
#endif
constexpr int kThreads = PAR_SHADE_THREADS;   // CTA size of k_shade (a multiple of 32 dividing 1600)
constexpr int kListCap = PAR_SHADE_LIST_CAP;  // boxes in the shared list (lo/hi float4 pairs, 32 B each)
constexpr int kHashSize = 2 * kListCap;       // de-duplication set
constexpr int kHashBits = kListCap == 1024 ? 11 : kListCap == 512 ? 10 : 9;
constexpr int kSegMax = 16;                   // (light, step range) segments per round
constexpr int kMaxRun = 32;                   // most walk steps per phase-1 thread
constexpr int kScratchRows = kListCap * 8 / kThreads;  // ints per thread in the idle box list
constexpr int kTilePixels = kBin * kBin;
constexpr int kPixPerThread = kTilePixels / kThreads;
constexpr int kNoGroup = 0x7fffffff;
constexpr unsigned kEmpty = 0xffffffffu;
constexpr int kWarps = kThreads / 32;
static_assert(kTilePixels % kThreads == 0 && kThreads % 32 == 0 && kWarps <= 16, "CTA size");
static_assert((1 << kHashBits) == kHashSize, "hash size");

struct Segment {
    float sx, sy, sz;  // bin_step_size (alternative.cpp:423-425)
    int light;         // index into ShadeParams::lights
    int steps;         // (int)largest_bin_distance of the whole walk
    int ka, kb;        // step range [ka, kb) covered by this segment
    int item0;         // first phase-1 work item
    int count;         // boxes found (before de-duplication)
    int base;          // first slot of the segment in the box list
    int fill;          // boxes stored (after de-duplication)
    int octant;        // >= 0: every ray of the group has this sign octant (bit a = component a negative) and
                       // the boxes are stored as (near, far) corners; -1: mixed, boxes stored as (lo, hi)
};

struct ShadeSmem {
    float4 list[2 * kListCap];
    unsigned hash[kHashSize];
    unsigned occ[kListCap];  // occupied bins found by the walk: segment << 28 | count << 25 | flat bin (< 2^25)
    unsigned out[kTilePixels];  // finished RGBA8; between rounds of a split group: the fp32 acc bits
    unsigned short pix[kTilePixels];
    unsigned char sh[kTilePixels];
    Segment seg[kSegMax];
    int warp_scan[kWarps];
    int scan_total;
    int group;
    int n_occ;
    int overflow;  // a segment kept more boxes than its share of the list: redo the round with fewer segments
    int org_min[3], org_max[3];  // actual bounds of the current group's ray origins (shaft cull)
    int n_items;
    int run;  // walk steps per phase-1 thread this round
};

// alternative.cpp:40-83, literal std::min/std::max (valid for every input, NaN included).
__device__ __forceinline__ bool slab_hit_exact(const float4 lo, const float4 hi, float ox, float oy,
                                               float oz, float ix, float iy, float iz) {
    // (float)(int - int) of 16-bit operands equals the float difference exactly, so the
    // reference's int subtract + convert is one FADD here.
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

// Same test when no operand can be NaN (finite non-zero direction => finite products): then
// std::min/std::max and fminf/fmaxf agree up to the sign of zero, which no comparison sees.
__device__ __forceinline__ bool slab_hit_fast(const float4 lo, const float4 hi, float ox, float oy,
                                              float oz, float ix, float iy, float iz) {
    float x1 = (lo.x - ox) * ix, x2 = (hi.x - ox) * ix;
    float y1 = (lo.y - oy) * iy, y2 = (hi.y - oy) * iy;
    float z1 = (lo.z - oz) * iz, z2 = (hi.z - oz) * iz;
    float tmin = fmaxf(fmaxf(fminf(x1, x2), fminf(y1, y2)), fminf(z1, z2));
    float tmax = fminf(fminf(fmaxf(x1, x2), fmaxf(y1, y2)), fmaxf(z1, z2));
    return tmax >= tmin;
}

// Same test again when, in addition, the signs of the three inverse-direction components are the
// same for every pixel of the group (the light lies strictly outside the group's origin bounds
// on every axis).  With lo <= hi, subtraction and a multiplication by a constant are monotonic in
// fp32, so min(x1,x2) IS the product of the near corner (lo for a positive component, hi for a
// negative one) and max(x1,x2) that of the far corner: the six inner min/max disappear.  The
// gather stores the box as (near, far) corners for such a segment, so one code path serves all
// eight sign octants: 6 FADD + 6 FMUL + 2 FMNMX3 per box.
__device__ __forceinline__ bool slab_hit_near_far(const float4 nr, const float4 fr, float ox, float oy,
                                                  float oz, float ix, float iy, float iz) {
    const float tmin = fmaxf(fmaxf((nr.x - ox) * ix, (nr.y - oy) * iy), (nr.z - oz) * iz);
    const float tmax = fminf(fminf((fr.x - ox) * ix, (fr.y - oy) * iy), (fr.z - oz) * iz);
    return tmax >= tmin;
}

// kMode 0: (near, far) storage, sign octant uniform; 1: NaN-free, (lo, hi) storage; 2: exact.
template <int kMode>
__device__ __forceinline__ bool any_box_hit(const float4* __restrict__ boxes, int n, int self,
                                            float ox, float oy, float oz, float ix, float iy,
                                            float iz) {
#pragma unroll 1
    for (int e = 0; e < n; e++) {
        const float4 lo = boxes[2 * e], hi = boxes[2 * e + 1];
        bool hit;
        if (kMode == 2) hit = slab_hit_exact(lo, hi, ox, oy, oz, ix, iy, iz);
        else if (kMode == 1) hit = slab_hit_fast(lo, hi, ox, oy, oz, ix, iy, iz);
        else hit = slab_hit_near_far(lo, hi, ox, oy, oz, ix, iy, iz);
        if (hit && __float_as_int(lo.w) != self) return true;  // quirk Q17: own entity never shadows
    }
    return false;
}

// Block-wide exclusive scan of one int per thread (10 warps); *total gets the sum.
__device__ __forceinline__ int block_exclusive_scan(int v, ShadeSmem& s) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s.warp_scan[w] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        int x = threadIdx.x < kWarps ? s.warp_scan[threadIdx.x] : 0;
        int xi = x;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, xi, o);
            if (lane >= o) xi += t;
        }
        if (threadIdx.x < kWarps) s.warp_scan[threadIdx.x] = xi - x;
        if (threadIdx.x == kWarps - 1) s.scan_total = xi;
    }
    __syncthreads();
    return s.warp_scan[w] + incl - v;
}

// Box -> shared list slot `at`: (lo, hi) corners, or (near, far) corners for a uniform sign octant.
__device__ __forceinline__ void store_box(ShadeSmem& s, int at, const Box& b, int ent, int octant) {
    const float lx = (float)b.px, ly = (float)b.py, lz = (float)b.pz;
    const float hx = (float)(b.px + b.ex), hy = (float)(b.py + b.ey), hz = (float)(b.pz + b.ez);
    const bool sx = octant > 0 && (octant & 1), sy = octant > 0 && (octant & 2), sz = octant > 0 && (octant & 4);
    s.list[2 * at] = make_float4(sx ? hx : lx, sy ? hy : ly, sz ? hz : lz, __int_as_float(ent));
    s.list[2 * at + 1] = make_float4(sx ? lx : hx, sy ? ly : hy, sz ? lz : hz, 0.f);
}

// Sign octant shared by every ray of the current group towards light lt, or -1.  Uses the measured
// integer bounds of the group's origins; strict separation also guarantees that no direction
// component is zero, i.e. that the NaN cases of quirk Q13 cannot occur for this (group, light).
__device__ __forceinline__ int group_octant(const ShadeSmem& s, short4 lt) {
    const int L[3] = {lt.x, lt.y, lt.z};
    int oct = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if (s.org_min[a] < -32768 || s.org_max[a] > 32767) return -1;  // origins are cast to short
        if (L[a] < s.org_min[a]) oct |= 1 << a;
        else if (!(L[a] > s.org_max[a])) return -1;
    }
    return oct;
}

__device__ __forceinline__ unsigned quantise(uchar4 c, float f) {  // sprites.hpp:8-16
    const unsigned r = (unsigned char)((float)c.x * f);
    const unsigned g = (unsigned char)((float)c.y * f);
    const unsigned b = (unsigned char)((float)c.z * f);
    return r | g << 8 | b << 16 | (unsigned)c.w << 24;
}

__global__ void __launch_bounds__(kThreads, PAR_SHADE_MIN_CTAS)
k_shade(const __grid_constant__ ShadeParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ShadeSmem& s = *reinterpret_cast<ShadeSmem*>(smem_raw);

    const ViewDims& d = p.d;
    const int tid = threadIdx.x, lane = tid & 31;
    const int bx = blockIdx.x % d.HW;
    const int ty = p.tile_row_first + (blockIdx.x / d.HW) * max(d.stripe_n, 1);
    const int ra = max(ty * kBin, d.row0), rb = min(ty * kBin + kBin, d.row1);
    const int n_lights = p.n_lights;

    // Optional barrier-to-barrier phase timing (debug; p.phase_cycles is NULL in production).
    enum { kPhLoad, kPhFind, kPhCompact, kPhSetup, kPhWalk, kPhCounts, kPhDecide, kPhGather, kPhShade, kPhTail };
    static_assert(kScratchRows >= 14, "scratch too small for two walk steps");
    long long t_mark = 0;
    if (p.phase_cycles && tid == 0) t_mark = clock64();
    auto mark = [&](int phase) {
        if (p.phase_cycles && tid == 0) {
            const long long now = clock64();
            atomicAdd(&p.phase_cycles[phase], (unsigned long long)(now - t_mark));
            t_mark = now;
        }
    };

    // ---- load the tile: world z of every hit pixel (its start-bin z, the "group", is z / 40: ray_bin_z,
    //      alternative.cpp:727, C division truncating toward zero); miss pixels are final ----
    constexpr int kNoZ = -0x7fffffff - 1;
    auto group_of = [](int z) { return z == kNoZ ? kNoGroup : z / kBin; };
    int zv[kPixPerThread];
    // (z, texel | sprite << 10) of the thread's pixels, staged through the (still idle) box list
    // with cp.async: all ten copies of a thread are in flight at once.  As plain loads next to
    // the per-pixel branches below, the compiler keeps at most three in flight (registers) and
    // every CTA pays several DRAM round trips in a row before it can start.  Each thread reads
    // back only what it copied itself, so no barrier is needed — only cp.async.wait_all.
    int2* const stage = reinterpret_cast<int2*>(s.list);
    static_assert(sizeof(int2) * kTilePixels <= sizeof(float4) * 2 * kListCap, "tile staging fits the box list");
#pragma unroll
    for (int m = 0; m < kPixPerThread; m++) {
        const int pidx = m * kThreads + tid;  // pixel (row pidx / 40, column pidx % 40) of the tile
        const int j = ty * kBin + pidx / kBin;
        if (j >= ra && j < rb) {
            const int2* src = reinterpret_cast<const int2*>(&p.gbuf[(size_t)j * d.W + bx * kBin + pidx % kBin]) + 1;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&stage[pidx]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
    for (int m = 0; m < kPixPerThread; m++) {
        const int pidx = m * kThreads + tid;
        const int j = ty * kBin + pidx / kBin;
        zv[m] = kNoZ;
        if (j >= ra && j < rb) {
            const int2 zw = stage[pidx];
            const int gz = zw.x, gw = zw.y;
            if (gw >= 0 && n_lights > 0) {
                zv[m] = gz;
            } else {
                uchar4 c = make_uchar4(127, 127, 127, 0);  // miss colour, alternative.cpp:281
                if (gw >= 0) c = p.palette[p.atlas_color[(gw >> 10) * kTexels + (gw & 1023)]];
                s.out[pidx] = quantise(c, std_min(1.f, 0.f + p.ambient));
            }
        }
    }

    // Precomputed walks (walks.cu) for this tile?  Groups are enumerated in ascending z both here
    // and in the primary kernel, so the k-th group found below is the k-th GroupMeta.
    const int tile = ty * d.HW + bx;
    const bool tile_pre = p.table && !(p.debug_flags & 4) && p.tile_ngroups[tile] >= 0;
    int group_index = -1;

    int last_group = -0x7fffffff - 1;
    for (;;) {
        // ---- next group: the smallest start-bin z not yet processed in this tile ----
        if (tid == 0) {
            s.group = kNoGroup;
            s.org_min[0] = s.org_min[1] = s.org_min[2] = 0x7fffffff;
            s.org_max[0] = s.org_max[1] = s.org_max[2] = -0x7fffffff - 1;
        }
        __syncthreads();
        mark(last_group == -0x7fffffff - 1 ? kPhLoad : kPhShade);
        int mine = kNoGroup;
#pragma unroll
        for (int m = 0; m < kPixPerThread; m++)
            if (group_of(zv[m]) > last_group) mine = min(mine, group_of(zv[m]));
#pragma unroll
        for (int o = 16; o; o >>= 1) mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, o));
        if (lane == 0 && mine != kNoGroup) atomicMin(&s.group, mine);
        __syncthreads();
        mark(kPhFind);
        const int group = s.group;
        if (group == kNoGroup) break;
        last_group = group;
        group_index++;

        // ---- compact the group's pixels into a dense list ----
        int my_n = 0;
#pragma unroll
        for (int m = 0; m < kPixPerThread; m++) my_n += (group_of(zv[m]) == group);
        int pos = block_exclusive_scan(my_n, s);
        const int npix = s.scan_total;
        int lo3[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi3[3] = {-0x7fffffff - 1, -0x7fffffff - 1, -0x7fffffff - 1};
#pragma unroll
        for (int m = 0; m < kPixPerThread; m++)
            if (group_of(zv[m]) == group) {
                const int pidx = m * kThreads + tid;
                s.pix[pos++] = (unsigned short)pidx;
                // ray origin of this pixel (alternative.cpp:720-722), for the group's bounds:
                // x = column, z from the G-buffer, y = world_j - z (quirk Q11: y + z == H - row)
                const int j = ty * kBin + pidx / kBin, i = bx * kBin + pidx % kBin;
                const int o3[3] = {i, (short)(d.H - j) - zv[m], zv[m]};
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    lo3[a] = min(lo3[a], o3[a]);
                    hi3[a] = max(hi3[a], o3[a]);
                }
            }
#pragma unroll
        for (int a = 0; a < 3; a++) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                lo3[a] = min(lo3[a], __shfl_xor_sync(0xffffffffu, lo3[a], o));
                hi3[a] = max(hi3[a], __shfl_xor_sync(0xffffffffu, hi3[a], o));
            }
            if (lane == 0 && lo3[a] <= hi3[a]) {
                atomicMin(&s.org_min[a], lo3[a]);
                atomicMax(&s.org_max[a], hi3[a]);
            }
        }

        // start bin of every pixel of the group (alternative.cpp:724-727, quirk Q11)
        const int start = flat_bin(d, bx, ty, group);

        // ---- rounds over (light, step range) segments ----
        int l_cur = 0, ka_cur = 0;  // next unprocessed step of light l_cur
        int kb_try = -1;            // trial end of segment 0 (-1 = the whole walk)
        int nseg_try = kSegMax;
        bool fresh = true;          // first round of the group: acc starts at 0
        while (l_cur < n_lights) {
            __syncthreads();  // previous round fully consumed (lists, segments, pix list complete)
            mark(fresh ? kPhCompact : kPhShade);
            int n_fit = 0;
            bool fetched = false;
            if (tile_pre && ka_cur == 0 && kb_try < 0) {
                // ---- fast path: the walks of this group were done by walks.cu; fetch their box lists ----
                const int nseg = min(kSegMax, n_lights - l_cur);
                if (tid < nseg) {
                    const int2 e = p.table[((size_t)tile * kMaxGroups + group_index) * n_lights + l_cur + tid];
                    Segment& g = s.seg[tid];
                    g.light = l_cur + tid;
                    g.ka = 0;
                    g.kb = g.steps = 1;  // a whole walk
                    g.item0 = e.x;       // pool offset
                    g.count = g.fill = e.y;
                    g.octant = group_octant(s, p.lights[l_cur + tid]);
                }
                __syncthreads();
                bool all = true;
                int total = 0;
                for (int q = 0; q < nseg; q++) {
                    const int c = s.seg[q].count;
                    all = all && c >= 0;
                    if (all && n_fit == q && total + c <= kListCap) {
                        total += c;
                        n_fit = q + 1;
                    }
                }
                if (all) {  // (a list is at most kWalkListCap <= kListCap boxes, so n_fit >= 1)
                    fetched = true;
                    if (tid < n_fit) {
                        int base = 0;
                        for (int q = 0; q < tid; q++) base += s.seg[q].count;
                        s.seg[tid].base = base;
                    }
                    for (int e0 = tid; e0 < total; e0 += kThreads) {
                        int q = 0, off = e0;
                        while (off >= s.seg[q].count) off -= s.seg[q++].count;
                        const int4 rec = p.pool[s.seg[q].item0 + off];
                        store_box(s, e0, unpack_box(rec), rec.w, s.seg[q].octant);
                    }
                    __syncthreads();
                    mark(kPhGather);
                } else {
                    n_fit = 0;
                    __syncthreads();  // everyone has read the segment table before it is rewritten
                }
            }
            if (!fetched) {
            // A. describe the trial segments (walk set-up, alternative.cpp:406-430)
            if (tid < nseg_try && l_cur + tid < n_lights) {
                const short4 lt = p.lights[l_cur + tid];
                // light bin, alternative.cpp:729-732 ('/' truncates toward zero)
                const int lbx = lt.x / kBin, lby = (d.H - lt.y - lt.z) / kBin, lbz = lt.z / kBin;
                const float dx = (float)lbx - (float)bx, dy = (float)lby - (float)ty,
                            dz = (float)lbz - (float)group;
                const float big = fmaxf(fmaxf(fabsf(dx), fabsf(dy)), fabsf(dz));
                Segment& g = s.seg[tid];
                g.steps = (int)big;  // 0 when big < 1 (then the NaN step is never used)
                g.sx = dx / big;
                g.sy = dy / big;
                g.sz = dz / big;
                g.light = l_cur + tid;
                g.ka = tid == 0 ? ka_cur : 0;
                g.kb = (tid == 0 && kb_try >= 0) ? kb_try : g.steps;
                g.count = 0;
                g.fill = 0;
                g.octant = group_octant(s, lt);
            }
            for (int i = tid; i < kHashSize; i += kThreads) s.hash[i] = kEmpty;
            __syncthreads();
            mark(kPhSetup);
            const int nseg = min(nseg_try, n_lights - l_cur);
            if (tid == 0) {
                // run length: about one work item per thread, so the serial part of a walk stays short
                int steps = 0;
                for (int q = 0; q < nseg; q++) steps += s.seg[q].kb - s.seg[q].ka;
                const int run = min(kMaxRun, max(1, (steps + kThreads - 1) / kThreads));
                int items = 0;
                for (int q = 0; q < nseg; q++) {
                    s.seg[q].item0 = items;
                    items += (s.seg[q].kb - s.seg[q].ka + run - 1) / run;
                }
                s.n_items = items;
                s.run = run;
                s.n_occ = 0;
                s.overflow = 0;
            }
            __syncthreads();
            mark(kPhSetup);

            // C. phase 1: walk.  One thread per run of kRun steps of one segment.  The run is
            // processed in sub-chunks: first the distinct probed bins of up to 8 steps are
            // listed (ALU only) in a thread-private column of shared scratch, then their 4-bit
            // counts are fetched as one batch of independent loads.
            const int n_items = s.n_items, kRun = s.run;
            int* scratch = reinterpret_cast<int*>(s.list);  // [kScratchRows][kThreads]; the box list is idle now
            for (int it = tid; it < n_items; it += kThreads) {
                int q = 0;
                while (q + 1 < nseg && s.seg[q + 1].item0 <= it) q++;
                const Segment& g = s.seg[q];
                const int k0 = g.ka + (it - g.item0) * kRun, k1 = min(k0 + kRun, g.kb);
                const float sx = g.sx, sy = g.sy, sz = g.sz;
                // sequential fp32 accumulation from the start bin (quirk Q15)
                float px = (float)bx, py = (float)ty, pz = (float)group;
                int k = 0;
                for (; k + 8 <= k0; k += 8) {
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        px = px + sx;
                        py = py + sy;
                        pz = pz + sz;
                    }
                }
                for (; k < k0; k++) {
                    px = px + sx;
                    py = py + sy;
                    pz = pz + sz;
                }
                int x0 = (int)px, y0 = (int)py, z0 = (int)pz;
                const int sxy = d.HH * d.HL;
                while (k < k1) {
                    int n = 0;
                    for (int u = 0; u < 8 && k < k1 && n + 7 <= kScratchRows; u++, k++) {
                        px = px + sx;
                        py = py + sy;
                        pz = pz + sz;
                        const int x1 = (int)px, y1 = (int)py, z1 = (int)pz;
                        // The 7 probes of a step are the bins {x0|x1} x {y0|y1} x {z0|z1} minus
                        // "all old" (the all-old bin was the previous step's last probe, or the
                        // start bin, which is skipped anyway: quirk Q16).  Distinct bins among
                        // them = the non-empty subsets of the axes whose bin changed.
                        const int changed = (x1 != x0) | (y1 != y0) << 1 | (z1 != z0) << 2;
                        const int fx0 = x0 * sxy, fx1 = x1 * sxy, fy0 = y0 * d.HL, fy1 = y1 * d.HL;
                        for (int sub = changed; sub; sub = (sub - 1) & changed) {
                            const int f = ((sub & 1) ? fx1 : fx0) + ((sub & 2) ? fy1 : fy0) + ((sub & 4) ? z1 : z0);
                            if (f == start || (unsigned)f >= (unsigned)d.V) continue;  // Q16 / Q18
                            scratch[n * kThreads + tid] = f;
                            n++;
                        }
                        x0 = x1;
                        y0 = y1;
                        z0 = z1;
                    }
                    // the 4-bit counts of the sub-chunk, in batches of 8 independent loads; occupied bins
                    // go straight to the shared list (the warp-uniform atomicAdd is aggregated by ptxas)
                    int boxes = 0;
                    for (int i0 = 0; i0 < n; i0 += 8) {
                        int f[8], c[8];
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            f[u] = i0 + u < n ? scratch[(i0 + u) * kThreads + tid] : -1;
                            c[u] = f[u] >= 0 ? (__ldg(&p.occ4[f[u] >> 3]) >> ((f[u] & 7) * 4)) & 7 : 0;
                        }
#pragma unroll
                        for (int u = 0; u < 8; u++)
                            if (c[u]) {
                                const int o = atomicAdd(&s.n_occ, 1);
                                if (o < kListCap) s.occ[o] = (unsigned)q << 28 | (unsigned)c[u] << 25 | (unsigned)f[u];
                                boxes += c[u];
                            }
                    }
                    if (boxes) atomicAdd(&s.seg[q].count, boxes);
                }
            }
            __syncthreads();
            mark(kPhWalk);
            const int n_occ = s.n_occ;  // every occupied bin holds >= 1 box, so n_occ <= sum of counts
            // D. how many leading segments go into this round?  (every thread, redundantly)
            // The de-duplication set must stay sparse (candidates <= 3/4 of its slots) and the
            // occupied-bin list must be complete.  The box list itself is shared out evenly: after
            // de-duplication and the shaft cull a segment keeps a small fraction of its candidates,
            // so each of the n_fit segments gets room for min(candidates, kListCap / n_fit) boxes;
            // if one needs more, the round is redone with half the segments.
            n_fit = 0;
            {
                int total = 0;
                while (n_fit < nseg && total + s.seg[n_fit].count <= kHashSize * 3 / 4) {
                    total += s.seg[n_fit].count;
                    n_fit++;
                }
            }
            mark(kPhDecide);
            if (n_fit == 0 || n_occ > kListCap) {  // shrink: fewer lights first, then fewer steps of the first light
                if (nseg > 1) {
                    nseg_try = max(1, nseg / 2);
                } else {
                    const int ka = s.seg[0].ka, kb = s.seg[0].kb;
                    kb_try = ka + max(1, (kb - ka) / 2);
                }
                continue;
            }
            const int share = kListCap / n_fit;
            if (tid < n_fit) {  // base of segment tid in the box list
                int base = 0;
                for (int q = 0; q < tid; q++) base += min(s.seg[q].count, share);
                s.seg[tid].base = base;
            }

            // E. phase 2: one lane per candidate slot.  Each warp takes 32 occupied bins, scans their
            // counts and expands them into (bin, slot) pairs with shuffles, so that the dependent
            // loads (entity id -> box) run with dense lanes: entity -> de-duplicate -> box ->
            // shaft cull -> the segment's part of the box list.
            // Bounds of the group's ray origins for the shaft cull.  The origin is cast to short in
            // the reference (alternative.cpp:720-722): only cull when nothing can wrap.
            float org_lo[3], org_hi[3];
            bool can_cull = !(p.debug_flags & 1);
#pragma unroll
            for (int a = 0; a < 3; a++) {
                org_lo[a] = (float)s.org_min[a];
                org_hi[a] = (float)s.org_max[a];
                can_cull = can_cull && s.org_min[a] >= -32768 && s.org_max[a] <= 32767;
            }
            if (p.debug_flags & 2) {  // A/B switch: analytic (loose) bounds instead of the measured ones
                const int zl = group > 0 ? group * kBin : group * kBin - (kBin - 1);
                const int zh = group >= 0 ? group * kBin + (kBin - 1) : group * kBin;
                const int wl = d.H - (ty * kBin + kBin - 1), wh = d.H - ty * kBin;
                org_lo[0] = (float)(bx * kBin);
                org_hi[0] = (float)(bx * kBin + kBin - 1);
                org_lo[1] = (float)(wl - zh);
                org_hi[1] = (float)(wh - zl);
                org_lo[2] = (float)zl;
                org_hi[2] = (float)zh;
            }
            for (int ob = tid - lane; ob < n_occ; ob += kThreads) {
                const unsigned mine = ob + lane < n_occ ? s.occ[ob + lane] : 0u;
                const int my_c = (int)(mine >> 28) < n_fit ? (mine >> 25) & 7 : 0;
                int incl = my_c;
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
                    const int slot = t - (__shfl_sync(0xffffffffu, incl, src) - (int)((desc >> 25) & 7));
                    if (t >= total) continue;
                    const int q = desc >> 28;
                    const int ent = p.ids[(desc & 0x1ffffffu) * kSlots + slot];
                    const unsigned key = (unsigned)q << 26 | (unsigned)ent;
                    unsigned h = (key * 2654435761u) >> (32 - kHashBits);
                    bool fresh_key;
                    for (;;) {
                        const unsigned old = atomicCAS(&s.hash[h], kEmpty, key);
                        if (old == kEmpty || old == key) {
                            fresh_key = old == kEmpty;
                            break;
                        }
                        h = (h + 1) & (kHashSize - 1);
                    }
                    if (!fresh_key) continue;
                    const Box b = unpack_box(p.boxes[ent]);
                    if (can_cull) {
                        const short4 lt = p.lights[s.seg[q].light];
                        const float blo[3] = {(float)b.px, (float)b.py, (float)b.pz};
                        const float bhi[3] = {(float)(b.px + b.ex), (float)(b.py + b.ey), (float)(b.pz + b.ez)};
                        const float lp[3] = {(float)lt.x, (float)lt.y, (float)lt.z};
                        if (!shaft_may_hit(blo, bhi, lp, org_lo, org_hi)) continue;
                    }
                    int base = 0;
                    for (int r = 0; r < q; r++) base += min(s.seg[r].count, share);
                    const int nth = atomicAdd(&s.seg[q].fill, 1);
                    if (nth >= min(s.seg[q].count, share)) {
                        s.overflow = 1;
                        continue;
                    }
                    store_box(s, base + nth, b, ent, s.seg[q].octant);
                }
            }
            __syncthreads();
            mark(kPhGather);
            if (s.overflow) {  // some segment kept more than its share: fewer segments, then fewer steps
                if (n_fit > 1) {
                    nseg_try = max(1, n_fit / 2);
                } else {
                    const int ka = s.seg[0].ka, kb = s.seg[0].kb;
                    kb_try = ka + max(1, (kb - ka) / 2);
                    nseg_try = 1;
                }
                continue;
            }
            if (p.phase_cycles && tid == 0) {  // debug: candidate boxes found / kept after de-dup + cull
                unsigned long long found = 0, kept = 0;
                for (int q = 0; q < n_fit; q++) {
                    found += s.seg[q].count;
                    kept += s.seg[q].fill;
                }
                atomicAdd(&p.phase_cycles[10], found);
                atomicAdd(&p.phase_cycles[11], kept);
                atomicAdd(&p.phase_cycles[12], (unsigned long long)npix * n_fit);
            }

            }  // !fetched

            // F. phase 3: one lane per pixel of the group
            const bool final_round = s.seg[n_fit - 1].light == n_lights - 1 &&
                                     s.seg[n_fit - 1].kb == s.seg[n_fit - 1].steps;
            for (int qb = tid - lane; qb < npix; qb += kThreads) {
                const int qi = qb + lane;
                const bool valid = qi < npix;
                const int pidx = valid ? s.pix[qi] : 0;
                const int j = ty * kBin + pidx / kBin, i = bx * kBin + pidx % kBin;
                int4 g = make_int4(0, 0, 0, 0);
                float nx = 0.f, ny = 0.f, nz = 0.f, acc = 0.f;
                bool shadowed = false;
                if (valid) {
                    g = __ldcs(&p.gbuf[(size_t)j * d.W + i]);
                    const float* nrm = p.atlas_normal + ((g.w >> 10) * kTexels + (g.w & 1023)) * 3;
                    nx = __ldg(nrm);
                    ny = __ldg(nrm + 1);
                    nz = __ldg(nrm + 2);
                    if (!fresh) acc = __uint_as_float(s.out[pidx]);
                    shadowed = s.sh[qi] != 0;  // only meaningful when segment 0 continues a light
                }
                // Ray origin, alternative.cpp:720-722
                const float ox = (float)(short)i, oy = (float)(short)g.y, oz = (float)(short)g.z;
                for (int q = 0; q < n_fit; q++) {
                    const Segment& sg = s.seg[q];
                    const short4 lt = p.lights[sg.light];
                    if (sg.ka == 0) shadowed = false;
                    // towards_light, L1-normalised (alternative.cpp:711-715, sprites.hpp:28-35)
                    float tx = (float)(lt.x - i), tyv = (float)(lt.y - g.y), tz = (float)(lt.z - g.z);
                    const float len = fabsf(tx) + fabsf(tyv) + fabsf(tz);
                    tx = tx / len;
                    tyv = tyv / len;
                    tz = tz / len;
                    // alternative.cpp:745-747; a term of 0 adds +0 whether visible or not (Q19)
                    const float lam = std_max(0.f, nx * tx + ny * tyv + nz * tz);
                    const bool lit_candidate = valid && lam > 0.f;
                    const int n = sg.fill;
                    const bool test = lit_candidate && !shadowed && n > 0;
                    // direction_inverse, alternative.cpp:717-719 (only needed when testing)
                    float ix = 0.f, iy = 0.f, iz = 0.f;
                    if (test) {
                        ix = 1.f / tx;
                        iy = 1.f / tyv;
                        iz = 1.f / tz;
                    }
                    // a NaN can only arise from a zero (or NaN) direction component (quirk Q13)
                    const bool nan_free = fabsf(tx) > 0.f && fabsf(tyv) > 0.f && fabsf(tz) > 0.f;
                    const float4* boxes = s.list + 2 * sg.base;
                    if (__any_sync(0xffffffffu, test)) {
                        // warp-uniform choice of the slab-test variant
                        bool hit = false;
                        if (sg.octant >= 0) {
                            if (test) hit = any_box_hit<0>(boxes, n, g.x, ox, oy, oz, ix, iy, iz);
                        } else if (!__any_sync(0xffffffffu, test && !nan_free)) {
                            if (test) hit = any_box_hit<1>(boxes, n, g.x, ox, oy, oz, ix, iy, iz);
                        } else {
                            if (test) hit = any_box_hit<2>(boxes, n, g.x, ox, oy, oz, ix, iy, iz);
                        }
                        if (hit) shadowed = true;
                    }
                    if (sg.kb == sg.steps && lit_candidate && !shadowed) acc = acc + lam;
                }
                if (valid) {
                    if (final_round) {
                        const uchar4 c = p.palette[p.atlas_color[(g.w >> 10) * kTexels + (g.w & 1023)]];
                        // alternative.cpp:735 / 757-758
                        s.out[pidx] = quantise(c, std_min(1.f, acc + p.ambient));
                    } else {
                        s.out[pidx] = __float_as_uint(acc);
                        s.sh[qi] = shadowed;
                    }
                }
            }

            // advance past the processed segments
            const Segment& last = s.seg[n_fit - 1];
            if (last.kb == last.steps) {
                l_cur = last.light + 1;
                ka_cur = 0;
            } else {  // a split light: n_fit == 1
                l_cur = last.light;
                ka_cur = last.kb;
            }
            kb_try = -1;
            nseg_try = kSegMax;
            fresh = false;
        }
    }

    // ---- 16-byte stores of the finished tile rows ----
    __syncthreads();
    mark(kPhTail);
    for (int v = tid; v < kTilePixels / 4; v += kThreads) {
        const int j = ty * kBin + v / (kBin / 4);
        if (j < ra || j >= rb) continue;
        const uint4 px = *reinterpret_cast<const uint4*>(&s.out[4 * v]);
        // raster row, or the row's slot in the stripe-major staging frame (rank-contiguous)
        const int jo = p.out_stripe_T ? ((ty % d.stripe_n) * p.out_stripe_T + ty / d.stripe_n) * kBin + (j - ty * kBin) : j;
        const size_t at = (size_t)jo * d.W + bx * kBin + 4 * (v % (kBin / 4));
        *reinterpret_cast<uint4*>(&p.out[at]) = px;
        // Fused exchange: the same chunk goes straight into every peer GPU's frame (posted writes over
        // NVLink), so no all-gather pass over the frame is needed afterwards — only a barrier.
        for (int r = 0; r < p.n_peer_out; r++) *reinterpret_cast<uint4*>(&p.peer_out[r][at]) = px;
    }
}

size_t shade_smem_bytes() { return sizeof(ShadeSmem); }

cudaError_t configure_shade() {
    return cudaFuncSetAttribute(k_shade, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(ShadeSmem));
}

cudaError_t launch_shade(const ShadeParams& p, cudaStream_t st) {
    const ViewDims& d = p.d;
    int first, tile_rows;
    owned_tile_rows(d, first, tile_rows);
    if (tile_rows <= 0) return cudaSuccess;
    k_shade<<<tile_rows * d.HW, kThreads, sizeof(ShadeSmem), st>>>(p);
    return cudaGetLastError();
}

}  // namespace par
