// tile.cu — the render kernel: ONE CTA per 40x40 screen tile does the whole per-pixel path of
// the reference frame loop for that tile,
//     trace_hash_for_pixel           /root/reference/src/alternative.cpp:271-383   (phase P)
//     shading loop                   alternative.cpp:702-760                       (phases G, R)
//     trace_hash_for_light           alternative.cpp:399-500
//     AABB::intersect                alternative.cpp:40-83
//     Vector::normalize, Color::operator*   sprites.hpp:28-35, 8-16
// generalised to N lights:  acc = sum over visible lights of max(0, n . t_l);
// out = color * min(1, acc + ambient).
//
// Why one kernel: a 40x40 tile is one column of bins for the primary rays AND the unit that shares
// shadow-ray grid walks, so the G-buffer record of a pixel (entity, z, texel) never has to leave the
// SM: it goes from the registers of phase P into 10 bytes of shared memory and is consumed there.
// The 16-byte-per-pixel G-buffer in HBM is written only when a caller asks for the parity
// checkpoint (par_render(out_gbuf), par_get_gbuffer).
//
//   P  primary rays (integer only).  The column's occupied bins are scanned and their entries
//      gathered into shared memory in the reference's (bin_z, slot) order, pre-digested into the
//      integers the per-pixel test needs (chunks of 256 entries).  A thread owns 10 pixels of ONE
//      screen column (rows 4 apart) and walks the list for 5 of them at a time: 2-D integer hit test
//      (quirk Q6), texel index with the sprite's own width (Q7 lifted), strict-greater depth select
//      (Q8), two-adjacent-bins early-out with empty-bin reset (Q9), miss pixel (Q10), record (Q11).
//   G  grouping.  Every hit pixel starts its shadow walks in bin (tile x, tile y, z / 40) (Q11) and
//      the probed-bin sequence of a walk depends only on (start bin, light bin): pixels are
//      counting-sorted by z / 40 into dense per-group lists (all groups of the tile at once; warp
//      ballots + redux aggregate the shared-memory atomics), with the integer bounds of each
//      group's ray origins for the shaft cull.
//   R  rounds over (group, light) segments, up to 32 per round:
//        walk    one thread per run of steps: the fp32 position chain is accumulated sequentially
//                exactly as the reference does (Q15); the 7 probes of a step collapse to the
//                distinct bins among them; 4-bit counts fetched as batches of independent loads;
//        gather  a warp scan expands the occupied bins into dense (bin, slot) lanes: entity -> box ->
//                shaft cull (shaft.cuh) -> de-duplicate the survivors per segment (shared hash set;
//                testing a box twice cannot change an OR) -> float corners in the segment's share of
//                a shared box list;
//        shade   one lane per pixel of the round's groups: per light the L1-normalised direction
//                (Q12), the Lambert term and — only when it is > 0 (Q19) — the slab tests of that
//                segment's boxes with unbounded-line semantics (Q14), self-entity skip (Q17) and
//                std::min/std::max NaN semantics (Q13) reproduced exactly (three variants, see
//                slab_hit_*).  A round that does not fit the shared lists is walked again with what
//                the failed walk measured (the leading segments that fit, or the step range of one
//                light that its density allows); the (acc, shadowed) state of an unfinished pixel is
//                parked in its own slot of the output frame.
//   Finished RGBA8 pixels are staged in shared memory and leave as 16-byte stores — into the own
//   frame and, fused frame exchange, in place into the frames of peer GPUs over NVLink.
// All fp32 arithmetic is IEEE round-to-nearest with no FMA contraction (-fmad=false).
#include <climits>

#include "par/par.h"
#include "par_kernels.cuh"
#include "shaft.cuh"

// This file is compiled twice.  On its own it gives the general configuration: 5 CTAs per SM (72 registers,
// 44.6 KB of shared memory) with lists sized for many-light rounds.  Through tile_one_light.cu
// (PAR_TILE_ONE_LIGHT) it gives the configuration for one-light frames of many CTA waves: 6 CTAs per SM (64
// registers, 36.9 KB) with the smallest lists the phases allow — a one-light round holds a tile's few z-groups
// only, and a sixth resident CTA covers more of the barrier and latency chain of the others' rounds (3840x2160
// default scene: 0.248 -> 0.230 ms; 16 lights: 2.14 -> 2.56 ms, which is why it is not the only one).  Same
// code, same results: rounds that overflow a list are re-walked in either (par_api.cu picks per frame).
#ifdef PAR_TILE_ONE_LIGHT
#define PAR_TILE_MIN_CTAS 6
#define PAR_TILE_LIST_CAP 288
#define PAR_TILE_HASH_BITS 9
#define PAR_TILE_OCC_CAP 384
#define PAR_TILE_ENTRY_CAP 176
#define k_tile k_tile_one_light
#endif

namespace par {

#ifndef PAR_TILE_MIN_CTAS
#define PAR_TILE_MIN_CTAS 5
#endif
#ifndef PAR_TILE_LIST_CAP
#define PAR_TILE_LIST_CAP 384
#endif
namespace {

constexpr int kT = kTileCtaThreads;             // 160 threads: 40 columns x 4 row phases
constexpr int kPix = kBin * kBin;               // 1600 pixels per tile
constexpr int kPPT = kPix / kT;                 // 10 pixels per thread (rows rsub + 4m)
constexpr int kHalf = kPPT / 2;                 // pixels per primary pass
constexpr int kListCap = PAR_TILE_LIST_CAP;     // boxes in the shared list (two float4 each)
#ifndef PAR_TILE_HASH_BITS
#define PAR_TILE_HASH_BITS 10
#endif
constexpr int kHashBits = PAR_TILE_HASH_BITS;
constexpr int kHashSize = 1 << kHashBits;       // de-duplication set
#ifndef PAR_TILE_OCC_CAP
#define PAR_TILE_OCC_CAP 896
#endif
constexpr int kOccCap = PAR_TILE_OCC_CAP;       // occupied bins found by the walks of one round
constexpr int kSegMax = 32;                     // (group, light, step range) segments per round
constexpr int kSegStart = 10;                   // ... of a many-light tile's first round (then adaptive)
constexpr int kGroupMax = 64;                   // z-groups handled per pass over the tile
#ifndef PAR_TILE_ENTRY_CAP
#define PAR_TILE_ENTRY_CAP 256
#endif
constexpr int kEntryCap = PAR_TILE_ENTRY_CAP;   // column entries staged at a time (phase P)
constexpr int kMaxRun = 32;                     // most walk steps per walk-phase thread
constexpr int kMaxHL = PAR_MAX_VIEW / kBin;     // 320 bins along z at most
constexpr int kNoGroup = INT_MAX;
constexpr int kDoneZ = -32768;                  // TileSmem::z of a pixel whose RGBA8 value is final
constexpr unsigned kEmpty = 0xffffffffu;
static_assert(kPix % kT == 0 && kT % 32 == 0 && kT == 4 * kBin && kPPT % 2 == 0 && kSegMax == 32, "CTA shape");
static_assert(kHashSize * 3 / 4 >= kListCap, "the hash set must hold a full list");

struct Seg {
    float sx, sy, sz;  // bin_step_size (alternative.cpp:423-425)
    int light;         // index into TileParams::lights
    int grp;           // index into TileSmem::grp
    int start;         // flat index of the start bin (quirk Q16)
    int steps;         // (int)largest_bin_distance of the whole walk
    int ka, kb;        // step range [ka, kb) covered by this segment
    int item0;         // first walk-phase work item
    int count;         // boxes found (before de-duplication)
    int occ;           // occupied bins found
    int base;          // first slot of the segment in the box list
    int fill;          // boxes stored (after de-duplication and cull)
    int octant;        // >= 0: every ray of the group has this sign octant (bit a = component a negative) and
                       // the boxes are stored as (near, far) corners; -1: mixed, boxes stored as (lo, hi)
    float rl[3], rh[3];  // shaft cull, the part that does not depend on the box (shaft_prepare)
    int cull;            // bits 0-2: axes without a constraint; 8: no cull for this segment
};

struct Grp {
    int gz;                  // start-bin z of the group = world z / 40 (ray_bin_z, alternative.cpp:727)
    int n;                   // pixels
    int p0;                  // first slot in TileSmem::pix
    int omin[3], omax[3];    // integer bounds of the group's ray origins (alternative.cpp:720-722)
};

struct RoundSmem {  // phases G/R
    float4 list[2 * kListCap];
    unsigned hash[kHashSize];
    unsigned occ_bin[kOccCap];       // flat bin
    unsigned char occ_meta[kOccCap]; // segment << 3 | count
};
struct PrimarySmem {  // phase P (aliases RoundSmem)
    int cnt[kMaxHL];
    int off[kMaxHL + 1];
    int4 A[kEntryCap];  // x0, x1 (exclusive), lo = py+pz, top = py+ey+pz+ez
    int4 B[kEntryCap];  // key0 = py-pz, ey, pz, texel base of the sprite
    int2 C[kEntryCap];  // entity, width << 2 | gap << 1 | first
};
constexpr int kStageTexels = (int)((sizeof(RoundSmem) - sizeof(PrimarySmem)) / sizeof(int));
static_assert(sizeof(PrimarySmem) + 800 * sizeof(int) <= sizeof(RoundSmem), "the 20x40 sprite's depths must fit");

struct TileSmem {
    union {
        RoundSmem r;
        struct {
            PrimarySmem p;
            int depth[kStageTexels];
        } pp;
    };
    int ent[kPix];            // hit entity (quirk Q17 needs it)
    unsigned w[kPix];         // global texel index of the hit; becomes the finished RGBA8 pixel
    short z[kPix];            // world z of the hit (fits: see par_set_atlas_sized)
    unsigned short pix[kPix]; // pixel indices sorted by group
    Grp grp[kGroupMax];
    int cursor[kGroupMax];
    Seg seg[kSegMax];
    unsigned bits[2];
    int more;     // some pixel's z-group lies beyond the current window of kGroupMax groups
    int gmin;
    int n_occ;
    int overflow;
    int n_keys;   // entries of the de-duplication set
    int n_items;
    int run;
};

constexpr int kScratchRows = (int)(sizeof(float4) * 2 * kListCap / sizeof(int) / kT);
static_assert(kScratchRows >= 14, "scratch too small for two walk steps");

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

// Box -> shared list slot `at`: (lo, hi) corners, or (near, far) corners for a uniform sign octant.
__device__ __forceinline__ void store_box(float4* list, int at, const Box& b, int ent, int octant) {
    const float lx = (float)b.px, ly = (float)b.py, lz = (float)b.pz;
    const float hx = (float)(b.px + b.ex), hy = (float)(b.py + b.ey), hz = (float)(b.pz + b.ez);
    const bool sx = octant > 0 && (octant & 1), sy = octant > 0 && (octant & 2), sz = octant > 0 && (octant & 4);
    list[2 * at] = make_float4(sx ? hx : lx, sy ? hy : ly, sz ? hz : lz, __int_as_float(ent));
    list[2 * at + 1] = make_float4(sx ? lx : hx, sy ? ly : hy, sz ? lz : hz, 0.f);
}

// Sign octant shared by every ray of a group towards light lt, or -1.  Uses the measured integer
// bounds of the group's origins; strict separation also guarantees that no direction component is
// zero, i.e. that the NaN cases of quirk Q13 cannot occur for this (group, light).
__device__ __forceinline__ int group_octant(const Grp& g, short4 lt) {
    const int L[3] = {lt.x, lt.y, lt.z};
    int oct = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        if (g.omin[a] < -32768 || g.omax[a] > 32767) return -1;  // origins are cast to short
        if (L[a] < g.omin[a]) oct |= 1 << a;
        else if (!(L[a] > g.omax[a])) return -1;
    }
    return oct;
}

__device__ __forceinline__ unsigned quantise(unsigned rgba, float f) {  // sprites.hpp:8-16
    const unsigned r = (unsigned char)((float)(rgba & 255u) * f);
    const unsigned g = (unsigned char)((float)((rgba >> 8) & 255u) * f);
    const unsigned b = (unsigned char)((float)((rgba >> 16) & 255u) * f);
    return r | g << 8 | b << 16 | (rgba & 0xff000000u);
}

constexpr unsigned kMissColor = 127u | 127u << 8 | 127u << 16;  // alternative.cpp:281, alpha 0

}  // namespace

// kChecks: the parity / debug outputs (G-buffer checkpoint, primary-only mode, fp32 intermediates, phase
// timing) are compiled into a second instantiation; production frames run the one without them.
template <bool kChecks>
__global__ void __launch_bounds__(kT, PAR_TILE_MIN_CTAS)
k_tile(const __grid_constant__ TileParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem& s = *reinterpret_cast<TileSmem*>(smem_raw);

    const ViewDims& d = p.d;
    const int tid = threadIdx.x, lane = tid & 31;
    const long long t_begin = p.tile_cost ? clock64() : 0;
    const int slot = p.tile_order ? p.tile_order[blockIdx.x] : (int)blockIdx.x;
    const int tps = tiles_per_stripe(d);
    const int stripe = p.tile_row_first + (slot / tps) * max(d.stripe_n, 1);  // (a stripe is a tile row unless stripe_s > 1)
    const int ty = stripe / stripe_segments(d);
    const int bx = stripe_column_segment(d, stripe) * tps + slot % tps;
    const int ra = max(ty * kBin, d.row0), rb = min(ty * kBin + kBin, d.row1);
    const int n_lights = p.n_lights;
    // The thread's pixels: rows rsub + 4m of column col.  A warp covers 8 adjacent columns (x 4 row
    // phases), so the x half of the primary hit test — a sprite is narrower than a tile — is uniform
    // in most warps instead of splitting every warp in two.
    const int col = tid >> 2, rsub = tid & 3;
    const int i = bx * kBin + col;
    const bool probe_here = p.probe_x / kBin == bx && p.probe_y / kBin == ty && p.probe_x >= 0;

    // Optional barrier-to-barrier phase timing (debug; p.phase_cycles is NULL in production).
    enum { kPhPrimary, kPhGroup, kPhSetup, kPhWalk, kPhGather, kPhShade, kPhTail };
    long long t_mark = 0;
    if (kChecks && p.phase_cycles && tid == 0) t_mark = clock64();
    auto mark = [&](int phase) {
        if (kChecks && p.phase_cycles && tid == 0) {
            const long long now = clock64();
            atomicAdd(&p.phase_cycles[phase], (unsigned long long)(now - t_mark));
            t_mark = now;
        }
    };

    // =============================== P. primary rays ===============================
    {
        PrimarySmem& ps = s.pp.p;
        const int col0 = flat_bin(d, bx, ty, 0);
        // the column's counts (the reference's wrapping count is cnt & 7) and, for small atlases, the depth tables
        for (int bz = tid; bz < d.HL; bz += kT) ps.cnt[bz] = __ldg(&p.cnt[col0 + bz]) & (kSlots - 1);
        const bool staged = p.atlas_texels <= kStageTexels;
        if (staged)
            for (int t = tid; t < p.atlas_texels; t += kT) s.pp.depth[t] = __ldg(&p.atlas_depth[t]);
        __syncthreads();
        if (tid < 32) {  // exclusive scan over bin_z
            int carry = 0;
            for (int base = 0; base < d.HL; base += 32) {
                const int bz = base + tid;
                const int v = bz < d.HL ? ps.cnt[bz] : 0;
                int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (tid >= o) incl += t;
                }
                if (bz < d.HL) ps.off[bz] = carry + incl - v;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (tid == 0) ps.off[d.HL] = carry;
        }
        __syncthreads();
        const int n_total = ps.off[d.HL];
        const bool single = n_total <= kEntryCap;  // the whole column fits one chunk: stage it once

#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            int best[kHalf], r_ent[kHalf], r_z[kHalf], r_w[kHalf], run[kHalf];
            unsigned any_bits = 0, live = 0;  // bit m: has_intersected in the current bin / pixel still marching
#pragma unroll
            for (int m = 0; m < kHalf; m++) {
                const int j = ty * kBin + rsub + 4 * (pass * kHalf + m);
                best[m] = INT_MIN;  // closest_entity_depth, alternative.cpp:289
                r_ent[m] = 0;       // miss pixel: entity 0, z 0 (quirk Q10)
                r_z[m] = 0;
                r_w[m] = -1;
                run[m] = 0;         // intersected_bin_count
                if (j >= ra && j < rb) live |= 1u << m;
            }
            const int wj0 = (short)(d.H - (ty * kBin + rsub + 4 * pass * kHalf));  // world_j of pixel m: wj0 - 4m (alternative.cpp:280)
            int b0 = 0;
#pragma unroll 1
            while (b0 < d.HL && n_total > 0) {
                int b1 = d.HL;
                if (!single) {  // largest b1 with off[b1] - off[b0] <= kEntryCap (a bin holds <= 7 entries)
                    b1 = b0 + 1;
                    while (b1 < d.HL && ps.off[b1 + 1] - ps.off[b0] <= kEntryCap) b1++;
                }
                const int e0 = ps.off[b0], n = ps.off[b1] - e0;
                if (!(single && pass == 1)) {
                    if (!(pass == 0 && b0 == 0)) __syncthreads();  // everyone is done with the previous chunk
                    // gather the chunk's entries in (bin_z ascending, slot ascending) order
                    for (int t = tid; t < (b1 - b0) * kSlots; t += kT) {
                        const int bz = b0 + (t >> 3), sl = t & 7, c = ps.cnt[bz];
                        if (sl >= c) continue;
                        const int ent = __ldg(&p.ids[(size_t)(col0 + bz) * kSlots + (c - 1 - sl)]);
                        const Box b = unpack_box(__ldg(&p.boxes[ent]));
                        const int2 dims = __ldg(&p.sprite_dims[b.sprite]);  // texel base, width | height << 16
                        const int pos = ps.off[bz] - e0 + sl;
                        ps.A[pos] = make_int4(b.px, b.px + b.ex, b.py + b.pz, b.py + b.ey + b.pz + b.ez);
                        ps.B[pos] = make_int4(b.py - b.pz, b.ey, b.pz, dims.x);
                        const int first = (sl == 0);
                        const int gap = first && (bz == 0 || ps.cnt[bz - 1] == 0);  // an empty bin precedes this one
                        ps.C[pos] = make_int2(ent, (dims.y & 0xffff) << 2 | gap << 1 | first);
                    }
                    __syncthreads();
                }
                // per-pixel walk: the entry list is traversed once per thread and pass; the x half of the
                // hit test (quirk Q6) is shared by the thread's pixels (one screen column)
                for (int k = 0; k < n && live; k++) {
                    const int2 c = ps.C[k];
                    if (c.y & 1) {  // first entry of a bin: close the previous bin (alternative.cpp:368-374)
#pragma unroll
                        for (int m = 0; m < kHalf; m++) {
                            run[m] += (any_bits >> m) & 1;
                            if (run[m] >= 2) live &= ~(1u << m);  // two adjacent hit bins end the march (quirk Q9)
                            if (c.y & 2) run[m] = 0;              // an empty bin in between resets the run (298-300)
                        }
                        any_bits = 0;
                        if (!live) break;
                    }
                    const int4 a = ps.A[k];
                    if (i < a.x || i >= a.y) continue;  // x half of quirk Q6
                    const int4 b = ps.B[k];
                    const int width = c.y >> 2;
#pragma unroll
                    for (int m = 0; m < kHalf; m++) {
                        const int wj = wj0 - 4 * m;
                        if (!((live >> m) & 1) || !(wj > a.z && wj <= a.w)) continue;  // y half of Q6
                        const int row = a.w - wj;
                        const int idx = b.w + row * width + (i - a.x);  // quirk Q7 with the sprite's own width
                        const int dep = staged ? s.pp.depth[idx] : __ldg(&p.atlas_depth[idx]);
                        const int key = b.x + min(0, b.y - row) - dep;  // quirk Q8
                        if (best[m] < key) {  // strict: ties keep the earlier (bin_z, slot)
                            best[m] = key;
                            r_ent[m] = c.x;
                            r_z[m] = b.z + dep;  // quirk Q11; y = world_j - z
                            r_w[m] = idx;
                            any_bits |= 1u << m;
                        }
                    }
                }
                b0 = b1;
            }
            // records of this pass -> shared memory (and the parity checkpoints, when asked for)
#pragma unroll
            for (int m = 0; m < kHalf; m++) {
                const int row = rsub + 4 * (pass * kHalf + m);
                const int pidx = row * kBin + col, j = ty * kBin + row;
                s.ent[pidx] = r_ent[m];
                s.z[pidx] = (short)r_z[m];
                s.w[pidx] = (unsigned)r_w[m];
                if (j < ra || j >= rb) {
                    s.w[pidx] = 0xfffffffeu;  // not rendered by this context
                    continue;
                }
                const int y = r_w[m] >= 0 ? (wj0 - 4 * m) - r_z[m] : 0;
                if (kChecks && p.gbuf) p.gbuf[(size_t)j * d.W + i] = make_int4(r_ent[m], y, r_z[m], r_w[m]);
                if (probe_here && i == p.probe_x && j == p.probe_y) {  // cursor probe (mouse_pixel, alternative.cpp:380-382)
                    float4 tx = make_float4(0.f, 0.f, 0.f, __uint_as_float(kMissColor));
                    if (r_w[m] >= 0) tx = __ldg(&p.texel_tab[r_w[m]]);
                    const int rec[7] = {__float_as_int(tx.x), __float_as_int(tx.y), __float_as_int(tx.z),
                                        __float_as_int(tx.w), y, r_z[m], r_ent[m]};
                    for (int k = 0; k < 7; k++) {
                        if (p.probe_a) p.probe_a[k] = rec[k];
                        if (p.probe_b) p.probe_b[k] = rec[k];
                    }
                }
            }
        }
    }
    __syncthreads();  // records complete; the staging area is free
    mark(kPhPrimary);

    // ---- miss pixels (and every pixel when there is no light) are final; group key of the others ----
    int gz[kPPT];
    const float amb_only = std_min(1.f, 0.f + p.ambient);
#pragma unroll
    for (int m = 0; m < kPPT; m++) {
        const int pidx = (rsub + 4 * m) * kBin + col;
        const unsigned w = s.w[pidx];
        gz[m] = kNoGroup;
        if (w == 0xfffffffeu) {
            s.z[pidx] = kDoneZ;
            continue;
        }
        if (w == 0xffffffffu) {
            s.w[pidx] = quantise(kMissColor, amb_only);
            s.z[pidx] = kDoneZ;
        } else if (n_lights == 0 || (kChecks && p.gbuf_only)) {
            s.w[pidx] = quantise(__float_as_uint(__ldg(&p.texel_tab[w]).w), amb_only);
            s.z[pidx] = kDoneZ;
        } else {
            gz[m] = s.z[pidx] / kBin;  // ray_bin_z (alternative.cpp:727, C division truncates toward zero)
        }
    }

    // =============================== G + R, kGroupMax groups at a time ===============================
    int g_done = INT_MIN;  // groups <= g_done are finished
    // Segments a round starts with.  A round that finds more candidate boxes than the de-duplication set
    // holds keeps only its leading segments — the walks of the others are thrown away — so the budget
    // follows the candidate density seen in this tile so far.
    int seg_budget = n_lights > 4 ? kSegStart : kSegMax;
#pragma unroll 1
    for (;;) {
        // ---- smallest unprocessed group ----
        if (tid == 0) {
            s.gmin = kNoGroup;
            s.bits[0] = s.bits[1] = 0u;
            s.more = 0;
        }
        __syncthreads();
        int mine = kNoGroup;
#pragma unroll
        for (int m = 0; m < kPPT; m++)
            if (gz[m] != kNoGroup && gz[m] > g_done) mine = min(mine, gz[m]);
        mine = __reduce_min_sync(0xffffffffu, mine);
        if (lane == 0 && mine != kNoGroup) atomicMin(&s.gmin, mine);
        __syncthreads();
        const int gmin = s.gmin;
        if (gmin == kNoGroup) break;
        // ---- the groups present in [gmin, gmin + 64) ----
        {
            unsigned b0 = 0u, b1 = 0u;
            bool beyond = false;
#pragma unroll
            for (int m = 0; m < kPPT; m++) {
                if (gz[m] == kNoGroup || gz[m] <= g_done) continue;
                const unsigned rel = (unsigned)(gz[m] - gmin);
                if (rel < 32u) b0 |= 1u << rel;
                else if (rel < 64u) b1 |= 1u << (rel - 32u);
                else beyond = true;
            }
            b0 = __reduce_or_sync(0xffffffffu, b0);
            b1 = __reduce_or_sync(0xffffffffu, b1);
            beyond = __any_sync(0xffffffffu, beyond);
            if (lane == 0) {
                if (b0) atomicOr(&s.bits[0], b0);
                if (b1) atomicOr(&s.bits[1], b1);
                if (beyond) s.more = 1;
            }
        }
        __syncthreads();
        const unsigned bits0 = s.bits[0], bits1 = s.bits[1];
        const bool more_windows = s.more != 0;  // else this pass is the tile's last: no further search
        const int n_groups = __popc(bits0) + __popc(bits1);
        auto group_index = [&](int g) -> int {  // index of group g among the present ones, -1 if not in this pass
            if (g == kNoGroup || g <= g_done) return -1;
            const unsigned rel = (unsigned)(g - gmin);
            if (rel < 32u) return __popc(bits0 & ((1u << rel) - 1u));
            if (rel < 64u) return __popc(bits0) + __popc(bits1 & ((1u << (rel - 32u)) - 1u));
            return -1;
        };
        if (tid < kGroupMax) {
            const unsigned present = tid < 32 ? (bits0 >> tid) & 1u : (bits1 >> (tid - 32)) & 1u;
            if (present) {
                Grp& g = s.grp[group_index(gmin + tid)];
                g.gz = gmin + tid;
                g.n = 0;
                g.omin[0] = g.omin[1] = g.omin[2] = INT_MAX;
                g.omax[0] = g.omax[1] = g.omax[2] = INT_MIN;
            }
        }
        __syncthreads();
        // ---- pixels per group and bounds of the ray origins (alternative.cpp:720-722).  Per warp: one
        //      iteration per distinct group among its 320 pixels; every lane folds its own pixels of that
        //      group, one redux per quantity, one lane does the shared-memory atomics ----
        int gi[kPPT], zz[kPPT];
        unsigned mine0 = 0u, mine1 = 0u;  // groups this thread's pixels belong to
#pragma unroll
        for (int m = 0; m < kPPT; m++) {
            gi[m] = group_index(gz[m]);
            zz[m] = s.z[(rsub + 4 * m) * kBin + col];
            if (gi[m] >= 0) {
                if (gi[m] < 32) mine0 |= 1u << gi[m];
                else mine1 |= 1u << (gi[m] - 32);
            }
        }
        const unsigned warp0 = __reduce_or_sync(0xffffffffu, mine0), warp1 = __reduce_or_sync(0xffffffffu, mine1);
        const int wj_top = (short)(d.H - (ty * kBin + rsub));  // world_j of pixel m is wj_top - 4m
        for (int half = 0; half < 2; half++) {
            unsigned todo = half ? warp1 : warp0;
            while (todo) {
                const int g = 32 * half + __ffs(todo) - 1;
                todo &= todo - 1;
                int cnt = 0, zlo = INT_MAX, zhi = INT_MIN, ylo = INT_MAX, yhi = INT_MIN;
#pragma unroll
                for (int m = 0; m < kPPT; m++)
                    if (gi[m] == g) {
                        cnt++;
                        // x = column, z from the record, y = world_j - z (quirk Q11: y + z == H - row)
                        const int y = wj_top - 4 * m - zz[m];
                        zlo = min(zlo, zz[m]);
                        zhi = max(zhi, zz[m]);
                        ylo = min(ylo, y);
                        yhi = max(yhi, y);
                    }
                const int xlo = __reduce_min_sync(0xffffffffu, cnt ? i : INT_MAX);
                const int xhi = __reduce_max_sync(0xffffffffu, cnt ? i : INT_MIN);
                cnt = __reduce_add_sync(0xffffffffu, cnt);
                zlo = __reduce_min_sync(0xffffffffu, zlo);
                zhi = __reduce_max_sync(0xffffffffu, zhi);
                ylo = __reduce_min_sync(0xffffffffu, ylo);
                yhi = __reduce_max_sync(0xffffffffu, yhi);
                if (lane == 0) {
                    Grp& G = s.grp[g];
                    atomicAdd(&G.n, cnt);
                    atomicMin(&G.omin[0], xlo);
                    atomicMax(&G.omax[0], xhi);
                    atomicMin(&G.omin[1], ylo);
                    atomicMax(&G.omax[1], yhi);
                    atomicMin(&G.omin[2], zlo);
                    atomicMax(&G.omax[2], zhi);
                }
            }
        }
        // With one or two groups a round covers (nearly) every pixel of the tile anyway: the shade phase then
        // takes the pixels in raster order and no per-group lists are built.
        const bool natural = n_groups <= 2;
        __syncthreads();
        if (!natural) {
        if (tid < 32) {  // exclusive scan of the group sizes (two groups per lane)
            const int a = tid < n_groups ? s.grp[tid].n : 0, b = tid + 32 < n_groups ? s.grp[tid + 32].n : 0;
            int ia = a, ib = b;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
                if (tid >= o) {
                    ia += ta;
                    ib += tb;
                }
            }
            const int total_a = __shfl_sync(0xffffffffu, ia, 31);
            if (tid < n_groups) s.grp[tid].p0 = s.cursor[tid] = ia - a;
            if (tid + 32 < n_groups) s.grp[tid + 32].p0 = s.cursor[tid + 32] = total_a + ib - b;
        }
        __syncthreads();
        for (int half = 0; half < 2; half++) {  // scatter the pixel indices into the per-group lists
            unsigned todo = half ? warp1 : warp0;
            while (todo) {
                const int g = 32 * half + __ffs(todo) - 1;
                todo &= todo - 1;
                int c = 0;
#pragma unroll
                for (int m = 0; m < kPPT; m++) c += (gi[m] == g);
                int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                int base = 0;
                if (lane == 31) base = atomicAdd(&s.cursor[g], incl);
                int at = __shfl_sync(0xffffffffu, base, 31) + incl - c;
#pragma unroll
                for (int m = 0; m < kPPT; m++)
                    if (gi[m] == g) s.pix[at++] = (unsigned short)((rsub + 4 * m) * kBin + col);
            }
        }
        }  // !natural
        // (the barrier at the top of the first round publishes pix and the group table)
        mark(kPhGroup);

        // ---- rounds over (group, light, step range) segments; segment index = group * n_lights + light ----
        const int n_seg_total = n_groups * n_lights;
        int s_cur = 0, ka_cur = 0;  // next unprocessed step of segment s_cur
        int step_cap = INT_MAX;     // most steps of the first segment a round tries (INT_MAX = the whole walk)
        int nseg_try = seg_budget;
#pragma unroll 1
        while (s_cur < n_seg_total) {
            __syncthreads();  // previous round fully consumed (lists, segments, pixel lists complete)
            // A. describe the trial segments (walk set-up, alternative.cpp:406-430)
            int nseg = min(nseg_try, n_seg_total - s_cur);
            if (step_cap != INT_MAX) {
                // Only the LAST segment of a round may cover part of its walk (the advance below relies on it), and the
                // step cap applies to the first one: a round whose first segment is cut short holds nothing else.
                const int grp0 = s_cur / n_lights;
                const short4 lt0 = p.lights[s_cur - grp0 * n_lights];
                const float dx0 = (float)(lt0.x / kBin) - (float)bx, dy0 = (float)((d.H - lt0.y - lt0.z) / kBin) - (float)ty;
                const float dz0 = (float)(lt0.z / kBin) - (float)s.grp[grp0].gz;
                const int steps0 = (int)fmaxf(fmaxf(fabsf(dx0), fabsf(dy0)), fabsf(dz0));
                if (step_cap < steps0 - ka_cur) nseg = 1;
            }
            if (tid < 32) {  // warp 0: one lane per segment (kSegMax == 32)
                int my_steps = 0;
                if (tid < nseg) {
                    const int sidx = s_cur + tid;
                    const int grp = sidx / n_lights, l = sidx - grp * n_lights;
                    const Grp& G = s.grp[grp];
                    const short4 lt = p.lights[l];
                    // light bin, alternative.cpp:729-732 ('/' truncates toward zero)
                    const int lbx = lt.x / kBin, lby = (d.H - lt.y - lt.z) / kBin, lbz = lt.z / kBin;
                    const float dx = (float)lbx - (float)bx, dy = (float)lby - (float)ty, dz = (float)lbz - (float)G.gz;
                    const float big = fmaxf(fmaxf(fabsf(dx), fabsf(dy)), fabsf(dz));
                    Seg& g = s.seg[tid];
                    g.steps = (int)big;  // 0 when big < 1 (then the NaN step is never used)
                    g.sx = dx / big;
                    g.sy = dy / big;
                    g.sz = dz / big;
                    g.light = l;
                    g.grp = grp;
                    g.start = flat_bin(d, bx, ty, G.gz);  // start bin of every pixel of the group (alternative.cpp:724-727)
                    g.ka = tid == 0 ? ka_cur : 0;
                    g.kb = (tid == 0 && step_cap < g.steps - g.ka) ? g.ka + step_cap : g.steps;
                    g.count = 0;
                    g.occ = 0;
                    g.fill = 0;
                    g.octant = group_octant(G, lt);
                    my_steps = g.kb - g.ka;
                }
                // run length: about one work item per thread, so the serial part of a walk stays short
                const int steps = __reduce_add_sync(0xffffffffu, my_steps);
                const int run = min(kMaxRun, max(1, (steps + kT - 1) / kT));
                const int my_items = (my_steps + run - 1) / run;
                int incl = my_items;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (tid >= o) incl += t;
                }
                if (tid < nseg) s.seg[tid].item0 = incl - my_items;
                if (tid == 31) {
                    s.n_items = incl;
                    s.run = run;
                    s.n_occ = 0;
                    s.overflow = 0;
                    s.n_keys = 0;
                }
            }
            if (tid >= 32 && tid < 32 + nseg) {  // warp 1, beside warp 0: the box-independent part of the shaft cull
                const int sidx = s_cur + tid - 32;
                const int grp = sidx / n_lights;
                const Grp& G = s.grp[grp];
                const short4 lt = p.lights[sidx - grp * n_lights];
                // Bounds of the group's ray origins.  The origin is cast to short in the reference
                // (alternative.cpp:720-722): only cull when nothing can wrap.
                bool can_cull = !(p.debug_flags & 1);
                float ol[3], oh[3], rl[3], rh[3];
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    ol[a] = (float)G.omin[a];
                    oh[a] = (float)G.omax[a];
                    can_cull = can_cull && G.omin[a] >= -32768 && G.omax[a] <= 32767;
                }
                const float lp[3] = {(float)lt.x, (float)lt.y, (float)lt.z};
                const unsigned free_axes = shaft_prepare(lp, ol, oh, rl, rh);
                Seg& g = s.seg[tid - 32];
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    g.rl[a] = rl[a];
                    g.rh[a] = rh[a];
                }
                g.cull = can_cull ? (int)free_axes : 8;
            }
            for (int t = tid; t < kHashSize; t += kT) s.r.hash[t] = kEmpty;
            __syncthreads();
            mark(kPhSetup);

            // C. walk.  One thread per run of kRun steps of one segment.  The run is processed in
            // sub-chunks: first the distinct probed bins of up to 8 steps are listed (ALU only) in a
            // thread-private column of shared scratch, then their 4-bit counts are fetched as one batch
            // of independent loads.
            const int n_items = s.n_items, kRun = s.run;
            int* scratch = reinterpret_cast<int*>(s.r.list);  // [kScratchRows][kT]; the box list is idle now
            for (int it = tid; it < n_items; it += kT) {
                int q = 0;
                while (q + 1 < nseg && s.seg[q + 1].item0 <= it) q++;
                const Seg& g = s.seg[q];
                const int k0 = g.ka + (it - g.item0) * kRun, k1 = min(k0 + kRun, g.kb);
                const float sx = g.sx, sy = g.sy, sz = g.sz;
                const int start = g.start;
                // sequential fp32 accumulation from the start bin (quirk Q15)
                float px = (float)bx, py = (float)ty, pz = (float)s.grp[g.grp].gz;
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
                            scratch[n * kT + tid] = f;
                            n++;
                        }
                        x0 = x1;
                        y0 = y1;
                        z0 = z1;
                    }
                    // the 4-bit counts of the sub-chunk, in batches of 8 independent loads; occupied bins
                    // go straight to the shared list
                    int boxes = 0, bins = 0;
                    for (int i0 = 0; i0 < n; i0 += 8) {
                        int f[8], c[8];
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            f[u] = i0 + u < n ? scratch[(i0 + u) * kT + tid] : -1;
                            c[u] = f[u] >= 0 ? (__ldg(&p.occ4[f[u] >> 3]) >> ((f[u] & 7) * 4)) & 7 : 0;
                        }
#pragma unroll
                        for (int u = 0; u < 8; u++)
                            if (c[u]) {
                                const int o = atomicAdd(&s.n_occ, 1);  // (warp-aggregating these was 7 % slower)
                                if (o < kOccCap) {
                                    s.r.occ_bin[o] = (unsigned)f[u];
                                    s.r.occ_meta[o] = (unsigned char)(q << 3 | c[u]);
                                }
                                boxes += c[u];
                                bins++;
                            }
                    }
                    if (boxes) {
                        atomicAdd(&s.seg[q].count, boxes);
                        atomicAdd(&s.seg[q].occ, bins);
                    }
                }
            }
            __syncthreads();
            mark(kPhWalk);
            const int n_occ = s.n_occ;  // every occupied bin holds >= 1 box, so n_occ <= sum of counts
            // D. how many leading segments go into this round?  (every thread, redundantly)
            // The occupied-bin list must be complete for them.  Candidates do not limit a round: the gather culls
            // before it de-duplicates, so the set only ever holds survivors (its fill and the box list are checked
            // while gathering; a round that overflows either is redone with half the segments).
            constexpr int kSetCap = kHashSize * 3 / 4;
            int n_fit = 0, cand_total = 0, occ_total = 0;
            while (n_fit < nseg && occ_total + s.seg[n_fit].occ <= kOccCap) {
                cand_total += s.seg[n_fit].count;
                occ_total += s.seg[n_fit].occ;
                n_fit++;
            }
            // A trial that does not fit is walked again — with what the failed walk measured, so that one more
            // walk is enough: the leading segments that do fit, or, when even the first one alone is too much, the
            // part of its step range that its occupied-bin density allows.
            if (n_fit == 0) {
                const int range = s.seg[0].kb - s.seg[0].ka;
                const float need = (float)s.seg[0].occ / (float)kOccCap;
                step_cap = max(1, min(range - 1, (int)((float)range * 0.85f / need)));
                nseg_try = 1;
                if (kChecks && p.phase_cycles && tid == 0) atomicAdd(&p.phase_cycles[13], 1ull);  // debug: walks thrown away
                continue;
            }
            if (n_occ > kOccCap) {  // the occupied-bin list is incomplete (which bins got a slot is arbitrary)
                nseg_try = n_fit;
                if (kChecks && p.phase_cycles && tid == 0) atomicAdd(&p.phase_cycles[15], 1ull);
                continue;
            }
            // List room of a segment: in proportion to its candidates (a segment keeps a small, similar fraction of
            // them after de-duplication and cull); the sum stays within the list.
            auto share_of = [&](int q) -> int {  // (candidates per round stay far below 2^32 / kListCap)
                const unsigned c = (unsigned)s.seg[q].count;
                return (int)min(c, (unsigned)(kListCap - n_fit) * c / (unsigned)max(cand_total, 1) + 1u);
            };
            if (tid < n_fit) {  // base of segment tid in the box list (read after the next barrier)
                int base = 0;
                for (int q = 0; q < tid; q++) base += share_of(q);
                s.seg[tid].base = base;
            }

            // E. gather: one lane per candidate slot.  Each warp takes 32 occupied bins, scans their
            // counts and expands them into (bin, slot) pairs with shuffles, so that the dependent
            // loads (entity id -> box) run with dense lanes: entity -> de-duplicate -> box ->
            // shaft cull -> the segment's part of the box list.
            const int seg_share = lane < n_fit ? share_of(lane) : 0;  // lane q: list room of segment q ...
            int seg_base = seg_share;                                  // ... and its first list slot
            {
                const int own = seg_base;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, seg_base, o);
                    if (lane >= o) seg_base += t;
                }
                seg_base -= own;
            }
            for (int ob = tid - lane; ob < n_occ; ob += kT) {
                const bool have = ob + lane < n_occ;
                const unsigned my_bin = have ? s.r.occ_bin[ob + lane] : 0u;
                const unsigned my_meta = have ? s.r.occ_meta[ob + lane] : 0u;
                const int my_c = (have && (int)(my_meta >> 3) < n_fit) ? (int)(my_meta & 7u) : 0;
                int incl = my_c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                // One candidate: where it sits (a binary search over the warp's prefix sums, all lanes together), then
                // entity id -> box -> cull -> de-duplicate -> list.
                struct Cand {
                    int ent, q, base, room;
                };
                auto locate = [&](int t) -> Cand {  // convergent; ent < 0: no candidate t
                    int src = 0;  // first lane whose inclusive prefix exceeds t
#pragma unroll
                    for (int step = 16; step; step >>= 1) {
                        const int v = __shfl_sync(0xffffffffu, incl, src + step - 1);
                        if (v <= t) src += step;
                    }
                    src = min(src, 31);
                    const unsigned bin = __shfl_sync(0xffffffffu, my_bin, src);
                    const unsigned meta = __shfl_sync(0xffffffffu, my_meta, src);
                    const int cnt_src = __shfl_sync(0xffffffffu, my_c, src);
                    const int slot_i = t - (__shfl_sync(0xffffffffu, incl, src) - cnt_src);
                    Cand c;
                    c.q = meta >> 3;
                    c.base = __shfl_sync(0xffffffffu, seg_base, c.q & 31);
                    c.room = __shfl_sync(0xffffffffu, seg_share, c.q & 31);
                    c.ent = t < total ? __ldg(&p.ids[(size_t)bin * kSlots + slot_i]) : -1;
                    return c;
                };
                auto process = [&](const Cand& c, const int4 raw) {
                    const int q = c.q, ent = c.ent;
                    const Seg& sg = s.seg[q];
                    const Grp& G = s.grp[sg.grp];
                    const Box b = unpack_box(raw);
                    if (!(sg.cull & 8)) {  // shaft cull: no ray of the group can hit this box
                        const float blo[3] = {(float)b.px, (float)b.py, (float)b.pz};
                        const float bhi[3] = {(float)(b.px + b.ex), (float)(b.py + b.ey), (float)(b.pz + b.ez)};
                        const float ol[3] = {(float)G.omin[0], (float)G.omin[1], (float)G.omin[2]};
                        const float oh[3] = {(float)G.omax[0], (float)G.omax[1], (float)G.omax[2]};
                        const float rl[3] = {sg.rl[0], sg.rl[1], sg.rl[2]}, rh[3] = {sg.rh[0], sg.rh[1], sg.rh[2]};
                        if (!shaft_may_hit_prepared(blo, bhi, ol, oh, rl, rh, (unsigned)sg.cull)) return;
                    }
                    // de-duplicate the survivors: the set stays small however many candidates the walks find
                    if (*(volatile int*)&s.overflow) return;  // (the set may be filling up: the round is redone anyway)
                    const unsigned key = (unsigned)q << 26 | (unsigned)ent;
                    unsigned h = (key * 2654435761u) >> (32 - kHashBits);
                    for (;;) {
                        const unsigned old = atomicCAS(&s.r.hash[h], kEmpty, key);
                        if (old == key) return;  // (segment, entity) seen before in this round
                        if (old == kEmpty) break;
                        h = (h + 1) & (kHashSize - 1);
                    }
                    if (atomicAdd(&s.n_keys, 1) >= kHashSize * 3 / 4) {
                        s.overflow = 1;
                        return;
                    }
                    if (kChecks && p.phase_cycles) atomicAdd(&p.phase_cycles[9], 1ull);  // debug: distinct survivors
                    const int nth = atomicAdd(&s.seg[q].fill, 1);
                    if (nth >= c.room) {
                        s.overflow = 1;
                        return;
                    }
                    store_box(s.r.list, c.base + nth, b, ent, sg.octant);
                };
                for (int t0 = 0; t0 < total; t0 += 32) {  // (two candidates per lane in flight was 4 % slower)
                    const Cand c = locate(t0 + lane);
                    if (c.ent >= 0) process(c, __ldg(&p.boxes[c.ent]));
                }
            }
            __syncthreads();
            mark(kPhGather);
            if (s.overflow) {  // some segment kept more than its share: fewer segments, then fewer steps
                if (kChecks && p.phase_cycles && tid == 0) atomicAdd(&p.phase_cycles[14], 1ull);  // debug: walks + gathers thrown away
                if (n_fit > 1) {
                    nseg_try = max(1, n_fit / 2);
                } else {
                    step_cap = max(1, (s.seg[0].kb - s.seg[0].ka) / 2);
                    nseg_try = 1;
                }
                continue;
            }
            if (kChecks && p.phase_cycles && tid == 0) {  // debug: candidate boxes found / kept after de-dup + cull
                unsigned long long found = 0, kept = 0;
                for (int q = 0; q < n_fit; q++) {
                    found += s.seg[q].count;
                    kept += s.seg[q].fill;
                }
                atomicAdd(&p.phase_cycles[10], found);
                atomicAdd(&p.phase_cycles[11], kept);
                atomicAdd(&p.phase_cycles[12], 1ull);
            }

            // F. shade: one lane per pixel of the groups this round touches
            const int grp_first = s.seg[0].grp, grp_last = s.seg[n_fit - 1].grp;
            const int pix0 = s.grp[grp_first].p0, pix1 = s.grp[grp_last].p0 + s.grp[grp_last].n;
            const int n_slots = natural ? kPix : pix1 - pix0;
            const bool last_seg_done = s.seg[n_fit - 1].kb == s.seg[n_fit - 1].steps;
            for (int qb = tid - lane; qb < n_slots; qb += kT) {
                const int pidx = natural ? qb + lane : (qb + lane < n_slots ? s.pix[pix0 + qb + lane] : 0);
                const int row = pidx / kBin;
                const int j = ty * kBin + row, ipx = bx * kBin + (pidx - row * kBin);
                const int g_z = s.z[pidx], g_ent = s.ent[pidx];
                const unsigned g_w = s.w[pidx];
                const int g_y = (short)(d.H - j) - g_z;  // quirk Q11
                int grp = group_index(g_z / kBin);
                const bool valid = natural ? (g_z != kDoneZ && grp >= grp_first && grp <= grp_last) : qb + lane < n_slots;
                if (!valid) grp = grp_first;
                // the pixel's segments in this round: [qa, qe)
                const int qa = max(0, grp * n_lights - s_cur), qe = min(n_fit, (grp + 1) * n_lights - s_cur);
                const bool fresh = grp * n_lights >= s_cur && s.seg[qa].ka == 0;        // light 0 starts here
                const bool final_round = (grp + 1) * n_lights - s_cur <= n_fit && (grp != grp_last || last_seg_done);
                // raster row, or the row's slot in the stripe-major staging frame (rank-contiguous)
                const int jo = p.out_stripe_T ? ((ty % d.stripe_n) * p.out_stripe_T + ty / d.stripe_n) * kBin + row : j;
                const size_t at = (size_t)jo * d.W + ipx;
                float4 tex = make_float4(0.f, 0.f, 0.f, 0.f);
                float acc = 0.f;
                bool shadowed = false;
                if (valid) {
                    tex = __ldg(&p.texel_tab[g_w]);  // normal + palette colour of the hit texel
                    if (!fresh) {  // state parked by the previous round (sign bit: shadowed so far by a split light)
                        const unsigned park = reinterpret_cast<const unsigned*>(p.out)[at];
                        acc = __uint_as_float(park & 0x7fffffffu);
                        shadowed = park >> 31;
                    }
                }
                // Ray origin, alternative.cpp:720-722
                const float ox = (float)(short)ipx, oy = (float)(short)g_y, oz = (float)(short)g_z;
                const int wqa = __reduce_min_sync(0xffffffffu, valid ? qa : n_fit);
                const int wqe = __reduce_max_sync(0xffffffffu, valid ? qe : 0);
                for (int q = wqa; q < wqe; q++) {
                    const Seg& sg = s.seg[q];
                    const bool mine = valid && q >= qa && q < qe;
                    const short4 lt = p.lights[sg.light];
                    if (sg.ka == 0 && mine) shadowed = false;
                    // towards_light, L1-normalised (alternative.cpp:711-715, sprites.hpp:28-35)
                    float tx = (float)(lt.x - ipx), tyv = (float)(lt.y - g_y), tz = (float)(lt.z - g_z);
                    const float len = fabsf(tx) + fabsf(tyv) + fabsf(tz);
                    tx = tx / len;
                    tyv = tyv / len;
                    tz = tz / len;
                    // alternative.cpp:745-747; a term of 0 adds +0 whether visible or not (Q19)
                    const float lam = std_max(0.f, tex.x * tx + tex.y * tyv + tex.z * tz);
                    if (kChecks && p.dbg_t && mine && sg.light == p.dbg_light && sg.ka == 0)
                        p.dbg_t[(size_t)j * d.W + ipx] = make_float4(tx, tyv, tz, lam);
                    const bool lit_candidate = mine && lam > 0.f;
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
                    const float4* boxes = s.r.list + 2 * sg.base;
                    if (__any_sync(0xffffffffu, test)) {
                        // warp-uniform choice of the slab-test variant
                        bool hit = false;
                        if (sg.octant >= 0) {
                            if (test) hit = any_box_hit<0>(boxes, n, g_ent, ox, oy, oz, ix, iy, iz);
                        } else if (!__any_sync(0xffffffffu, test && !nan_free)) {
                            if (test) hit = any_box_hit<1>(boxes, n, g_ent, ox, oy, oz, ix, iy, iz);
                        } else {
                            if (test) hit = any_box_hit<2>(boxes, n, g_ent, ox, oy, oz, ix, iy, iz);
                        }
                        if (hit) shadowed = true;
                    }
                    if (sg.kb == sg.steps && lit_candidate && !shadowed) acc = acc + lam;
                }
                if (valid && qa < qe) {
                    if (final_round) {
                        if (kChecks && p.dbg_factor) p.dbg_factor[(size_t)j * d.W + ipx] = acc + p.ambient;
                        // alternative.cpp:735 / 757-758
                        s.w[pidx] = quantise(__float_as_uint(tex.w), std_min(1.f, acc + p.ambient));
                        s.z[pidx] = kDoneZ;
                    } else {
                        reinterpret_cast<unsigned*>(p.out)[at] = __float_as_uint(acc) | (shadowed ? 0x80000000u : 0u);
                    }
                }
            }
            mark(kPhShade);

            // advance past the processed segments
            if (last_seg_done) {
                s_cur += n_fit;
                ka_cur = 0;
            } else {  // a split light: n_fit == 1
                s_cur += n_fit - 1;
                ka_cur = s.seg[n_fit - 1].kb;
            }
            if (s_cur < n_seg_total) {  // next round's budgets, from this round's densities (aiming well below the
                                        // capacities: densities vary a lot from light to light)
                int max_fill = 1, total_fill = 0;
                for (int q = 0; q < n_fit; q++) {
                    max_fill = max(max_fill, s.seg[q].fill);
                    total_fill += s.seg[q].fill;
                }
                const int by_set = (kSetCap * 7 / 8) * n_fit / max(s.n_keys, 1);
#ifndef PAR_TILE_OCC_FILL
#define PAR_TILE_OCC_FILL 4  // eighths of the occupied-bin list a round aims at (densities vary a lot from light to light)
#endif
                const int by_occ = (kOccCap * PAR_TILE_OCC_FILL / 8) * n_fit / max(occ_total, 1);
                // (list room is shared out in proportion to the candidates: what counts is the total kept, with headroom
                // for segments that keep a larger fraction than the others)
#ifndef PAR_TILE_LIST_FILL
#define PAR_TILE_LIST_FILL 2  // eighths of the box list
#endif
                const int by_list = (kListCap * PAR_TILE_LIST_FILL / 8) * n_fit / max(total_fill, 1);
                seg_budget = max(1, min(kSegMax, min(min(by_set, by_occ), by_list)));
                // a split light goes on with the step range its density so far allows; a new light tries its whole walk
                step_cap = INT_MAX;
                if (!last_seg_done) {
                    const Seg& g = s.seg[n_fit - 1];
                    const float need = fmaxf((float)g.occ / (float)kOccCap, (float)(g.fill + g.fill / 4 + 1) / (float)kListCap);
                    step_cap = max(1, (int)((float)(g.kb - g.ka) * 0.85f / fmaxf(need, 1e-3f)));
                }
            }
            nseg_try = seg_budget;
        }
        if (!more_windows) break;
        g_done = gmin + (kGroupMax - 1);
    }

    // ---- 16-byte stores of the finished tile rows ----
    __syncthreads();
    if (!(kChecks && p.gbuf_only)) {
        for (int v = tid; v < kPix / 4; v += kT) {
            const int row = v / (kBin / 4), j = ty * kBin + row;
            if (j < ra || j >= rb) continue;
            const uint4 px = *reinterpret_cast<const uint4*>(&s.w[4 * v]);
            const int jo = p.out_stripe_T ? ((ty % d.stripe_n) * p.out_stripe_T + ty / d.stripe_n) * kBin + row : j;
            const size_t at = (size_t)jo * d.W + bx * kBin + 4 * (v % (kBin / 4));
            *reinterpret_cast<uint4*>(&p.out[at]) = px;
            // Fused exchange: the same chunk goes straight into every peer GPU's frame (posted writes over
            // NVLink), so no all-gather pass over the frame is needed afterwards.
            for (int r = 0; r < p.n_peer_out; r++) *reinterpret_cast<uint4*>(&p.peer_out[r][at]) = px;
        }
    }
    mark(kPhTail);
    if (p.tile_cost && tid == 0) {
        const long long dt = clock64() - t_begin;
        p.tile_cost[ty * d.HW + bx] = (unsigned)min(dt, (long long)0xffffffffu);
    }
}

#ifdef PAR_TILE_ONE_LIGHT
// ---- the one-light configuration (see the top of the file): production frames only ----
size_t tile_one_light_smem_bytes() { return sizeof(TileSmem); }

cudaError_t configure_tile_one_light() {
    return cudaFuncSetAttribute(k_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem));
}

int tile_one_light_ctas_per_sm() {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_tile<false>, kT, sizeof(TileSmem)) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    return n;
}

cudaError_t launch_tile_one_light(const TileParams& p, int n_tiles, cudaStream_t st) {
    k_tile<false><<<n_tiles, kT, sizeof(TileSmem), st>>>(p);
    return cudaGetLastError();
}

}  // namespace par
#else
size_t tile_smem_bytes() { return sizeof(TileSmem); }

cudaError_t configure_tile() {
    cudaError_t e = cudaFuncSetAttribute(k_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem));
    if (e != cudaSuccess) return e;
    return configure_tile_one_light();
}

// CTAs of the production render kernel one SM holds at a time (occupancy query; 5 on sm_100).
int tile_ctas_per_sm() {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_tile<false>, kT, sizeof(TileSmem)) != cudaSuccess) {
        cudaGetLastError();
        n = PAR_TILE_MIN_CTAS;
    }
    return n > 0 ? n : 1;
}

cudaError_t launch_tile(const TileParams& p, cudaStream_t st) {
    int first, tile_rows;
    owned_tile_rows(p.d, first, tile_rows);
    if (tile_rows <= 0) return cudaSuccess;
    const int n_tiles = tile_rows * tiles_per_stripe(p.d);  // (tile_rows counts stripes)
    if (p.gbuf || p.gbuf_only || p.dbg_t || p.dbg_factor || p.phase_cycles)
        k_tile<true><<<n_tiles, kT, sizeof(TileSmem), st>>>(p);
    else if (p.one_light_config)
        return launch_tile_one_light(p, n_tiles, st);
    else
        k_tile<false><<<n_tiles, kT, sizeof(TileSmem), st>>>(p);
    return cudaGetLastError();
}

// ---- tile order: longest tiles first ---------------------------------------------------------------
// The cost of a tile (shadow-walk lengths, occluder counts) varies by more than an order of
// magnitude over a many-light frame, and the hardware hands CTAs to SMs in launch order: with the
// raster order a few expensive tiles that happen to start late leave most SMs idle at the end of the
// kernel.  Every CTA records its cycle count; before the next frame the owned tiles are
// counting-sorted by the previous frame's cost, most expensive first (LPT scheduling).  The order
// only permutes which CTA renders which tile, never what is rendered.
__global__ void __launch_bounds__(1024)
k_tile_order(const unsigned* __restrict__ cost, int* __restrict__ order, ViewDims d, int tile_row_first, int tile_rows) {
    constexpr int kBuckets = 256;
    __shared__ unsigned s_max;
    __shared__ int s_hist[kBuckets], s_base[kBuckets];
    const int tps = tiles_per_stripe(d), seg = stripe_segments(d);
    const int n = tile_rows * tps, tid = threadIdx.x;
    const int stripe = max(d.stripe_n, 1);
    auto cost_of = [&](int t) {  // owned tile t (the CTA slot numbering of k_tile) -> its cost
        const int v = tile_row_first + (t / tps) * stripe;
        return cost[(v / seg) * d.HW + stripe_column_segment(d, v) * tps + t % tps];
    };
    if (tid == 0) s_max = 1u;
    if (tid < kBuckets) s_hist[tid] = 0;
    __syncthreads();
    unsigned mx = 0u;
    for (int t = tid; t < n; t += blockDim.x) mx = max(mx, cost_of(t));
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((tid & 31) == 0) atomicMax(&s_max, mx);
    __syncthreads();
    const float scale = (float)(kBuckets - 1) / (float)s_max;
    auto bucket = [&](int t) {  // bucket 0 = most expensive
        const unsigned c = cost_of(t);
        return (kBuckets - 1) - min(kBuckets - 1, (int)((float)c * scale));
    };
    for (int t = tid; t < n; t += blockDim.x) atomicAdd(&s_hist[bucket(t)], 1);
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int b = 0; b < kBuckets; b++) {
            s_base[b] = run;
            run += s_hist[b];
        }
    }
    __syncthreads();
    for (int t = tid; t < n; t += blockDim.x) order[atomicAdd(&s_base[bucket(t)], 1)] = t;
}

cudaError_t launch_tile_order(const unsigned* cost, int* order, const ViewDims& d, cudaStream_t st) {
    int first, tile_rows;
    owned_tile_rows(d, first, tile_rows);
    if (tile_rows <= 0) return cudaSuccess;
    k_tile_order<<<1, 1024, 0, st>>>(cost, order, d, first, tile_rows);
    return cudaGetLastError();
}

}  // namespace par
#endif  // PAR_TILE_ONE_LIGHT
