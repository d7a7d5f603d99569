// par_device.cuh — shared device-side definitions of the B200 render path.
//
// Data layout in HBM (all per context):
//   boxes   int4[N]      one 16-byte record per entity = the reference AABB (alternative.cpp:35-38)
//                        with the sprite-atlas index stored in the 4 padding bytes:
//                        .x = px | py<<16, .y = pz | ex<<16, .z = ey | ez<<16, .w = sprite id
//   cnt     int32[V]     number of inserts per bin this frame (the reference's wrapping count is cnt & 7)
//   ids     int32[V*8]   per bin the up-to-7 HIGHEST entity indices that inserted, descending;
//                        the reference's slot s of a bin is ids[bin*8 + (cnt&7) - 1 - s]   (quirk Q2)
//   gbuf    int4[W*H]    compact G-buffer: .x entity, .y world y, .z world z,
//                        .w = texel | sprite<<10, or -1 for a miss pixel
//   frame   uchar4[W*H]  RGBA8
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace par {

constexpr int kBin = 40;        // single_bin_cubic_size, alternative.cpp:116
constexpr int kSlots = 8;       // sparse_bin_size,       alternative.cpp:131
constexpr int kSpriteW = 20;    // alternative.cpp:330
constexpr int kTexels = 800;    // sprites.hpp:68-70
constexpr int kTileThreads = 320;  // 40 columns x 8 rows of a 40x40 screen tile
constexpr int kTileRowsPerThread = 5;

struct ViewDims {
    int W, H, L;     // view_width / view_height / view_length
    int HW, HH, HL;  // hash_width / hash_height / hash_length (alternative.cpp:120-122)
    int V;           // hash_volume
    int row0, row1;  // band rendered by this context
    int stripe_n;    // >= 1: only the stripes v with v % stripe_n == stripe_i are rendered
    int stripe_i;    //       (40-row stripes interleaved over ranks: load balance)
    int stripe_s;    // >= 1: a tile row is cut into stripe_s stripes of HW / stripe_s tiles each; stripe
                     //       v = tile row * stripe_s + segment (equal stripe counts per rank when HH % ranks != 0)
    int stripe_rot;  // lcm(stripe_n, stripe_s): the stripes after which the column segments rotate by one (below)
};

// Stripe v -> its tile row, and the tiles per stripe.
__host__ __device__ __forceinline__ int stripe_segments(const ViewDims& d) { return d.stripe_s > 1 ? d.stripe_s : 1; }
__host__ __device__ __forceinline__ int tiles_per_stripe(const ViewDims& d) { return d.HW / stripe_segments(d); }

// Column segment of stripe v within its tile row v / stripe_s: v % stripe_s rotated by one for every
// lcm(ranks, stripe_s) stripes.  Without the rotation rank v % ranks would own the same columns in every one of
// its tile rows whenever ranks and stripe_s have a common factor (2 stripes per row over 8 ranks: the even ranks
// the left half of the image, the odd ranks the right half — measured 12 % apart on the 8K frames); with it a
// rank's stripes visit all column segments in turn.  All stripes of a tile row share v / lcm (the lcm is a
// multiple of stripe_s), so the segments of a row are permuted, never doubled.
__host__ __device__ __forceinline__ int stripe_column_segment(const ViewDims& d, int v) {
    const int s = stripe_segments(d);
    return s > 1 ? (v % s + v / d.stripe_rot) % s : 0;
}
// ... and back: the stripe that holds column segment `seg` of tile row t.
__host__ __device__ __forceinline__ int stripe_of_segment(const ViewDims& d, int t, int seg) {
    const int s = stripe_segments(d);
    if (s == 1) return t;
    const int rot = (t * s / d.stripe_rot) % s;
    return t * s + (seg - rot + s) % s;
}

// Stripes first, first + stripe_n, ... (count of them) owned by this context within its band.  With stripe_s == 1 a
// stripe IS a tile row.
__host__ __device__ __forceinline__ void owned_tile_rows(const ViewDims& d, int& first, int& count) {
    const int seg = stripe_segments(d);
    const int t0 = (d.row0 / kBin) * seg, t1 = ((d.row1 + kBin - 1) / kBin) * seg;  // band's stripes [t0, t1)
    const int n = d.stripe_n > 1 ? d.stripe_n : 1, i = d.stripe_n > 1 ? d.stripe_i : 0;
    first = t0 + ((i - t0) % n + n) % n;
    count = first < t1 ? (t1 - first + n - 1) / n : 0;
}

// alternative.cpp:180-182
__host__ __device__ __forceinline__ int flat_bin(const ViewDims& d, int x, int y, int z) {
    return x * d.HH * d.HL + y * d.HL + z;
}

struct Box {
    int px, py, pz, ex, ey, ez, sprite;
};

__device__ __forceinline__ Box unpack_box(int4 r) {
    Box b;
    b.px = (short)(r.x & 0xffff);
    b.py = r.x >> 16;
    b.pz = (short)(r.y & 0xffff);
    b.ex = r.y >> 16;
    b.ey = (short)(r.z & 0xffff);
    b.ez = r.z >> 16;
    b.sprite = r.w;
    return b;
}

// Bin ranges an entity is inserted into (alternative.cpp:212-240); false = culled.
struct BinRange {
    int x0, x1, y0, y1, z0, z1;
};

__device__ __forceinline__ bool cull_and_range(const ViewDims& d, const Box& b, BinRange& r) {
    int x0 = b.px, y0 = b.py, z0 = b.pz;
    int x1 = x0 + b.ex, y1 = y0 + b.ey, z1 = z0 + b.ez;
    // quirk Q5: hard-coded slack of one bin
    if (x1 < 0 || x0 >= d.W) return false;
    if (y1 < 0 - z1) return false;
    if (y0 >= d.H - z0 + kBin) return false;
    if (z1 < -b.ez - kBin) return false;
    if (z0 > d.L + kBin) return false;
    // quirk Q4: the grid's y axis is the screen-row axis; '/' truncates toward zero
    r.x0 = max(0, x0 / kBin);
    r.y0 = max(0, (d.H - y1 - z1) / kBin);
    r.z0 = max(0, z0 / kBin);
    r.x1 = min(d.HW, (x1 + kBin - 1) / kBin);
    r.y1 = min(d.HH, (d.H - y0 - z0 + kBin - 1) / kBin);
    r.z1 = min(d.HL, (z1 + kBin - 1) / kBin);
    return true;
}

// std::min / std::max exactly as libstdc++ defines them, including what they do with NaN
// (quirk Q13): min(a,b) = (b<a)?b:a, max(a,b) = (a<b)?b:a.  Never fminf/fmaxf here.
__device__ __forceinline__ float std_min(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float std_max(float a, float b) { return (a < b) ? b : a; }

}  // namespace par
