// par_kernels.cuh — parameter blocks and host-side launchers of the three kernel groups
// (scene_loader.cu, primary.cu, shade.cu); shared with the C ABI layer (par_api.cu).
#pragma once
#include "par_device.cuh"

namespace par {

// ---- scene loader (alternative.cpp:690-693, 195-269) ----
struct LoaderCounters {
    int n_survivors;
    int n_inserts;
    int reserved;
    int bad_scene;  // set when an AABB would index outside the 20x40 sprite (quirk Q7)
};
cudaError_t launch_scene_loader(const int4* raw, const int* sprite_ids, int n, int n_sprites,
                                const ViewDims& d, int4* boxes, int* cnt, int* ids,
                                unsigned* occ4, int* survivors, LoaderCounters* ctr,
                                LoaderCounters* host_a, LoaderCounters* host_b,
                                cudaStream_t s, int* launches);

// ---- per-tile shadow-walk work descriptors (primary -> walks -> shade) ----
constexpr int kMaxGroups = 24;    // z-groups per tile with precomputed walks (more: shade walks itself)
constexpr int kWalkListCap = 16;  // boxes kept per (tile, group, light) list in the pool
struct GroupMeta {                // one z-group of a tile: all its pixels start their shadow walk in
    int z;                        //   bin (tile x, tile y, z)  (alternative.cpp:724-727, quirk Q11)
    int npix;
    int omin[3], omax[3];         // integer bounds of the group's ray origins (alternative.cpp:720-722)
};

// ---- primary rays (alternative.cpp:271-383) ----
struct PrimaryParams {
    ViewDims d;
    const int* cnt;
    const int* ids;
    const int4* boxes;
    const int* atlas_depth;  // [n_sprites][800]
    int n_sprites;
    int4* gbuf;
    int tile_row_first;  // first owned tile row (bin_y); the next ones are stripe_n apart
    int* tile_ngroups;   // [HW*HH]: z-groups of the tile (ascending z), -1 = more than kMaxGroups
    GroupMeta* groups;   // [HW*HH][kMaxGroups]
};
size_t primary_smem_bytes(const ViewDims& d, int n_sprites);
cudaError_t configure_primary(size_t smem);  // per device, before the first launch
cudaError_t launch_primary(const PrimaryParams& p, cudaStream_t s);

// ---- shadow walks (alternative.cpp:399-476 minus the slab tests): one warp per (tile, group, light) ----
struct WalkParams {
    ViewDims d;
    const int* ids;
    const unsigned* occ4;
    const int4* boxes;
    const int* tile_ngroups;
    const GroupMeta* groups;
    int2* table;       // [HW*HH][kMaxGroups][n_lights]: (pool offset, box count) or count -1 = not available
    int4* pool;        // kept boxes: .x.y.z = the packed 16-bit box record, .w = entity index
    int* pool_cursor;
    int pool_cap;
    int n_lights;
    int tile_row_first;
    int debug_flags;
    short4 lights[64];
};
cudaError_t launch_walks(const WalkParams& p, cudaStream_t s);

// ---- shading + shadow rays + RGBA8 pack (alternative.cpp:702-760, 399-500, 40-83) ----
constexpr int kMaxLights = 64;
struct ShadeParams {
    ViewDims d;
    const int* cnt;
    const int* ids;
    const unsigned* occ4;      // 4 bits per bin: cnt & 7 (8 bins per word)
    const int4* boxes;
    const int4* gbuf;
    const float* atlas_normal;         // [n_sprites][800][3]
    const unsigned char* atlas_color;  // [n_sprites][800] palette index
    const uchar4* palette;
    uchar4* out;  // full frame, W*H
    int n_lights;
    float ambient;
    int tile_row_first;
    int out_stripe_T;  // 0: raster output; T > 0: stripe-major staging, T stripes per rank
    int n_peer_out;    // fused frame exchange: every finished 16-byte chunk is also stored, in place, into
    uchar4* peer_out[7];  //   the raster frames of the other GPUs (peer memory over NVLink / NVSwitch)
    int debug_flags;   // bit 0: disable the shaft cull, bit 2: ignore precomputed walks (A/B measurements only)
    const int* tile_ngroups;  // precomputed walks (walks.cu); NULL = shade walks itself
    const int2* table;
    const int4* pool;
    unsigned long long* phase_cycles;  // optional debug instrumentation: 16 counters (NULL in production)
    short4 lights[kMaxLights];         // x, y, z, radius (alternative.cpp:619-622)
};
size_t shade_smem_bytes();
cudaError_t configure_shade();  // per device, before the first launch
cudaError_t launch_shade(const ShadeParams& p, cudaStream_t s);

}  // namespace par
