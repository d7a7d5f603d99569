// par_kernels.cuh — parameter blocks and host-side launchers of the kernel groups
// (scene_loader.cu, tile.cu); shared with the C ABI layer (par_api.cu).
#pragma once
#include "par_device.cuh"

namespace par {

// ---- scene loader (alternative.cpp:690-693, 195-269) ----
struct LoaderCounters {
    int n_survivors;
    int n_inserts;
    int bad_entity;  // first (lowest) entity whose box would index outside its sprite, or INT_MAX
    int bad_scene;   // set when an inserted AABB would index outside its sprite (quirk Q7)
    int n_list;      // entries of the survivor list (>= n_survivors: moved entities are appended)
    int blocks_done; // grid-wide completion counter of the occupancy pass
    int pad[2];
};

// One generation of the view-hash grid.  Two generations alternate: while frame k
// is built into one, the bins the other one's survivors touched are cleared in the same launch.
struct GridBuffers {
    int* cnt;         // [V]   inserts per bin (the reference's wrapping count is cnt & 7)
    int* ids;         // [V*8] the <= 7 highest inserting entity ids per bin, descending; -1 = empty
    unsigned* occ4;   // 4 bits per bin: cnt & 7
    int4* boxes;      // [cap] packed box records this grid was built from (.w = sprite id)
    int* survivors;   // [cap] entities inserted into this grid (clear list for its next reuse)
    LoaderCounters* ctr;
};

struct LoaderParams {
    ViewDims d;
    const int4* raw;         // uploaded AABB records (alternative.cpp:35-38)
    const int* sprite_ids;   // optional
    const int2* sprite_dims; // [n_sprites]: texel base, width | height << 16
    int n, n_sprites;
    GridBuffers cur;         // built by this launch
    GridBuffers old;         // its touched bins are cleared by this launch (old.cnt == NULL: nothing to clear)
    int old_n_list_cap;      // upper bound of old's survivor list (threads to spend on the clear)
    LoaderCounters* host_a;  // page-locked host copies of the counters (either may be NULL)
    LoaderCounters* host_b;
};
cudaError_t launch_scene_loader(const LoaderParams& p, cudaStream_t s, int* launches);
// Clear the bins a generation's last build touched and re-arm its counters (n_list_cap: host-side upper bound
// of its survivor list).
cudaError_t launch_clear_touched(const ViewDims& d, const GridBuffers& g, int n_list_cap, cudaStream_t s, int* launches);
// Whole-grid clear (context creation / recovery): insert totals 0, slots -1, occupancy 0, counters 0.
cudaError_t launch_clear_grid(const GridBuffers& g, int V, cudaStream_t s);

// Incremental update (alternative.cpp:641-681: only entity 0 and light 0 ever move): entities
// [first, first + count) get new boxes; only the bins their old and new boxes span are rebuilt.
constexpr int kMaxUpdate = 8;  // entities per incremental update (more: full rebuild)
struct UpdateParams {
    ViewDims d;
    GridBuffers g;
    const int2* sprite_dims;
    int n, n_sprites;
    int count;                   // updated entities
    int first;                   // index of the first one
    int4 fresh[kMaxUpdate];      // their new packed records (.w = sprite id)
    int keep_sprite_ids;         // 1: .w of fresh is ignored, the entities keep their sprites
    int4* raw;                   // the resident upload buffer is kept in step (later full re-bins read it)
    int* raw_sprite_ids;         // likewise (may be NULL)
    LoaderCounters* host_a;
    LoaderCounters* host_b;
};
cudaError_t launch_scene_update(const UpdateParams& p, cudaStream_t s, int* launches);
cudaError_t launch_publish_counters(const GridBuffers& g, LoaderCounters* host_a, LoaderCounters* host_b, cudaStream_t s);
size_t loader_counter_bytes();  // LoaderCounters + the update scratch that lives right behind them

// ---- the render kernel: primary rays + shading + shadow rays + RGBA8 pack, one CTA per tile
//      (alternative.cpp:271-383, 702-760, 399-500, 40-83) ----
constexpr int kTileCtaThreads = 160;
constexpr int kMaxLights = 64;
struct TileParams {
    ViewDims d;
    const int* cnt;
    const int* ids;
    const unsigned* occ4;      // 4 bits per bin: cnt & 7 (8 bins per word)
    const int4* boxes;
    const int* atlas_depth;    // [atlas_texels] depth offsets (sprites.hpp:69)
    const float4* texel_tab;   // [atlas_texels] normal.xyz (sprites.hpp:70) + palette colour of the texel (RGBA8 bits in .w)
    const int2* sprite_dims;   // [n_sprites]: texel base, width | height << 16
    int atlas_texels;
    uchar4* out;               // full frame, W*H
    int4* gbuf;                // optional parity checkpoint: entity, y, z, global texel index (-1 = miss)
    int gbuf_only;             // 1: primary rays only (par_get_gbuffer after a frame rendered without gbuf)
    int n_lights;
    float ambient;
    int tile_row_first;        // first owned tile row (bin_y); the next ones are stripe_n apart
    int out_stripe_T;          // 0: raster output; T > 0: stripe-major staging, T stripes per rank
    int n_peer_out;            // fused frame exchange: every finished 16-byte chunk is also stored, in place, into
    uchar4* peer_out[7];       //   the raster frames of the other GPUs (peer memory over NVLink / NVSwitch)
    const int* tile_order;     // optional: CTA -> tile slot (longest-first order of the previous frame's costs)
    unsigned* tile_cost;       // optional: cycles per tile, indexed ty * HW + bx
    int probe_x, probe_y;      // cursor probe pixel (-1: off) and where its 28-byte record goes (mapped host memory)
    int* probe_a;
    int* probe_b;
    int debug_flags;           // bit 0: disable the shaft cull (A/B measurements and tests only)
    int one_light_config;      // 1: launch the 6-CTAs-per-SM build of the kernel (tile_one_light.cu; production frames only)
    unsigned long long* phase_cycles;  // optional debug instrumentation: 16 counters (NULL in production)
    int dbg_light;             // optional export of fp32 intermediates: t.xyz + Lambert term of this light
    float4* dbg_t;             //   [W*H]
    float* dbg_factor;         //   [W*H] acc + ambient
    short4 lights[kMaxLights]; // x, y, z, radius (alternative.cpp:619-622)
};
size_t tile_smem_bytes();
cudaError_t configure_tile();  // per device, before the first launch
int tile_ctas_per_sm();        // resident CTAs of the production kernel per SM (after configure_tile)
cudaError_t launch_tile(const TileParams& p, cudaStream_t s);
// the second build of the same kernel for one-light frames (tile_one_light.cu); launch_tile dispatches to it
size_t tile_one_light_smem_bytes();
cudaError_t configure_tile_one_light();
int tile_one_light_ctas_per_sm();  // 6 on sm_100; 0 = the query failed
cudaError_t launch_tile_one_light(const TileParams& p, int n_tiles, cudaStream_t s);
cudaError_t launch_tile_order(const unsigned* cost, int* order, const ViewDims& d, cudaStream_t s);

}  // namespace par
