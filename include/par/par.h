/* par.h — C ABI of the B200-native render path (libpar_b200.so).
 *
 * Drop-in boundary for the per-frame hot path of Cons-Cat/Pixel-Art-Raytracer, i.e. the
 * body of its frame loop, /root/reference/src/alternative.cpp:689-760:
 *
 *     memset(count) + count_entities_in_bins(...)      alternative.cpp:690-693, 195-269
 *     trace_hash_for_pixel(...)                        alternative.cpp:694,     271-383
 *     shading loop + trace_hash_for_light(...)         alternative.cpp:702-760, 399-500, 40-83
 *
 * The reference has no plugin/FFI layer; the seam is that line range.  A host keeps the
 * reference's own scene types (same layouts, checked by static_asserts below and mirrored
 * for C++ in par/reference_types.hpp) and replaces the range with
 *
 *     par_set_scene(ctx, aabbs, sprite_ids, n);              // upload + device grid build
 *     par_render(ctx, lights, n_lights, rgba, gbuf, &stats); // primary + shade + readback
 *
 * Conventions: every function returns PAR_OK (0) or a negative par_status; no exception
 * crosses the boundary; par_last_error() gives a thread-local message.  The caller owns all
 * host buffers and may reuse them as soon as a call returns.  One context serves one host
 * thread at a time and owns its device memory, stream and events.  There is NO CPU
 * fallback: without a CUDA device par_create fails with PAR_ERR_NO_DEVICE.
 */
#ifndef PAR_PAR_H
#define PAR_PAR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAR_BIN_SIZE 40       /* single_bin_cubic_size, alternative.cpp:116 */
#define PAR_BIN_SLOTS 8       /* sparse_bin_size,       alternative.cpp:131 */
#define PAR_SPRITE_W 20       /* sprite width hard-coded at alternative.cpp:330 */
#define PAR_SPRITE_H 40
#define PAR_SPRITE_TEXELS 800 /* sprites.hpp:68-70 */
#define PAR_MAX_SPRITE_DIM 1024   /* par_set_atlas_sized: width and height of a sprite            */
#define PAR_MAX_SPRITE_DEPTH 4095 /* ... and |depth| of a texel (world z of a hit stays 16-bit)   */
#define PAR_MAX_UPDATE 8      /* entities par_update_entities patches in place (more: re-bin all) */
#define PAR_MAX_LIGHTS 64
#define PAR_MAX_VIEW 12800    /* W, H, L upper bound (one grid-walk step per thread) */

typedef enum par_status {
    PAR_OK = 0,
    PAR_ERR_INVALID_ARG = -1,   /* null pointer, bad size, bad band ...              */
    PAR_ERR_NO_DEVICE = -2,     /* no CUDA device / wrong architecture               */
    PAR_ERR_CUDA = -3,          /* a CUDA runtime call failed, see par_last_error()  */
    PAR_ERR_OUT_OF_MEMORY = -4, /* host or device allocation failed                  */
    PAR_ERR_BAD_SCENE = -5,     /* an inserted AABB would index outside its sprite   */
    PAR_ERR_STATE = -6,         /* call order: atlas and scene must be set first     */
    PAR_ERR_NCCL = -7           /* multi-GPU gather failed                           */
} par_status;

/* Reference PODs, byte-identical to the reference's (file:line in comments). */
typedef struct par_aabb { /* AABB, alternative.cpp:35-38 (alignas 16) */
    int16_t px, py, pz; /* position */
    int16_t ex, ey, ez; /* extent   */
    int16_t pad[2];
} par_aabb;

typedef struct par_color { /* Color, sprites.hpp:5-6 */
    uint8_t r, g, b, a;
} par_color;

typedef struct par_sprite { /* Sprite, sprites.hpp:67-71 */
    int32_t color[PAR_SPRITE_TEXELS]; /* palette index */
    int32_t depth[PAR_SPRITE_TEXELS];
    float normal[PAR_SPRITE_TEXELS][3];
} par_sprite;

typedef struct par_pixel { /* Pixel, sprites.hpp:53-58 — the G-buffer record */
    float nx, ny, nz;
    par_color color;
    int32_t y, z;
    int32_t entity;
} par_pixel;

typedef struct par_light { /* Light, alternative.cpp:619-622 */
    int16_t x, y, z;
    int16_t radius; /* never read by the reference */
} par_light;

/* Replaces the compile-time constants of alternative.cpp:116-131. */
typedef struct par_config {
    int32_t width;     /* view_width  (multiple of 40, <= PAR_MAX_VIEW) */
    int32_t height;    /* view_height (multiple of 40)                  */
    int32_t length;    /* view_length (multiple of 40)                  */
    int32_t device;    /* CUDA device ordinal                           */
    int32_t row_begin; /* this context renders rows [row_begin,row_end); */
    int32_t row_end;   /* both 0 = the whole frame (row-band multi-GPU)  */
    float ambient;     /* ambient_light, alternative.cpp:702; 0 selects 0.25f */
    int32_t stripe_count; /* > 1: of the band, render only the 40-row tile rows t with          */
    int32_t stripe_index; /* t % stripe_count == stripe_index (interleaved stripes balance the   */
                          /* per-row cost over GPUs far better than contiguous bands); 0/1 = all */
    int32_t tile_order;   /* longest-tile-first CTA order from the previous frame's per-tile cost: */
                          /* 0 = automatic (>= 2 lights, or one light when the tiles make 1-3 waves */
                          /* of resident CTAs), 1 = always, -1 = never                              */
    int32_t stripe_split; /* > 1: every 40-row tile row is cut into this many stripes of equal width  */
                          /* (must divide width / 40) and stripe v = tile row * split + k goes to rank   */
                          /* v % stripe_count: equal stripe counts per rank when height / 40 is not a    */
                          /* multiple of the rank count (e.g. 108 tile rows over 8 GPUs: split 2).  Its  */
                          /* columns are segment (k + v / lcm(stripe_count, split)) % split of the row,  */
                          /* so a rank's stripes take every column segment in turn.                      */
                          /* 0 / 1 = whole tile rows.  Not with the stripe-major staging calls.          */
    int32_t reserved;
} par_config;

/* Filled by par_render / par_get_stats; GPU times are CUDA-event milliseconds on the
 * context's stream for the most recent build / frame. */
typedef struct par_stats {
    float ms_grid_build;   /* scene loader kernels (cull + bin + select)        */
    float ms_render;       /* the render kernel: primary rays + shading + shadow rays + RGBA8 pack */
    float ms_reserved;     /* (0)                                               */
    float ms_total;        /* first kernel start to last kernel end of the frame */
    int32_t kernel_launches; /* kernels launched by the most recent build + frame */
    int32_t n_entities;
    int32_t n_survivors;   /* entities that passed the cull (alternative.cpp:212-219) */
    int32_t n_inserts;     /* (entity, bin) insertions (alternative.cpp:243-267)       */
    uint64_t rays;         /* reference-equivalent rays: rows*W*(1+n_lights)           */
    uint64_t slab_tests;   /* reserved (0)                                              */
    float ms_reserved2;    /* (0)                                                       */
    float ms_readback;     /* par_wait_frame only: kernels done -> frame complete on the host
                            * (queueing behind the previous frame's copy + the D2H itself);
                            * ms_total is submit -> complete, the per-kernel times stay 0 */
    int32_t reserved[2];
} par_stats;

typedef struct par_ctx par_ctx;

/* -- lifetime ------------------------------------------------------------------------- */
int par_create(par_ctx** out, const par_config* cfg);
void par_destroy(par_ctx* ctx);
const char* par_last_error(void);
const char* par_version(void);

/* Use an external CUDA stream (cudaStream_t cast to void*) for all work of this context;
 * NULL restores the context's own stream.  Lets a host framework order and time the work. */
int par_set_stream(par_ctx* ctx, void* cuda_stream);
void* par_get_stream(par_ctx* ctx); /* the cudaStream_t all work of the context is ordered on */
int par_sync(par_ctx* ctx);

/* Pinned host memory for frame/scene buffers (optional; any host pointer is accepted). */
void* par_alloc_host(size_t bytes);
void par_free_host(void* p);

/* -- scene ---------------------------------------------------------------------------- */
/* Sprite atlas + palette (replaces Entities::sprites, alternative.cpp:95, and color_palette,
 * sprites.hpp:60-65).  Copied; call once, or again whenever sprites change. */
int par_set_atlas(par_ctx* ctx, const par_sprite* sprites, int n_sprites,
                  const par_color* palette, int n_palette);

/* Same with per-sprite dimensions (lifts quirk Q7: the reference hard-codes a 20-texel row at
 * alternative.cpp:330 and 800 texels per sprite at sprites.hpp:68-70).  Sprite s is widths[s] x
 * heights[s] texels (1..PAR_MAX_SPRITE_DIM each), row-major; the three tables hold the sprites
 * back to back (color: palette index, depth: |d| <= PAR_MAX_SPRITE_DEPTH, normal: 3 floats per
 * texel).  A texel index is row * widths[s] + column; an entity of sprite s must have
 * extent.x <= widths[s] and extent.y + extent.z <= heights[s] (else PAR_ERR_BAD_SCENE). */
int par_set_atlas_sized(par_ctx* ctx, int n_sprites, const int32_t* widths, const int32_t* heights,
                        const int32_t* color, const int32_t* depth, const float* normal,
                        const par_color* palette, int n_palette);

/* Per-frame scene: replaces Entities::aabbs (alternative.cpp:94) and runs the device scene
 * loader = memset + count_entities_in_bins (alternative.cpp:690-693).  sprite_ids may be
 * NULL (every entity uses atlas entry 0, which is what Entities::insert produces,
 * alternative.cpp:105-108).  Host data is copied before return of the NEXT synchronising
 * call at the latest when it is pinned; pageable buffers are copied before return. */
int par_set_scene(par_ctx* ctx, const par_aabb* aabbs, const int32_t* sprite_ids, int n);

/* Re-run the device scene loader on the scene already resident in HBM (no upload). */
int par_rebuild_grid(par_ctx* ctx);

/* Incremental scene update: entities [first, first + count) of the RESIDENT scene get new boxes
 * (and, with sprite_ids != NULL, new sprites).  The reference's input handling moves one entity
 * per key (alternative.cpp:641-660) yet re-bins all of them every frame (689-693); here up to
 * PAR_MAX_UPDATE entities are patched in place — their 16-byte records travel as kernel
 * arguments, and only the bins their old and new boxes span are rebuilt — so a moving player
 * costs 16 bytes of upload instead of the whole scene.  Larger updates upload the range and
 * re-bin everything on the device.  The resulting grid is identical to a full par_set_scene. */
int par_update_entities(par_ctx* ctx, int first, int count, const par_aabb* aabbs,
                        const int32_t* sprite_ids);

/* -- frame ---------------------------------------------------------------------------- */
/* Render one frame and read it back.  out_rgba: host, W*H par_color, caller-owned; rows of
 * the context's band are written at their place in the full frame (others untouched).
 * out_gbuf: optional host W*H par_pixel (band rows written).  Synchronous at the API. */
int par_render(par_ctx* ctx, const par_light* lights, int n_lights, par_color* out_rgba,
               par_pixel* out_gbuf, par_stats* stats);

/* Render into DEVICE memory, asynchronously on the context's stream: d_rgba points at a
 * full W*H*4-byte frame in HBM on the context's device; band rows are written in place
 * (so an in-place all-gather over bands completes the frame).  NULL renders into the
 * context's own frame buffer.  Use par_sync / stream ordering before consuming it. */
int par_render_device(par_ctx* ctx, const par_light* lights, int n_lights, void* d_rgba);

/* Striped multi-GPU rendering: like par_render_device, but the context's stripes are written
 * STRIPE-MAJOR into d_staging (par_staging_bytes() bytes, laid out [stripe_count][T][40 rows][W]
 * with T = ceil(H/40 / stripe_count)), i.e. every rank's output is one contiguous block at
 * offset stripe_index * T * 40 * W * 4 — exactly what an in-place all-gather wants.
 * par_unstripe_device then turns a gathered staging frame into the raster W*H frame. */
int par_render_device_striped(par_ctx* ctx, const par_light* lights, int n_lights, void* d_staging);
size_t par_staging_bytes(const par_ctx* ctx);
int par_unstripe_device(par_ctx* ctx, const void* d_staging, void* d_rgba);

/* Fused frame exchange (striped contexts): k_shade stores every finished 16-byte chunk of its
 * stripes into its own raster frame AND, in place, into the raster frames of all other ranks
 * through peer memory (NVLink / NVSwitch), so the frame is complete on every GPU as soon as all
 * ranks' kernels have finished — a barrier replaces the all-gather + un-stripe passes.
 *   one process per GPU : par_peer_export (64-byte CUDA IPC handle of the own frame), exchange the
 *                         handles, par_peer_import(rank, handle) for every other rank;
 *   one process, N GPUs : par_peer_set(rank, par_device_frame(other ctx)) after enabling peer access.
 * par_render_device_peers stores into EVERY frame imported so far: import all other ranks for an
 * all-gather (the frame complete on every GPU), or import only the root's frame on the other
 * ranks (the root imports nothing) for a gather-to-root — 1/(N-1) of the NVLink traffic, the
 * right choice when one GPU feeds the display / encoder.  It renders asynchronously; the caller
 * provides the cross-rank barrier. */
int par_peer_export(par_ctx* ctx, void* handle64);
int par_peer_import(par_ctx* ctx, int rank, const void* handle64);
int par_peer_set(par_ctx* ctx, int rank, void* d_peer_frame);
int par_render_device_peers(par_ctx* ctx, const par_light* lights, int n_lights);
/* Enqueue a D2H copy of the context's whole raster frame on its stream (asynchronous for pinned
 * memory; par_sync or stream ordering before reading the host buffer). */
int par_read_frame(par_ctx* ctx, par_color* out_rgba);
/* Enqueue a D2H copy of only the rows this context OWNS (its band / interleaved stripes) into
 * their raster position inside a full W*H host frame — one strided DMA.  With one context per
 * GPU all pointing at the same pinned host frame (one process: par_alloc_host; several
 * processes: shared memory + par_register_host in each), the N GPUs move the frame over their
 * N PCIe links in parallel and no GPU-to-GPU exchange is needed for a host-side consumer
 * (the SDL texture upload of alternative.cpp:786-788 reads host memory). */
int par_read_stripes(par_ctx* ctx, par_color* host_frame);
/* Pipelined frames — the throughput form of par_set_scene + par_render for a host that produces
 * a scene and consumes a frame every iteration (the reference's loop, alternative.cpp:641-788).
 * par_submit_frame enqueues upload, grid build, both kernels and the readback of the rows this
 * context owns into out_rgba (a full W*H frame) and returns; at most two frames are in flight.
 * par_wait_frame blocks until the OLDEST frame in flight is complete in its out_rgba (status
 * and stats are that frame's; ms_total = submit to completion).  The readback of frame k runs
 * on a second stream beside the upload and kernels of frame k+1 (PCIe is full duplex), so a
 * frame costs max(readback, upload + kernels) instead of their sum.  aabbs / sprite_ids /
 * out_rgba should be page-locked (par_alloc_host) and must stay untouched until that frame has
 * been waited for — a moving scene alternates two AABB arrays the way it alternates two frames. */
int par_submit_frame(par_ctx* ctx, const par_aabb* aabbs, const int32_t* sprite_ids, int n,
                     const par_light* lights, int n_lights, par_color* out_rgba);
int par_wait_frame(par_ctx* ctx, par_stats* stats);
/* par_submit_frame for a scene that is already resident: applies par_update_entities(first, count,
 * aabbs, sprite_ids) (count may be 0) and renders — no scene upload at all.  aabbs may be any host
 * memory and may be reused as soon as the call returns. */
int par_submit_update(par_ctx* ctx, int first, int count, const par_aabb* aabbs, const int32_t* sprite_ids,
                      const par_light* lights, int n_lights, par_color* out_rgba);
/* Row pitch, in bytes, of every HOST frame this context writes (par_render, par_submit_*,
 * par_read_frame, par_read_stripes): the blit contract of alternative.cpp:774-788, where the
 * destination is a locked texture with its own pitch.  0 (default) = packed rows of W * 4 bytes. */
int par_set_output_pitch(par_ctx* ctx, size_t pitch_bytes);
/* D2H of the whole raster frame into rows `pitch_bytes` apart (one strided DMA, asynchronous for
 * pinned memory) — SDL_LockTexture + the row loop + SDL_UnlockTexture of alternative.cpp:774-783. */
int par_read_frame_pitched(par_ctx* ctx, void* dst, size_t pitch_bytes);

/* Render a frame from the RESIDENT scene: device scene loader + render kernel, asynchronously on
 * the context's stream, into the context's own frame (par_device_frame).  The current grid generation
 * is always the grid of the resident scene (every scene call rebuilds or patches it at once), so the
 * frame renders from it while the loader — the reference re-bins every frame, alternative.cpp:689-693 —
 * rebuilds the OTHER generation from the same scene on a side branch, for the next frame.  The launch
 * sequence is captured once into CUDA graphs and replayed while the lights stay the same.  With an exchange
 * set up (par_exchange_setup) the call also carries the multi-GPU frame exchange: the render
 * kernel's stores go to the consumers' frames as well, arrival is signalled through flags in the
 * consumers' frame footers, consumers wait for all producers on their stream, and a producer does
 * not overwrite a consumer's frame before that consumer has started its next call — no NCCL
 * collective, no host round trip. */
int par_render_resident(par_ctx* ctx, const par_light* lights, int n_lights);
/* Multi-GPU exchange of par_render_resident for striped contexts: every other rank must have been
 * imported (par_peer_import / par_peer_set).  root >= 0: gather-to-root (only that rank's frame
 * is completed); root = -1: all-gather (every rank's frame is completed). */
int par_exchange_setup(par_ctx* ctx, int root);
/* The reference keeps a pointer to the G-buffer record under the mouse cursor (mouse_pixel,
 * alternative.cpp:380-382) for its debug overlay (762-772).  par_set_cursor selects that pixel
 * (it must be one this context renders; x < 0 switches the probe off); from then on every
 * frame also delivers its 28-byte record — no G-buffer readback needed for the overlay.
 * par_cursor_pixel returns the record of the frame most recently completed by par_render /
 * par_wait_frame (after par_render_device it waits for the stream first). */
int par_set_cursor(par_ctx* ctx, int x, int y);
int par_cursor_pixel(par_ctx* ctx, par_pixel* out);
/* Page-lock / unlock host memory the caller already owns (e.g. a shared-memory frame), so that
 * copies into it are asynchronous DMA. */
int par_register_host(void* p, size_t bytes);
int par_unregister_host(void* p);

/* Device pointer of the context's own W*H*4 frame buffer. */
void* par_device_frame(par_ctx* ctx);

/* Parity checkpoints of the most recent frame / build (synchronous copies to host):
 *   gbuf   W*H par_pixel, texel W*H int32 (sprite texel index of the hit, -1 = miss)
 *   count  int32[volume]  (p_aabb_count_in_bin), ids int32[volume*8] (entity-index map in
 *          slot order; slots >= count[bin] are -1).  Any pointer may be NULL. */
int par_get_gbuffer(par_ctx* ctx, par_pixel* gbuf, int32_t* texel);
/* fp32 intermediates of the most recent frame, recomputed by the render kernel itself with the
 * export switched on (SURVEY.md 8d "Tolerance"): t_lam = W*H x 4 floats, the L1-normalised
 * direction towards light `light` (alternative.cpp:711-715) and its Lambert term (745-747);
 * factor = W*H floats, acc + ambient (the operand of the final min, alternative.cpp:757-758).
 * Only hit pixels are written (the kernel never evaluates t for a miss pixel, quirk Q19); the
 * rest stays 0.  lights must be the ones the frame was rendered with.  Either may be NULL. */
int par_debug_intermediates(par_ctx* ctx, const par_light* lights, int n_lights, int light,
                            float* t_lam, float* factor);
int par_get_grid(par_ctx* ctx, int32_t* count, int32_t* ids);
int par_get_stats(par_ctx* ctx, par_stats* stats);
int par_grid_volume(const par_ctx* ctx);
/* Debug aid: barrier-to-barrier cycle totals of the shade kernel's phases (16 counters,
 * summed over CTAs).  enable != 0 switches the instrumentation on and zeroes it. */
int par_debug_phase_timing(par_ctx* ctx, int enable, uint64_t* out16);

/* -- single-process multi-GPU: interleaved stripes ------------------------------------------- */
/* One par_ctx per device renders the tile rows t with t % n == i of the same scene (the scene
 * and grid are replicated).  Host consumer (out_rgba != NULL): every device DMAs its own stripes
 * into out_rgba over its own PCIe link (par_read_stripes; pass par_alloc_host memory so the N
 * copies run concurrently) — no GPU-to-GPU traffic.  Device consumer (out_rgba == NULL): the
 * frame is completed on every device, by peer-memory stores fused into the shade kernel
 * (par_peer_set) or, without peer access / with PAR_MULTI_EXCHANGE=nccl, by stripe-major
 * staging + one in-place ncclAllGather + an un-stripe copy.  NCCL is loaded lazily with dlopen,
 * only when that fallback is taken. */
typedef struct par_multi par_multi;
int par_multi_create(par_multi** out, const par_config* cfg, const int* devices, int n_devices);
void par_multi_destroy(par_multi* m);
int par_multi_size(const par_multi* m);
par_ctx* par_multi_context(par_multi* m, int i); /* device i's context (G-buffer, stats, grid) */
int par_multi_set_atlas(par_multi* m, const par_sprite* sprites, int n_sprites,
                        const par_color* palette, int n_palette);
int par_multi_set_scene(par_multi* m, const par_aabb* aabbs, const int32_t* sprite_ids, int n);
/* Renders every stripe set into out_rgba (host, W*H), or with out_rgba == NULL leaves the frame
 * in HBM, complete on every device (par_device_frame(par_multi_context(m, i))).  Synchronous. */
int par_multi_render(par_multi* m, const par_light* lights, int n_lights, par_color* out_rgba,
                     par_stats* stats);
const char* par_multi_last_error(void);

/* -- host-side pieces of the reference that sit either side of the path ---------------- */
/* make_tile_floor(), sprites.hpp:73-364, and color_palette, sprites.hpp:60-65. */
void par_sprite_tile_floor(par_sprite* out);
void par_palette_default(par_color out[4]);
/* Default scene of alternative.cpp:519-599 (scene constants 480/320/320) and its light
 * (alternative.cpp:624-626).  Returns the entity count; writes at most cap records. */
int par_scene_default(par_aabb* out, int cap);
void par_light_default(par_light* out);
/* SURVEY.md §8(d) synthetic recipe (C3/C5): n cubes + n_lights lights from splitmix64.  Stated for
 * width, length > 20 and height > 600; smaller views clamp the affected coordinate range to 1. */
void par_scene_synthetic(int width, int height, int length, uint64_t seed, int n,
                         par_aabb* out_aabbs, int n_lights, par_light* out_lights);
/* Key semantics of alternative.cpp:641-681 on entity 0 / light 0.  key: 'L','R' arrows,
 * 'U','D' arrows, 'P'/'p' page up/down, and the literal light keys a k j u h o. */
void par_apply_key(int key, par_aabb* player, par_light* light);
/* Debug overlay of alternative.cpp:139-175, 762-772 drawn into a host frame: a red line from the
 * surface point under the cursor to light 0.  par_draw_overlay reads the record from a whole
 * G-buffer, par_draw_overlay_at takes the single record (par_cursor_pixel). */
void par_draw_overlay(int width, int height, const par_pixel* gbuf, const par_light* light,
                      int cursor_x, int cursor_y, par_color* frame);
void par_draw_overlay_at(int width, int height, const par_pixel* under_cursor, const par_light* light,
                         int cursor_x, par_color* frame);
/* FNV-1a-64 over a byte range — the per-frame hash of the 240-frame golden sequences,
 * taken over the RGBA bytes handed to the frame sink (alternative.cpp:774-788). */
uint64_t par_fnv1a64(const void* data, size_t bytes);

#ifdef __cplusplus
} /* extern "C" */
#if __cplusplus >= 201103L
static_assert(sizeof(par_aabb) == 16, "AABB must stay 16 bytes (alternative.cpp:88)");
static_assert(sizeof(par_sprite) == 16000, "Sprite layout (sprites.hpp:67-71)");
static_assert(sizeof(par_pixel) == 28, "Pixel layout (sprites.hpp:53-58)");
static_assert(sizeof(par_light) == 8, "Light layout (alternative.cpp:619-622)");
static_assert(sizeof(par_color) == 4, "Color layout (sprites.hpp:5-6)");
#endif
#endif
#endif /* PAR_PAR_H */
