/* oracle.c — CPU restatement of the reference render path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product (libpar_b200.so)
 * never links, loads or calls anything in oracle/.
 *
 * It restates, in plain C and in behaviour (not in text), the per-frame path of
 * Cons-Cat/Pixel-Art-Raytracer:
 *     grid build        /root/reference/src/alternative.cpp:195-269  (orc_grid_build)
 *     primary rays      alternative.cpp:271-383                       (orc_trace_primary)
 *     slab test         alternative.cpp:40-83                         (slab_hit)
 *     shadow grid walk  alternative.cpp:399-500                       (light_visible)
 *     shading loop      alternative.cpp:702-760, sprites.hpp:8-16,28-35 (orc_shade)
 *     debug overlay     alternative.cpp:139-175, 762-772              (orc_draw_overlay)
 *     default scene     alternative.cpp:519-599, 626                  (orc_scene_default)
 *     key map           alternative.cpp:641-681                       (orc_apply_key)
 *     sprite table      sprites.hpp:73-364                            (orc_make_tile_floor)
 * generalised to a runtime view size (W,H,L), a sprite atlas with per-entity ids and N
 * lights (SURVEY.md §8d multi-light rule), and with the reference's out-of-bounds reads
 * (quirk Q18) DEFINED: a bin-count read at a flat index outside [0, volume) yields 0.
 *
 * PARITY PIN: the reference ships no tests or golden vectors.  This oracle is pinned
 * against the reference executable itself, built headless by oracle/build_ref.sh
 * (tests/test_oracle_vs_reference.py, tests/golden/reference_hashes.json).
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp; no FMA contraction — the
 * reference is built for baseline x86-64, which has none).
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_BIN 40       /* single_bin_cubic_size, alternative.cpp:116 */
#define ORC_SLOTS 8      /* sparse_bin_size,       alternative.cpp:131 */
#define ORC_SPR_W 20     /* sprite width hard-coded at alternative.cpp:330 */
#define ORC_SPR_TEXELS 800

typedef struct {
    int16_t px, py, pz, ex, ey, ez;
    int16_t pad[2];
} orc_aabb; /* alternative.cpp:35-38, 16 bytes */

typedef struct {
    int32_t color[ORC_SPR_TEXELS];
    int32_t depth[ORC_SPR_TEXELS];
    float normal[ORC_SPR_TEXELS][3];
} orc_sprite; /* sprites.hpp:67-71, 16000 bytes */

typedef struct {
    uint8_t r, g, b, a;
} orc_color; /* sprites.hpp:5-6 */

typedef struct {
    float nx, ny, nz;
    orc_color color;
    int32_t y, z;
    int32_t entity;
} orc_pixel; /* sprites.hpp:53-58, 28 bytes */

typedef struct {
    int16_t x, y, z, radius;
} orc_light; /* alternative.cpp:619-622 */

typedef struct {
    int32_t W, H, L; /* view_width / view_height / view_length */
} orc_view;

/* Sprite atlas with per-entry dimensions (lifts quirk Q7: the reference hard-codes width 20 at
 * alternative.cpp:330 and 800 texels at sprites.hpp:68-70).  Sprite s is w[s] x h[s] texels,
 * row-major, at off[s] in the three concatenated tables; texel index = row * w[s] + column,
 * which reduces to the reference's row * 20 + column for the 20x40 sprite. */
typedef struct {
    int32_t n;
    const int32_t* w;
    const int32_t* h;
    const int32_t* off;
    const int32_t* color;
    const int32_t* depth;
    const float* normal; /* 3 floats per texel */
} orc_atlas;

/* §8(d) counters, in the order of SURVEY.md's weight table. */
typedef struct {
    uint64_t pixels, primary_bins, primary_slot_tests, primary_passed, primary_accepts;
    uint64_t shaded_px_lights, lit_px_lights, shadow_probes, shadow_slot_entries, slab_tests;
    uint64_t pixels_hit;
} orc_counters;

_Static_assert(sizeof(orc_aabb) == 16, "AABB layout");
_Static_assert(sizeof(orc_sprite) == 16000, "Sprite layout");
_Static_assert(sizeof(orc_pixel) == 28, "Pixel layout");
_Static_assert(sizeof(orc_light) == 8, "Light layout");

static inline int hw(const orc_view* v) { return v->W / ORC_BIN; }
static inline int hh(const orc_view* v) { return v->H / ORC_BIN; }
static inline int hl(const orc_view* v) { return v->L / ORC_BIN; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

int orc_grid_volume(const orc_view* v) { return hw(v) * hh(v) * hl(v); }

/* alternative.cpp:180-182 — z fastest, then screen-row bin, then x. */
static inline int flat_bin(const orc_view* v, int x, int y, int z) {
    return x * hh(v) * hl(v) + y * hl(v) + z;
}

/* std::min / std::max with their exact NaN / argument-order behaviour (quirk Q13). */
static inline float std_minf(float a, float b) { return (b < a) ? b : a; }
static inline float std_maxf(float a, float b) { return (a < b) ? b : a; }

/* ------------------------------------------------------------------ sprite table */

/* sprites.hpp:73-364, stated as rules instead of tables: rows 0-19 are the cube's top
 * face (normal +y, depth 19..0), rows 20-39 its front face (normal -z, depth 0). */
void orc_make_tile_floor(orc_sprite* s) {
    for (int row = 0; row < 40; row++) {
        for (int col = 0; col < ORC_SPR_W; col++) {
            int idx = row * ORC_SPR_W + col;
            int c;
            if (row < 20) {
                c = 0;
                if (row >= 4 && row < 16 && col >= 4 && col < 16) {
                    int left = col < 10, upper = row < 10;
                    c = (left == upper) ? 2 : 3;
                }
            } else if (row < 38) {
                c = (col < 2 || col >= 18) ? 1 : 2;
            } else {
                c = 1;
            }
            s->color[idx] = c;
            s->depth[idx] = row < 20 ? 19 - row : 0;
            s->normal[idx][0] = 0.f;
            s->normal[idx][1] = row < 20 ? 1.f : 0.f;
            s->normal[idx][2] = row < 20 ? 0.f : -1.f;
        }
    }
}

void orc_default_palette(orc_color pal[4]) { /* sprites.hpp:60-65 */
    static const uint8_t g[4] = {100, 140, 200, 240};
    for (int i = 0; i < 4; i++) pal[i] = (orc_color){g[i], g[i], g[i], 0};
}

/* ------------------------------------------------------------------ scenes */

static int push_box(orc_aabb* out, int cap, int n, int x, int y, int z) {
    if (out && n < cap) {
        out[n] = (orc_aabb){(int16_t)x, (int16_t)y, (int16_t)z, 20, 20, 20, {0, 0}};
    }
    return n + 1;
}

/* alternative.cpp:519-599 with the scene constants pinned at 480/320/320 (the view may be
 * larger; the scene is always the default one — SURVEY.md §8c tier 1).  Returns the entity
 * count (162308); writes at most cap records. */
int orc_scene_default(orc_aabb* out, int cap) {
    const int SW = 480, SL = 320;
    int n = 0;
    n = push_box(out, cap, n, SW / 2, 36, SL / 4); /* the "player", entity 0 */
    for (int i = 0; i < SW; i++)                   /* floor with a hole */
        for (int j = 0; j < SL; j++) {
            int x = i * 20, z = j * 20;
            if (x >= SW / 2 - 40 && x < SW / 2 + 40 && z < SL / 2 + 40 && z > SL / 2 - 40)
                continue;
            n = push_box(out, cap, n, x, 0, z);
        }
    for (int i = 0; i < 6; i++) /* left wall block */
        for (int j = 0; j < SL - 10; j++)
            for (int k = 1; k < 6; k++) {
                if (i >= 4 && k >= 4) continue;
                n = push_box(out, cap, n, i * 20, k * 20, SL - j * 20);
            }
    for (int i = 1; i < 3; i++) /* right strip */
        for (int j = 0; j < SL; j++) n = push_box(out, cap, n, SW - i * 20, 20, j * 20);
    for (int i = 1; i < 20; i++) /* back row */
        n = push_box(out, cap, n, SW - 40 - i * 20, 20, SL - 60);
    return n;
}

void orc_light_default(orc_light* l) { /* alternative.cpp:624-626 */
    *l = (orc_light){480, 160, 80, 10};
}

static uint64_t splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* SURVEY.md §8(d) C3/C5 recipe: n cubes then n_lights lights from one splitmix64 stream. */
void orc_scene_synthetic(const orc_view* v, uint64_t seed, int n, orc_aabb* out, int n_lights,
                         orc_light* lights) {
    uint64_t s = seed;
    for (int e = 0; e < n; e++) {
        int x = (int)(splitmix64(&s) % (uint64_t)(v->W - 20));
        int y = (int)(splitmix64(&s) % 200u);
        int z = (int)(splitmix64(&s) % (uint64_t)(v->L - 20));
        out[e] = (orc_aabb){(int16_t)x, (int16_t)y, (int16_t)z, 20, 20, 20, {0, 0}};
    }
    for (int l = 0; l < n_lights; l++) {
        int x = (int)(splitmix64(&s) % (uint64_t)v->W);
        int y = 40 + (int)(splitmix64(&s) % 400u);
        int z = (int)(splitmix64(&s) % (uint64_t)(v->H - 600));
        lights[l] = (orc_light){(int16_t)x, (int16_t)y, (int16_t)z, 10};
    }
}

/* Key semantics of alternative.cpp:641-681.  Keys are named by a letter code:
 * 'L','R' = left/right arrows, 'U','D' = up/down arrows, 'p','P' = page down/up,
 * and the literal light keys a k j u h o. */
void orc_apply_key(int key, orc_aabb* player, orc_light* light) {
    switch (key) {
        case 'L': player->px -= 5; break;
        case 'R': player->px += 5; break;
        case 'U': player->pz += 5; break;
        case 'D': player->pz -= 5; break;
        case 'p': player->py -= 5; break;
        case 'P': player->py += 5; break;
        case 'a': light->z -= 5; break;
        case 'k': light->z += 5; break;
        case 'j': light->y -= 5; break;
        case 'u': light->y += 5; break;
        case 'h': light->x -= 5; break;
        case 'o': light->x += 5; break;
        default: break;
    }
}

/* Script C: the player key delivered before frame f (f >= 1), 0 if none. */
int orc_script_c_key(int f) {
    static const int runs[8][2] = {{30, 'R'}, {20, 'U'}, {50, 'L'}, {30, 'D'},
                                   {30, 'P'}, {40, 'R'}, {30, 'p'}, {9, 'U'}};
    int k = f - 1;
    if (k < 0) return 0;
    for (int r = 0; r < 8; r++) {
        if (k < runs[r][0]) return runs[r][1];
        k -= runs[r][0];
    }
    return 0;
}

/* ------------------------------------------------------------------ grid build */

/* alternative.cpp:195-269 (+ the memset at 690): cull, then push a copy of the box and its
 * entity index into every bin it spans, slot = count, count = (count+1) & 7 (quirk Q2).
 * count: int32[V]; bin_box: orc_aabb[V*8]; bin_ent: int32[V*8].  Entities are visited in
 * index order — the ring makes the result order dependent. */
void orc_grid_build(const orc_view* v, const orc_aabb* boxes, int n, int32_t* count,
                    orc_aabb* bin_box, int32_t* bin_ent) {
    const int W = v->W, H = v->H, L = v->L;
    memset(count, 0, sizeof(int32_t) * (size_t)orc_grid_volume(v));
    for (int e = 0; e < n; e++) {
        const orc_aabb b = boxes[e];
        int x0 = b.px, y0 = b.py, z0 = b.pz;
        int x1 = x0 + b.ex, y1 = y0 + b.ey, z1 = z0 + b.ez;
        /* cull with the reference's hard-coded slack (quirk Q5) */
        if (x1 < 0 || x0 >= W) continue;
        if (y1 < 0 - z1) continue;
        if (y0 >= H - z0 + ORC_BIN) continue;
        if (z1 < -b.ez - ORC_BIN) continue;
        if (z0 > L + ORC_BIN) continue;
        /* bin ranges; the grid's y axis is the screen-row axis (quirk Q4); C integer
         * division truncates toward zero exactly as the reference's does */
        int bx0 = imax(0, x0 / ORC_BIN);
        int by0 = imax(0, (H - y1 - z1) / ORC_BIN);
        int bz0 = imax(0, z0 / ORC_BIN);
        int bx1 = imin(hw(v), (x1 + ORC_BIN - 1) / ORC_BIN);
        int by1 = imin(hh(v), (H - y0 - z0 + ORC_BIN - 1) / ORC_BIN);
        int bz1 = imin(hl(v), (z1 + ORC_BIN - 1) / ORC_BIN);
        for (int bx = bx0; bx < bx1; bx++)
            for (int by = by0; by < by1; by++)
                for (int bz = bz0; bz < bz1; bz++) {
                    int f = flat_bin(v, bx, by, bz);
                    int c = count[f];
                    bin_ent[(size_t)f * ORC_SLOTS + c] = e;
                    bin_box[(size_t)f * ORC_SLOTS + c] = b;
                    count[f] = (c + 1) & (ORC_SLOTS - 1);
                }
    }
}

/* ------------------------------------------------------------------ primary rays */

/* alternative.cpp:271-383 for rows [row0,row1).  gbuf: orc_pixel[W*H] (only the band is
 * written); texel: optional int32[W*H] (sprite texel index of the winning hit, -1 = miss);
 * sprite_ids: optional per-entity atlas index (NULL = all 0). */
static void trace_primary_atlas(const orc_view* v, const int32_t* count, const orc_aabb* bin_box,
                                const int32_t* bin_ent, const orc_atlas* atlas, const int32_t* sprite_ids,
                                const orc_color* palette, orc_pixel* gbuf, int32_t* texel, int row0,
                                int row1, orc_counters* ctr) {
    const int W = v->W, H = v->H, HL = hl(v);
    uint64_t c_bins = 0, c_tests = 0, c_pass = 0, c_acc = 0, c_hit = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : c_bins, c_tests, c_pass, c_acc, c_hit)
    for (int j = row0; j < row1; j++) {
        for (int i = 0; i < W; i++) {
            int world_j = (int16_t)(H - j);
            orc_pixel px = {0.f, 0.f, 0.f, {127, 127, 127, 0}, 0, 0, 0}; /* quirk Q10 */
            int tex = -1;
            int best = INT_MIN;
            int run = 0; /* intersected_bin_count, quirk Q9 */
            int bx = i / ORC_BIN, by = (int16_t)(j / ORC_BIN);
            for (int bz = 0; bz < HL; bz++) {
                c_bins++;
                int f = flat_bin(v, bx, by, bz);
                int cnt = count[f];
                if (cnt == 0) run = 0;
                int any = 0;
                for (int k = 0; k < cnt; k++) {
                    c_tests++;
                    const orc_aabb* b = &bin_box[(size_t)f * ORC_SLOTS + k];
                    int top = b->py + b->ey + b->pz + b->ez;
                    if (!(i >= b->px && i < b->px + b->ex && world_j > b->py + b->pz &&
                          world_j <= top))
                        continue; /* quirk Q6 */
                    c_pass++;
                    int ent = bin_ent[(size_t)f * ORC_SLOTS + k];
                    const int sid = sprite_ids ? sprite_ids[ent] : 0;
                    const int base = atlas->off[sid];
                    int row = top - world_j;
                    int idx = row * atlas->w[sid] + (i - b->px); /* quirk Q7, width per sprite */
                    int d = atlas->depth[base + idx];
                    int key = b->py - b->pz + imin(0, b->ey - row) - d; /* quirk Q8 */
                    if (best >= key) continue;
                    c_acc++;
                    best = key;
                    px.nx = atlas->normal[3 * (size_t)(base + idx) + 0];
                    px.ny = atlas->normal[3 * (size_t)(base + idx) + 1];
                    px.nz = atlas->normal[3 * (size_t)(base + idx) + 2];
                    px.color = palette[atlas->color[base + idx]];
                    px.y = b->py + b->ey + b->ez - row - d; /* quirk Q11 */
                    px.z = b->pz + d;
                    px.entity = ent;
                    tex = idx;
                    any = 1;
                }
                run += any;
                if (run >= 2) break;
            }
            gbuf[(size_t)j * W + i] = px;
            if (texel) texel[(size_t)j * W + i] = tex;
            c_hit += tex >= 0;
        }
    }
    if (ctr) {
        ctr->pixels += (uint64_t)(row1 - row0) * W;
        ctr->primary_bins += c_bins;
        ctr->primary_slot_tests += c_tests;
        ctr->primary_passed += c_pass;
        ctr->primary_accepts += c_acc;
        ctr->pixels_hit += c_hit;
    }
}

/* The reference's fixed 20x40 sprites (sprites.hpp:67-71) as a ragged atlas.  Returns 0 / -1. */
typedef struct {
    orc_atlas a;
    int32_t *w, *h, *off, *color, *depth;
    float* normal;
} atlas_store;

static void atlas_free(atlas_store* st) {
    free(st->w);
    free(st->h);
    free(st->off);
    free(st->color);
    free(st->depth);
    free(st->normal);
    memset(st, 0, sizeof *st);
}

static int atlas_from_sprites(const orc_sprite* sprites, int n, atlas_store* st) {
    memset(st, 0, sizeof *st);
    size_t nt = (size_t)n * ORC_SPR_TEXELS;
    st->w = malloc(sizeof(int32_t) * (size_t)n);
    st->h = malloc(sizeof(int32_t) * (size_t)n);
    st->off = malloc(sizeof(int32_t) * (size_t)n);
    st->color = malloc(sizeof(int32_t) * nt);
    st->depth = malloc(sizeof(int32_t) * nt);
    st->normal = malloc(sizeof(float) * 3 * nt);
    if (!st->w || !st->h || !st->off || !st->color || !st->depth || !st->normal) {
        atlas_free(st);
        return -1;
    }
    for (int s = 0; s < n; s++) {
        st->w[s] = ORC_SPR_W;
        st->h[s] = ORC_SPR_TEXELS / ORC_SPR_W;
        st->off[s] = s * ORC_SPR_TEXELS;
        memcpy(st->color + (size_t)s * ORC_SPR_TEXELS, sprites[s].color, sizeof sprites[s].color);
        memcpy(st->depth + (size_t)s * ORC_SPR_TEXELS, sprites[s].depth, sizeof sprites[s].depth);
        memcpy(st->normal + 3 * (size_t)s * ORC_SPR_TEXELS, sprites[s].normal, sizeof sprites[s].normal);
    }
    st->a = (orc_atlas){n, st->w, st->h, st->off, st->color, st->depth, st->normal};
    return 0;
}

/* Number of atlas entries the scene refers to (max sprite id + 1). */
static int atlas_entries_used(const int32_t* sprite_ids, int n) {
    int m = 0;
    if (sprite_ids)
        for (int e = 0; e < n; e++) m = imax(m, sprite_ids[e]);
    return m + 1;
}

/* ------------------------------------------------------------------ shadow rays */

typedef struct {
    float ix, iy, iz; /* direction_inverse */
    int16_t ox, oy, oz;
} orc_ray; /* alternative.cpp:30-33 */

/* alternative.cpp:40-83: unbounded-line slab test (quirk Q14) with std::min/std::max
 * argument order preserved (quirk Q13). */
static inline int slab_hit(const orc_aabb* b, const orc_ray* r) {
    float x1 = (float)(b->px - r->ox) * r->ix;
    float x2 = (float)(b->px + b->ex - r->ox) * r->ix;
    float tmin = std_minf(x1, x2);
    float tmax = std_maxf(x1, x2);
    float y1 = (float)(b->py - r->oy) * r->iy;
    float y2 = (float)(b->py + b->ey - r->oy) * r->iy;
    tmin = std_maxf(tmin, std_minf(y1, y2));
    tmax = std_minf(tmax, std_maxf(y1, y2));
    float z1 = (float)(b->pz - r->oz) * r->iz;
    float z2 = (float)(b->pz + b->ez - r->oz) * r->iz;
    tmin = std_maxf(tmin, std_minf(z1, z2));
    tmax = std_minf(tmax, std_maxf(z1, z2));
    return tmax >= tmin;
}

/* The reference's occlusion predicate for ONE box and ONE pixel: ray set-up of
 * alternative.cpp:712-722 (L1-normalised direction, quirk Q12; reciprocal) + AABB::intersect.
 * Exported for property tests of exact-output culls (tests/test_shaft_cull_property.py). */
int orc_slab_hit_point(const orc_aabb* box, int ox, int oy, int oz, const orc_light* lt) {
    float tx = (float)(lt->x - ox), ty = (float)(lt->y - oy), tz = (float)(lt->z - oz);
    float len = fabsf(tx) + fabsf(ty) + fabsf(tz);
    tx = tx / len;
    ty = ty / len;
    tz = tz / len;
    orc_ray ray = {1.f / tx, 1.f / ty, 1.f / tz, (int16_t)ox, (int16_t)oy, (int16_t)oz};
    return slab_hit(box, &ray);
}

typedef struct {
    uint64_t probes, entries, slabs;
} walk_ctr;

/* alternative.cpp:399-500: fp32 grid walk from the pixel's bin to the light's bin.  Each of
 * (int)maxabs steps probes the six partial advances (x, y, z, xy, xz, yz) from the last
 * full position and then the full advance xyz, which becomes the new position (quirk Q15).
 * Bins whose flat index equals the start bin's are skipped (Q16); the pixel's own entity is
 * skipped (Q17); out-of-range flat indices read a count of 0 (Q18, defined here). */
static int light_visible(const orc_view* v, const int32_t* count, const orc_aabb* bin_box,
                         const int32_t* bin_ent, int sx, int sy, int sz, int lx, int ly, int lz,
                         int self, const orc_ray* ray, walk_ctr* wc) {
    const int V = orc_grid_volume(v);
    float dx = (float)lx - (float)sx, dy = (float)ly - (float)sy, dz = (float)lz - (float)sz;
    float big = std_maxf(std_maxf(fabsf(dx), fabsf(dy)), fabsf(dz));
    float stx = dx / big, sty = dy / big, stz = dz / big;
    float px = (float)sx, py = (float)sy, pz = (float)sz;
    const int start = flat_bin(v, sx, sy, sz);
    const int steps = (int)big;
    /* advance masks in the reference's order: x, y, z, xy, xz, yz, xyz */
    static const int adv[7] = {1, 2, 4, 3, 5, 6, 7};
    for (int s = 0; s < steps; s++) {
        float nx = px + stx, ny = py + sty, nz = pz + stz;
        for (int p = 0; p < 7; p++) {
            float cx = (adv[p] & 1) ? nx : px;
            float cy = (adv[p] & 2) ? ny : py;
            float cz = (adv[p] & 4) ? nz : pz;
            wc->probes++;
            int f = flat_bin(v, (int)cx, (int)cy, (int)cz);
            if (f == start) continue;
            int cnt = (f >= 0 && f < V) ? count[f] : 0;
            for (int k = 0; k < cnt; k++) {
                wc->entries++;
                if (bin_ent[(size_t)f * ORC_SLOTS + k] == self) continue;
                wc->slabs++;
                if (slab_hit(&bin_box[(size_t)f * ORC_SLOTS + k], ray)) return 0;
            }
        }
        px = nx;
        py = ny;
        pz = nz;
    }
    return 1;
}

/* sprites.hpp:8-16: per-channel float multiply, truncating cast, alpha passthrough. */
static inline orc_color color_scale(orc_color c, float s) {
    orc_color o;
    o.r = (uint8_t)((float)c.r * s);
    o.g = (uint8_t)((float)c.g * s);
    o.b = (uint8_t)((float)c.b * s);
    o.a = c.a;
    return o;
}

/* alternative.cpp:702-760 for rows [row0,row1), N lights:
 *   acc = sum over visible lights of max(0, n . t_l);  out = color * min(1, acc + 0.25)
 * which is bit-identical to the reference for one light. */
static void shade_impl(const orc_view* v, const int32_t* count, const orc_aabb* bin_box,
                       const int32_t* bin_ent, const orc_pixel* gbuf, const orc_light* lights,
                       int n_lights, orc_color* out, int row0, int row1, orc_counters* ctr,
                       int dbg_light, float* dbg_t, float* dbg_factor) {
    const int W = v->W, H = v->H;
    const float ambient = 0.25f;
    uint64_t c_shaded = 0, c_lit = 0, c_probe = 0, c_entry = 0, c_slab = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : c_shaded, c_lit, c_probe, c_entry, c_slab)
    for (int j = row0; j < row1; j++) {
        for (int i = 0; i < W; i++) {
            const orc_pixel* px = &gbuf[(size_t)j * W + i];
            int wx = i, wy = px->y, wz = px->z;
            float acc = 0.f;
            for (int l = 0; l < n_lights; l++) {
                const orc_light* lt = &lights[l];
                float tx = (float)(lt->x - wx), ty = (float)(lt->y - wy),
                      tz = (float)(lt->z - wz);
                float len = fabsf(tx) + fabsf(ty) + fabsf(tz); /* L1 norm, quirk Q12 */
                tx = tx / len;
                ty = ty / len;
                tz = tz / len;
                orc_ray ray = {1.f / tx, 1.f / ty, 1.f / tz, (int16_t)wx, (int16_t)wy,
                               (int16_t)wz};
                int sx = wx / ORC_BIN, sy = (H - wy - wz) / ORC_BIN, sz = wz / ORC_BIN;
                int lx = lt->x / ORC_BIN, ly = (H - lt->y - lt->z) / ORC_BIN,
                    lz = lt->z / ORC_BIN;
                walk_ctr wc = {0, 0, 0};
                c_shaded++;
                int vis = light_visible(v, count, bin_box, bin_ent, sx, sy, sz, lx, ly, lz,
                                        px->entity, &ray, &wc);
                c_probe += wc.probes;
                c_entry += wc.entries;
                c_slab += wc.slabs;
                float dot = px->nx * tx + px->ny * ty + px->nz * tz;
                float lam = std_maxf(0.f, dot);
                if (dbg_t && l == dbg_light) { /* fp32 intermediates of SURVEY.md §8(d) "Tolerance" */
                    float* o = dbg_t + 4 * ((size_t)j * W + i);
                    o[0] = tx;
                    o[1] = ty;
                    o[2] = tz;
                    o[3] = lam;
                }
                if (vis) {
                    c_lit++;
                    acc = acc + lam;
                }
            }
            if (dbg_factor) dbg_factor[(size_t)j * W + i] = acc + ambient;
            out[(size_t)j * W + i] = color_scale(px->color, std_minf(1.f, acc + ambient));
        }
    }
    if (ctr) {
        ctr->shaded_px_lights += c_shaded;
        ctr->lit_px_lights += c_lit;
        ctr->shadow_probes += c_probe;
        ctr->shadow_slot_entries += c_entry;
        ctr->slab_tests += c_slab;
    }
}

void orc_shade(const orc_view* v, const int32_t* count, const orc_aabb* bin_box,
               const int32_t* bin_ent, const orc_pixel* gbuf, const orc_light* lights,
               int n_lights, orc_color* out, int row0, int row1, orc_counters* ctr) {
    shade_impl(v, count, bin_box, bin_ent, gbuf, lights, n_lights, out, row0, row1, ctr, -1, NULL, NULL);
}

/* ------------------------------------------------------------------ debug overlay */

/* alternative.cpp:139-175 + 762-772: red Bresenham line from the cursor pixel's projected
 * surface point to light 0, bounds-checked per pixel.  (cx,cy) is the cursor; the reference
 * starts with the cursor at (0,0). */
void orc_draw_overlay(const orc_view* v, const orc_pixel* gbuf, const orc_light* light, int cx,
                      int cy, orc_color* frame) {
    const int W = v->W, H = v->H;
    const orc_pixel* mp = &gbuf[(size_t)cy * W + cx];
    int x = cx, y = H - (mp->y + mp->z);
    int xe = light->x, ye = H - (light->y + light->z);
    int dx = abs(xe - x), dy = -abs(ye - y);
    int sx = x < xe ? 1 : -1, sy = y < ye ? 1 : -1;
    int err = dx + dy;
    const orc_color red = {255, 0, 0, 255};
    for (;;) {
        if (x >= 0 && y >= 0 && x < W && y < H) frame[(size_t)y * W + x] = red;
        if (x == xe && y == ye) return;
        int e2 = 2 * err;
        if (e2 >= dy) {
            if (x == xe) return;
            err += dy;
            x += sx;
        }
        if (e2 <= dx) {
            if (y == ye) return;
            err += dx;
            y += sy;
        }
    }
}

/* ------------------------------------------------------------------ whole frame */

/* The reference frame loop body alternative.cpp:689-760 (no overlay) for rows [row0,row1), over a
 * ragged atlas.  Scratch for the grid is allocated per call.  dbg_*: optional fp32 intermediates
 * of light dbg_light — t.xyz and the Lambert term (4 floats per pixel) — and acc + ambient (1
 * float per pixel), the operands SURVEY.md §8(d) puts a <= 1 ULP tolerance on.
 * Returns 0, or -1 on allocation failure. */
int orc_render_frame_atlas(const orc_view* v, const orc_aabb* boxes, const int32_t* sprite_ids, int n,
                           const orc_atlas* atlas, const orc_color* palette, const orc_light* lights,
                           int n_lights, orc_color* out_rgba, orc_pixel* out_gbuf, int32_t* out_texel,
                           int row0, int row1, orc_counters* ctr, int dbg_light, float* dbg_t,
                           float* dbg_factor) {
    size_t V = (size_t)orc_grid_volume(v);
    int32_t* count = malloc(sizeof(int32_t) * V);
    orc_aabb* bin_box = malloc(sizeof(orc_aabb) * V * ORC_SLOTS);
    int32_t* bin_ent = malloc(sizeof(int32_t) * V * ORC_SLOTS);
    orc_pixel* gbuf = out_gbuf ? out_gbuf : malloc(sizeof(orc_pixel) * (size_t)v->W * v->H);
    if (!count || !bin_box || !bin_ent || !gbuf) return -1;
    orc_grid_build(v, boxes, n, count, bin_box, bin_ent);
    trace_primary_atlas(v, count, bin_box, bin_ent, atlas, sprite_ids, palette, gbuf, out_texel,
                        row0, row1, ctr);
    if (out_rgba)
        shade_impl(v, count, bin_box, bin_ent, gbuf, lights, n_lights, out_rgba, row0, row1, ctr,
                   dbg_light, dbg_t, dbg_factor);
    if (!out_gbuf) free(gbuf);
    free(count);
    free(bin_box);
    free(bin_ent);
    return 0;
}

/* Same with the reference's fixed 20x40 Sprite records (sprites.hpp:67-71). */
int orc_render_frame(const orc_view* v, const orc_aabb* boxes, const int32_t* sprite_ids, int n,
                     const orc_sprite* atlas, const orc_color* palette, const orc_light* lights,
                     int n_lights, orc_color* out_rgba, orc_pixel* out_gbuf, int32_t* out_texel,
                     int row0, int row1, orc_counters* ctr) {
    atlas_store st;
    if (atlas_from_sprites(atlas, atlas_entries_used(sprite_ids, n), &st)) return -1;
    int rc = orc_render_frame_atlas(v, boxes, sprite_ids, n, &st.a, palette, lights, n_lights, out_rgba,
                                    out_gbuf, out_texel, row0, row1, ctr, -1, NULL, NULL);
    atlas_free(&st);
    return rc;
}

/* alternative.cpp:271-383 with the fixed Sprite records (kept for callers that drive the passes
 * one by one). */
int orc_trace_primary(const orc_view* v, const int32_t* count, const orc_aabb* bin_box,
                      const int32_t* bin_ent, const orc_sprite* atlas, int n_sprites,
                      const int32_t* sprite_ids, const orc_color* palette, orc_pixel* gbuf,
                      int32_t* texel, int row0, int row1, orc_counters* ctr) {
    atlas_store st;
    if (atlas_from_sprites(atlas, n_sprites, &st)) return -1;
    trace_primary_atlas(v, count, bin_box, bin_ent, &st.a, sprite_ids, palette, gbuf, texel, row0, row1, ctr);
    atlas_free(&st);
    return 0;
}

uint64_t orc_fnv1a64(const uint8_t* p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

/* algorithmic lane-ops of a frame = sum(counter x weight). */
double orc_algorithmic_ops(const orc_counters* c) {
    return 2.0 * c->pixels + 12.0 * c->primary_bins + 12.0 * c->primary_slot_tests +
           13.0 * c->primary_passed + 5.0 * c->primary_accepts + 59.0 * c->shaded_px_lights +
           17.0 * c->lit_px_lights + 19.0 * c->shadow_probes + 5.0 * c->shadow_slot_entries +
           32.0 * c->slab_tests;
}
