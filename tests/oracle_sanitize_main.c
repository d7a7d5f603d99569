/* Test driver: the CPU oracle under AddressSanitizer + UndefinedBehaviorSanitizer (the
 * reference itself reads outside its arrays on its own default scene, quirk Q18, which ASan reports at
 * alternative.cpp:476 — the oracle DEFINES those reads, so it must run clean).  Built and run by
 * tests/test_oracle_sanitizers.py; #includes the oracle's translation unit to reach it without a library. */
#include <stdio.h>

#include "../oracle/oracle.c"

static uint64_t frame_hash(const orc_view* v, const orc_aabb* boxes, int n, const orc_light* lights, int n_lights) {
    orc_sprite* sprite = malloc(sizeof *sprite);
    orc_color pal[4];
    orc_make_tile_floor(sprite);
    orc_default_palette(pal);
    size_t px = (size_t)v->W * v->H;
    orc_color* rgba = malloc(px * sizeof *rgba);
    orc_counters ctr;
    memset(&ctr, 0, sizeof ctr);
    if (!sprite || !rgba || orc_render_frame(v, boxes, NULL, n, sprite, pal, lights, n_lights, rgba, NULL, NULL, 0, v->H, &ctr)) {
        fprintf(stderr, "render failed\n");
        exit(2);
    }
    uint64_t h = orc_fnv1a64((const uint8_t*)rgba, px * 4);
    free(rgba);
    free(sprite);
    return h;
}

int main(void) {
    /* 1. the reference's default scene and light at its built-in view: the light's bin lies outside the grid */
    orc_view c1 = {480, 320, 320};
    int n = orc_scene_default(NULL, 0);
    orc_aabb* boxes = malloc(sizeof *boxes * (size_t)n);
    orc_scene_default(boxes, n);
    orc_light light;
    orc_light_default(&light);
    printf("c1 %016llx\n", (unsigned long long)frame_hash(&c1, boxes, n, &light, 1));
    /* 2. lights far outside the grid on every side, entities straddling every face of the view volume */
    orc_view v = {200, 120, 160};
    orc_aabb edge[12];
    const int16_t pos[12][3] = {{-15, 0, 0},  {190, 0, 0},   {0, -15, 10},  {0, 130, 0},  {0, 0, -15},  {0, 0, 150},
                                {-30, 60, 80}, {199, 119, 159}, {100, 60, -39}, {100, -39, 100}, {100, 60, 199}, {50, 50, 50}};
    for (int k = 0; k < 12; k++) edge[k] = (orc_aabb){pos[k][0], pos[k][1], pos[k][2], 20, 20, 20, {0, 0}};
    orc_light far[6] = {{-3000, 50, 50, 10}, {3000, 50, 50, 10}, {50, -3000, 50, 10}, {50, 3000, 50, 10}, {50, 50, -3000, 10}, {32767, 32767, 32767, 10}};
    printf("edges %016llx\n", (unsigned long long)frame_hash(&v, edge, 12, far, 6));
    /* 3. the synthetic recipe at a small view, 16 lights (ring overflow, Q2) */
    orc_view s = {640, 680, 680};
    orc_aabb* syn = malloc(sizeof *syn * 3000);
    orc_light sl[16];
    orc_scene_synthetic(&s, 0xB200, 3000, syn, 16, sl);
    printf("synthetic %016llx\n", (unsigned long long)frame_hash(&s, syn, 3000, sl, 16));
    free(syn);
    free(boxes);
    return 0;
}
