// scene_loader.cu — device scene loader: replaces memset + count_entities_in_bins
// (/root/reference/src/alternative.cpp:690-693, 195-269).
//
// The reference inserts entities sequentially into 8-slot rings: slot = count,
// count = (count + 1) & 7 (quirk Q2).  After n inserts a bin reads as its LAST (n mod 8)
// inserts in entity order.  That is a pure function of the SET of inserting entities, so it
// can be built in parallel and deterministically:
//   k_load_cull_insert  one thread per entity: pack the box record, validate, cull
//                       (alternative.cpp:212-219), and for every spanned bin count the insert
//                       (atomicAdd) and push the entity index through the bin's 7 slots with an
//                       atomicMax chain: each slot keeps the larger of (old, new) and hands the
//                       smaller one down, so whatever the interleaving, slot r ends up holding
//                       the (r+1)-th highest inserting entity index.  Survivors are appended to
//                       a compact list.
//   k_occupancy         one thread per survivor: write each spanned bin's wrapped count into a
//                       4-bit-per-bin table (read by the shadow walk: one small load answers
//                       "occupied?" and "how many?").
// Readers map the reference's slot s to ids[bin*8 + (cnt&7) - 1 - s].
#include <algorithm>

#include "par_kernels.cuh"

namespace par {

__global__ void __launch_bounds__(256)
k_load_cull_insert(const int4* __restrict__ raw, const int* __restrict__ sprite_ids, int n,
                   int n_sprites, ViewDims d, int4* __restrict__ boxes, int* __restrict__ cnt,
                   int* ids, int* __restrict__ survivors, LoaderCounters* __restrict__ ctr) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    int4 r = raw[e];
    r.w = sprite_ids ? sprite_ids[e] : 0;
    boxes[e] = r;
    Box b = unpack_box(r);
    if (b.ex < 0 || b.ex > kSpriteW || b.ey < 0 || b.ez < 0 || b.ey + b.ez > 2 * kSpriteW ||
        r.w < 0 || r.w >= n_sprites) {
        ctr->bad_scene = 1;
        return;
    }
    BinRange g;
    if (!cull_and_range(d, b, g)) return;
    survivors[atomicAdd(&ctr->n_survivors, 1)] = e;
    int inserts = 0;
    for (int x = g.x0; x < g.x1; x++)
        for (int y = g.y0; y < g.y1; y++)
            for (int z = g.z0; z < g.z1; z++) {
                const int f = flat_bin(d, x, y, z);
                atomicAdd(&cnt[f], 1);
                int v = e;  // top-7 insertion: slots start at -1
                for (int slot = 0; slot < kSlots - 1 && v >= 0; slot++) {
                    const int old = atomicMax(&ids[f * kSlots + slot], v);
                    v = min(old, v);
                }
                inserts++;
            }
    if (inserts) atomicAdd(&ctr->n_inserts, inserts);
}

__global__ void __launch_bounds__(256)
k_occupancy(const int* __restrict__ survivors, const int4* __restrict__ boxes, ViewDims d,
            const int* __restrict__ cnt, unsigned* __restrict__ occ4,
            const LoaderCounters* __restrict__ ctr) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ctr->n_survivors) return;
    Box b = unpack_box(boxes[survivors[t]]);
    BinRange g;
    cull_and_range(d, b, g);
    for (int x = g.x0; x < g.x1; x++)
        for (int y = g.y0; y < g.y1; y++)
            for (int z = g.z0; z < g.z1; z++) {
                const int f = flat_bin(d, x, y, z);
                const unsigned keep = cnt[f] & (kSlots - 1);
                if (keep) atomicOr(&occ4[f >> 3], keep << ((f & 7) * 4));  // idempotent: same value from every inserter
            }
}

// Clears the grid for a new build in ONE launch: insert totals 0, slots -1, occupancy 0, counters 0.
// (cudaMemsetAsync may be served by a copy engine, where it would queue behind the previous
// frame's readback when frames are pipelined — and four memset nodes cost more than one kernel.)
__global__ void __launch_bounds__(256)
k_clear_grid(int* __restrict__ cnt, int* __restrict__ ids, unsigned* __restrict__ occ4,
             LoaderCounters* __restrict__ ctr, int V) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    int4* ids4 = reinterpret_cast<int4*>(ids);  // kSlots = 8 ints per bin = 2 int4
    for (size_t i = tid; i < (size_t)V * (kSlots / 4); i += nthr) ids4[i] = make_int4(-1, -1, -1, -1);
    for (size_t i = tid; i < (size_t)V; i += nthr) cnt[i] = 0;
    for (size_t i = tid; i < ((size_t)V + 7) / 8; i += nthr) occ4[i] = 0u;
    if (tid == 0) *ctr = LoaderCounters{0, 0, 0, 0};
}

// The counters go to the host through mapped pinned memory, written by the GPU itself: a D2H
// memcpy would queue on the copy engine behind the previous frame's 33 MB readback and stall the
// stream (pipelined frames, par_submit_frame).
__global__ void k_publish_counters(const LoaderCounters* ctr, LoaderCounters* host_a, LoaderCounters* host_b) {
    const LoaderCounters c = *ctr;
    if (host_a) *host_a = c;
    if (host_b) *host_b = c;
}

// Host-side launcher (called from par_api.cu).  host_a / host_b: page-locked host copies of the
// counters (either may be NULL).
cudaError_t launch_scene_loader(const int4* raw, const int* sprite_ids, int n, int n_sprites,
                                const ViewDims& d, int4* boxes, int* cnt, int* ids,
                                unsigned* occ4, int* survivors, LoaderCounters* ctr,
                                LoaderCounters* host_a, LoaderCounters* host_b,
                                cudaStream_t s, int* launches) {
    static_assert(kSlots % 4 == 0, "k_clear_grid writes the slots as int4");
    const int clear_blocks = (int)std::min<size_t>(((size_t)d.V * (kSlots / 4) + 255) / 256, 148 * 8);
    k_clear_grid<<<clear_blocks > 0 ? clear_blocks : 1, 256, 0, s>>>(cnt, ids, occ4, ctr, d.V);
    *launches += 1;
    if (n > 0) {
        int blocks = (n + 255) / 256;
        k_load_cull_insert<<<blocks, 256, 0, s>>>(raw, sprite_ids, n, n_sprites, d, boxes, cnt, ids,
                                                  survivors, ctr);
        // The survivor count lives on the device; size the grid for the worst case (every
        // entity survives) and let surplus threads exit on the device-side count.
        k_occupancy<<<blocks, 256, 0, s>>>(survivors, boxes, d, cnt, occ4, ctr);
        *launches += 2;
    }
    if (host_a || host_b) {
        k_publish_counters<<<1, 1, 0, s>>>(ctr, host_a, host_b);
        *launches += 1;
    }
    return cudaGetLastError();
}

}  // namespace par
