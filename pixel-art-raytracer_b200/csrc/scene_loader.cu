// scene_loader.cu — device scene loader: replaces memset + count_entities_in_bins
// (/root/reference/src/alternative.cpp:690-693, 195-269).
//
// The reference inserts entities sequentially into 8-slot rings: slot = count,
// count = (count + 1) & 7 (quirk Q2).  After n inserts a bin reads as its LAST (n mod 8)
// inserts in entity order.  That is a pure function of the SET of inserting entities, so it
// can be built in parallel and deterministically:
//   1. k_load_cull_count   one thread per entity: pack the box record, cull
//                          (alternative.cpp:212-219), count inserts per bin with atomicAdd,
//                          append the survivors to a compact list;
//   2. k_select_round<r>   r = 0..6: every surviving (entity, bin) pair whose bin keeps more
//                          than r entries proposes itself with atomicMax if it is smaller than
//                          the bin's round r-1 winner — after round r, ids[bin*8+r] is the
//                          (r+1)-th highest inserting entity index.  Round 0 also writes the
//                          bin's wrapped count into a 4-bit-per-bin table (read by the walk:
//                          one small load answers "occupied?" and "how many?").
// Readers map the reference's slot s to ids[bin*8 + (cnt&7) - 1 - s].
#include "par_kernels.cuh"

namespace par {

__global__ void __launch_bounds__(256)
k_load_cull_count(const int4* __restrict__ raw, const int* __restrict__ sprite_ids, int n,
                  int n_sprites, ViewDims d, int4* __restrict__ boxes, int* __restrict__ cnt,
                  int* __restrict__ survivors, LoaderCounters* __restrict__ ctr) {
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
    int inserts = 0, worst = 0;
    for (int x = g.x0; x < g.x1; x++)
        for (int y = g.y0; y < g.y1; y++)
            for (int z = g.z0; z < g.z1; z++) {
                int old = atomicAdd(&cnt[flat_bin(d, x, y, z)], 1);
                worst = max(worst, old + 1);
                inserts++;
            }
    if (inserts) {
        atomicAdd(&ctr->n_inserts, inserts);
        atomicMax(&ctr->max_inserts_per_bin, worst);
    }
}

__global__ void __launch_bounds__(256)
k_select_round(int round, const int* __restrict__ survivors, const int4* __restrict__ boxes,
               ViewDims d, const int* __restrict__ cnt, int* ids, unsigned* __restrict__ occ4,
               const LoaderCounters* __restrict__ ctr) {
    if (round >= ctr->max_inserts_per_bin) return;  // no bin keeps more than `round` entries
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ctr->n_survivors) return;
    int e = survivors[t];
    Box b = unpack_box(boxes[e]);
    BinRange g;
    cull_and_range(d, b, g);
    for (int x = g.x0; x < g.x1; x++)
        for (int y = g.y0; y < g.y1; y++)
            for (int z = g.z0; z < g.z1; z++) {
                int f = flat_bin(d, x, y, z);
                int keep = cnt[f] & (kSlots - 1);
                if (round >= keep) continue;
                if (round == 0) atomicOr(&occ4[f >> 3], (unsigned)keep << ((f & 7) * 4));  // idempotent: same value from every inserter
                int prev = round ? ids[f * kSlots + round - 1] : 0x7fffffff;
                if (e < prev) atomicMax(&ids[f * kSlots + round], e);
            }
}

// Host-side launcher (called from par_api.cu).
cudaError_t launch_scene_loader(const int4* raw, const int* sprite_ids, int n, int n_sprites,
                                const ViewDims& d, int4* boxes, int* cnt, int* ids,
                                unsigned* occ4, int* survivors, LoaderCounters* ctr,
                                cudaStream_t s, int* launches) {
    cudaError_t err;
    if ((err = cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)d.V, s))) return err;
    if ((err = cudaMemsetAsync(ids, 0xff, sizeof(int) * (size_t)d.V * kSlots, s))) return err;
    if ((err = cudaMemsetAsync(occ4, 0, sizeof(unsigned) * (((size_t)d.V + 7) / 8), s))) return err;
    if ((err = cudaMemsetAsync(ctr, 0, sizeof(LoaderCounters), s))) return err;
    if (n > 0) {
        int blocks = (n + 255) / 256;
        k_load_cull_count<<<blocks, 256, 0, s>>>(raw, sprite_ids, n, n_sprites, d, boxes, cnt,
                                                 survivors, ctr);
        // The survivor count lives on the device; size the round grids for the worst case
        // (every entity survives) and let surplus threads exit on the device-side count.
        for (int r = 0; r < kSlots - 1; r++)
            k_select_round<<<blocks, 256, 0, s>>>(r, survivors, boxes, d, cnt, ids, occ4, ctr);
        *launches += kSlots;
    }
    return cudaGetLastError();
}

}  // namespace par
