// scene_loader.cu — device scene loader: replaces memset + count_entities_in_bins
// (/root/reference/src/alternative.cpp:690-693, 195-269), plus the incremental update the
// reference's input handling calls for (alternative.cpp:641-681 moves entity 0 only).
//
// The reference inserts entities sequentially into 8-slot rings: slot = count,
// count = (count + 1) & 7 (quirk Q2).  After n inserts a bin reads as its LAST (n mod 8)
// inserts in entity order.  That is a pure function of the SET of inserting entities, so it
// can be built in parallel and deterministically.  A full build is TWO launches:
//   k_load_insert   (a) one thread per entity: pack the box record, cull (alternative.cpp:212-219),
//                   validate the survivors that span a bin against their sprite's size, and for every
//                   spanned bin count the insert (atomicAdd) and push the entity index through the
//                   bin's 7 slots with an atomicMax chain: each slot keeps the larger of (old, new) and
//                   hands the smaller one down, so whatever the interleaving, slot r ends up holding
//                   the (r+1)-th highest inserting entity index.  Survivors are appended to a list.
//                   (b) the same threads clear the OTHER grid generation — only the bins its own
//                   survivor list touched (C2: 32 k of 280 k bins, C5: 42 k of 2.2 M), instead of a
//                   memset of the whole grid.  Two generations alternate, so the clear of the grid
//                   frame k-1 used never sits between frame k's kernels.
//   (k_clear_touched + k_reset_counters clear a generation on its own, for the pipelined resident frame.)
//   k_occupancy     one thread per survivor: write each spanned bin's wrapped count into a 4-bit-
//                   per-bin table (read by the shadow walk: one small load answers "occupied?" and
//                   "how many?"); the last block to finish publishes the counters straight into
//                   mapped pinned host memory (a D2H memcpy would queue on the copy engine behind the
//                   previous frame's readback when frames are pipelined).
// Readers map the reference's slot s to ids[bin*8 + (cnt&7) - 1 - s].
//
// Incremental update (k_update_*): the bins spanned by the old and new boxes of the moved entities
// ("dirty" bins) are cleared and rebuilt from ALL entities — a bin's content depends on the whole
// set of its inserters (the ring keeps the last n mod 8), so it cannot be patched from the slots
// alone — while every other bin is left as it is.
#include <algorithm>
#include <climits>

#include "par_kernels.cuh"

namespace par {

namespace {

// Does the box index inside its sprite (quirk Q7: texel = row * width + column)?
__device__ __forceinline__ bool fits_sprite(const Box& b, const int2* __restrict__ sprite_dims, int n_sprites) {
    if (b.sprite < 0 || b.sprite >= n_sprites) return false;
    const int wh = sprite_dims[b.sprite].y;
    return b.ex >= 0 && b.ey >= 0 && b.ez >= 0 && b.ex <= (wh & 0xffff) && b.ey + b.ez <= (wh >> 16);
}

__device__ __forceinline__ int range_volume(const BinRange& g) {
    return max(g.x1 - g.x0, 0) * max(g.y1 - g.y0, 0) * max(g.z1 - g.z0, 0);
}

__device__ __forceinline__ void insert_into_bin(const GridBuffers& g, int f, int e) {
    atomicAdd(&g.cnt[f], 1);
    int v = e;  // top-7 insertion: slots start at -1
    for (int slot = 0; slot < kSlots - 1 && v >= 0; slot++) {
        const int old = atomicMax(&g.ids[(size_t)f * kSlots + slot], v);
        v = min(old, v);
    }
}

__device__ __forceinline__ void clear_bin(const GridBuffers& g, int f) {
    g.cnt[f] = 0;
    int4* slots = reinterpret_cast<int4*>(g.ids + (size_t)f * kSlots);
    slots[0] = make_int4(-1, -1, -1, -1);
    slots[1] = make_int4(-1, -1, -1, -1);
}

}  // namespace

// Clear the bins listed survivor t of generation g touched in its last build.
__device__ __forceinline__ void clear_touched(const ViewDims& d, const GridBuffers& g, int t) {
    const Box b = unpack_box(g.boxes[g.survivors[t]]);
    BinRange r;
    if (cull_and_range(d, b, r))
        for (int x = r.x0; x < r.x1; x++)
            for (int y = r.y0; y < r.y1; y++)
                for (int z = r.z0; z < r.z1; z++) {
                    const int f = flat_bin(d, x, y, z);
                    clear_bin(g, f);
                    g.occ4[f >> 3] = 0u;  // every non-zero nibble belongs to a touched bin: all end up 0
                }
}

__global__ void __launch_bounds__(256)
k_load_insert(const __grid_constant__ LoaderParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const ViewDims& d = p.d;
    // (b) clear what the other generation's last build touched
    if (p.old.cnt && t < p.old.ctr->n_list) clear_touched(d, p.old, t);
    // (a) this frame's entity
    if (t >= p.n) return;
    int4 r = p.raw[t];
    r.w = p.sprite_ids ? p.sprite_ids[t] : 0;
    p.cur.boxes[t] = r;
    const Box b = unpack_box(r);
    BinRange g;
    if (!cull_and_range(d, b, g)) return;
    p.cur.survivors[atomicAdd(&p.cur.ctr->n_survivors, 1)] = t;
    const int inserts = range_volume(g);
    if (inserts == 0) return;
    // Only a box that is inserted can ever be indexed by a primary ray (alternative.cpp:324-332): a
    // culled entity may carry any extents or sprite id, exactly as in the reference.
    if (!fits_sprite(b, p.sprite_dims, p.n_sprites)) {
        p.cur.ctr->bad_scene = 1;
        atomicMin(&p.cur.ctr->bad_entity, t);
        return;
    }
    for (int x = g.x0; x < g.x1; x++)
        for (int y = g.y0; y < g.y1; y++)
            for (int z = g.z0; z < g.z1; z++) insert_into_bin(p.cur, flat_bin(d, x, y, z), t);
    atomicAdd(&p.cur.ctr->n_inserts, inserts);
}

__device__ __forceinline__ void publish(LoaderCounters* ctr, LoaderCounters* host_a, LoaderCounters* host_b) {
    LoaderCounters c = *ctr;
    c.blocks_done = 0;
    if (host_a) *host_a = c;
    if (host_b) *host_b = c;
}

__global__ void __launch_bounds__(256)
k_occupancy(const __grid_constant__ LoaderParams p) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const ViewDims& d = p.d;
    LoaderCounters* ctr = p.cur.ctr;
    const int n_surv = ctr->n_survivors;
    if (t < n_surv) {
        const Box b = unpack_box(p.cur.boxes[p.cur.survivors[t]]);
        BinRange g;
        if (cull_and_range(d, b, g) && fits_sprite(b, p.sprite_dims, p.n_sprites))
            for (int x = g.x0; x < g.x1; x++)
                for (int y = g.y0; y < g.y1; y++)
                    for (int z = g.z0; z < g.z1; z++) {
                        const int f = flat_bin(d, x, y, z);
                        const unsigned keep = p.cur.cnt[f] & (kSlots - 1);
                        if (keep) atomicOr(&p.cur.occ4[f >> 3], keep << ((f & 7) * 4));  // idempotent: same value from every inserter
                    }
    }
    // the last block to get here publishes the counters and re-arms the other generation's
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(&ctr->blocks_done, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        ctr->n_list = n_surv;
        ctr->blocks_done = 0;
        publish(ctr, p.host_a, p.host_b);
        if (p.old.ctr) *p.old.ctr = LoaderCounters{0, 0, INT_MAX, 0, 0, 0, {0, 0}};
    }
}

__global__ void __launch_bounds__(256)
k_clear_grid(GridBuffers g, int V) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    int4* ids4 = reinterpret_cast<int4*>(g.ids);  // kSlots = 8 ints per bin = 2 int4
    for (size_t i = tid; i < (size_t)V * (kSlots / 4); i += nthr) ids4[i] = make_int4(-1, -1, -1, -1);
    for (size_t i = tid; i < (size_t)V; i += nthr) g.cnt[i] = 0;
    for (size_t i = tid; i < ((size_t)V + 7) / 8; i += nthr) g.occ4[i] = 0u;
    if (tid == 0) *g.ctr = LoaderCounters{0, 0, INT_MAX, 0, 0, 0, {0, 0}};
}

// A generation cleared on its own (the pipelined resident frame rebuilds the generation the previous frame
// rendered from while the current frame renders from the other one): the touched bins first, then — a second
// launch, because the list length lives in the counters — the counters.
__global__ void __launch_bounds__(256)
k_clear_touched(ViewDims d, GridBuffers g) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < g.ctr->n_list) clear_touched(d, g, t);
}

__global__ void k_reset_counters(LoaderCounters* ctr) { *ctr = LoaderCounters{0, 0, INT_MAX, 0, 0, 0, {0, 0}}; }

cudaError_t launch_clear_touched(const ViewDims& d, const GridBuffers& g, int n_list_cap, cudaStream_t s, int* launches) {
    k_clear_touched<<<std::max(1, (n_list_cap + 255) / 256), 256, 0, s>>>(d, g);
    k_reset_counters<<<1, 1, 0, s>>>(g.ctr);
    *launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_clear_grid(const GridBuffers& g, int V, cudaStream_t s) {
    static_assert(kSlots % 4 == 0, "the slots of a bin are cleared as int4");
    const int blocks = (int)std::min<size_t>(((size_t)V * (kSlots / 4) + 255) / 256, 148 * 8);
    k_clear_grid<<<blocks > 0 ? blocks : 1, 256, 0, s>>>(g, V);
    return cudaGetLastError();
}

cudaError_t launch_scene_loader(const LoaderParams& p, cudaStream_t s, int* launches) {
    // The survivor counts live on the device; size the grids for the worst case and let surplus
    // threads exit on the device-side counts.
    const int work = std::max(p.n, p.old.cnt ? p.old_n_list_cap : 0);
    const int blocks = std::max(1, (work + 255) / 256);
    k_load_insert<<<blocks, 256, 0, s>>>(p);
    k_occupancy<<<std::max(1, (p.n + 255) / 256), 256, 0, s>>>(p);
    *launches += 2;
    return cudaGetLastError();
}

// ---- incremental update ------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ bool in_range(const BinRange& r, int x, int y, int z) {
    return x >= r.x0 && x < r.x1 && y >= r.y0 && y < r.y1 && z >= r.z0 && z < r.z1;
}

}  // namespace

// The dirty ranges of an update, written by k_update_begin and read by the two kernels after it.
// They live right behind the counters (LoaderCounters is 32 bytes; the scratch follows).
struct UpdateScratch {
    int n;
    int pad[3];
    BinRange r[2 * kMaxUpdate];
};
__device__ __forceinline__ UpdateScratch* scratch_of(LoaderCounters* ctr) {
    return reinterpret_cast<UpdateScratch*>(ctr + 1);
}

// One block.  Old and new bin ranges of the moved entities -> dirty list; clear the dirty bins; swap
// in the new boxes; keep the counters (survivors, inserts) and the survivor list up to date.
__global__ void __launch_bounds__(256)
k_update_begin(const __grid_constant__ UpdateParams p) {
    const ViewDims& d = p.d;
    UpdateScratch* sc = scratch_of(p.g.ctr);
    __shared__ BinRange s_r[2 * kMaxUpdate];
    __shared__ int s_n;
    if (threadIdx.x == 0) {
        int n = 0;
        LoaderCounters* ctr = p.g.ctr;
        for (int u = 0; u < p.count; u++) {
            const int e = p.first + u;
            int4 fresh = p.fresh[u];
            if (p.keep_sprite_ids) fresh.w = p.g.boxes[e].w;
            const Box ob = unpack_box(p.g.boxes[e]), nb = unpack_box(fresh);
            BinRange og, ng;
            const bool o_surv = cull_and_range(d, ob, og), n_surv = cull_and_range(d, nb, ng);
            const bool o_ins = o_surv && range_volume(og) > 0 && fits_sprite(ob, p.sprite_dims, p.n_sprites);
            const bool n_has = n_surv && range_volume(ng) > 0;
            const bool n_ok = n_has && fits_sprite(nb, p.sprite_dims, p.n_sprites);
            if (o_ins) {
                s_r[n++] = og;
                ctr->n_inserts -= range_volume(og);
            }
            if (n_ok) {
                s_r[n++] = ng;
                ctr->n_inserts += range_volume(ng);
            } else if (n_has) {
                ctr->bad_scene = 1;
                ctr->bad_entity = min(ctr->bad_entity, e);
            }
            ctr->n_survivors += (int)n_surv - (int)o_surv;
            if (n_surv && !o_surv) p.g.survivors[ctr->n_list++] = e;  // listed entities stay listed (clear list)
            p.g.boxes[e] = fresh;
            int4 raw = fresh;
            raw.w = p.raw[e].w;  // the upload buffer keeps the caller's padding bytes
            p.raw[e] = raw;
            if (p.raw_sprite_ids) p.raw_sprite_ids[e] = fresh.w;
        }
        s_n = n;
        sc->n = n;
        for (int k = 0; k < n; k++) sc->r[k] = s_r[k];
    }
    __syncthreads();
    for (int k = 0; k < s_n; k++) {
        const BinRange g = s_r[k];
        const int ny = g.y1 - g.y0, nz = g.z1 - g.z0, vol = range_volume(g);
        for (int t = threadIdx.x; t < vol; t += blockDim.x) {
            const int f = flat_bin(d, g.x0 + t / (ny * nz), g.y0 + (t / nz) % ny, g.z0 + t % nz);
            clear_bin(p.g, f);
            atomicAnd(&p.g.occ4[f >> 3], ~(0xfu << ((f & 7) * 4)));
        }
    }
}

// One thread per entity: re-insert into the dirty bins it spans (each bin once, even where dirty
// ranges overlap).
__global__ void __launch_bounds__(256)
k_update_insert(const __grid_constant__ UpdateParams p) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n) return;
    const ViewDims& d = p.d;
    const UpdateScratch* sc = scratch_of(p.g.ctr);
    const int n = sc->n;
    if (n == 0) return;
    const Box b = unpack_box(p.g.boxes[e]);
    BinRange g;
    if (!cull_and_range(d, b, g)) return;
    if (e >= p.first && e < p.first + p.count && !fits_sprite(b, p.sprite_dims, p.n_sprites)) return;
    for (int k = 0; k < n; k++) {
        const BinRange r = sc->r[k];
        const int x0 = max(g.x0, r.x0), x1 = min(g.x1, r.x1), y0 = max(g.y0, r.y0), y1 = min(g.y1, r.y1);
        const int z0 = max(g.z0, r.z0), z1 = min(g.z1, r.z1);
        for (int x = x0; x < x1; x++)
            for (int y = y0; y < y1; y++)
                for (int z = z0; z < z1; z++) {
                    bool earlier = false;
                    for (int q = 0; q < k; q++) earlier = earlier || in_range(sc->r[q], x, y, z);
                    if (!earlier) insert_into_bin(p.g, flat_bin(d, x, y, z), e);
                }
    }
}

// One block: occupancy nibbles of the dirty bins, then publish the counters.
__global__ void __launch_bounds__(256)
k_update_end(const __grid_constant__ UpdateParams p) {
    const ViewDims& d = p.d;
    const UpdateScratch* sc = scratch_of(p.g.ctr);
    for (int k = 0; k < sc->n; k++) {
        const BinRange g = sc->r[k];
        const int ny = g.y1 - g.y0, nz = g.z1 - g.z0, vol = range_volume(g);
        for (int t = threadIdx.x; t < vol; t += blockDim.x) {
            const int f = flat_bin(d, g.x0 + t / (ny * nz), g.y0 + (t / nz) % ny, g.z0 + t % nz);
            const unsigned keep = p.g.cnt[f] & (kSlots - 1);
            // overlapping dirty ranges write the same value twice: clear-then-or keeps it idempotent
            if (keep) atomicOr(&p.g.occ4[f >> 3], keep << ((f & 7) * 4));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) publish(p.g.ctr, p.host_a, p.host_b);
}

cudaError_t launch_scene_update(const UpdateParams& p, cudaStream_t s, int* launches) {
    k_update_begin<<<1, 256, 0, s>>>(p);
    k_update_insert<<<std::max(1, (p.n + 255) / 256), 256, 0, s>>>(p);
    k_update_end<<<1, 256, 0, s>>>(p);
    *launches += 3;
    return cudaGetLastError();
}

__global__ void k_publish_counters(GridBuffers g, LoaderCounters* host_a, LoaderCounters* host_b) {
    publish(g.ctr, host_a, host_b);
}

cudaError_t launch_publish_counters(const GridBuffers& g, LoaderCounters* host_a, LoaderCounters* host_b, cudaStream_t s) {
    k_publish_counters<<<1, 1, 0, s>>>(g, host_a, host_b);
    return cudaGetLastError();
}

size_t loader_counter_bytes() { return sizeof(LoaderCounters) + sizeof(UpdateScratch); }

}  // namespace par
