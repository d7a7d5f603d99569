// primary.cu — ray generation + ray-sprite intersection + nearest-hit selection:
// replaces trace_hash_for_pixel (/root/reference/src/alternative.cpp:271-383).
//
// Every primary ray has slope <0,-1,+1> and stays in one column of bins (pixel x / 40,
// pixel row / 40) while bin_z runs near to far (alternative.cpp:287-296).  One CTA owns one
// 40x40 screen tile = one bin column:
//   1. stage the column: per-bin counts -> exclusive scan -> the column's entries (in the
//      reference's bin_z-then-slot order) are gathered into shared memory ONCE per tile,
//      pre-digested into the integers the per-pixel test needs;
//   2. one thread per pixel (40 columns x 8 rows per pass, 5 passes) walks that list: 2-D
//      integer hit test (quirk Q6), texel index (Q7), depth key with strict-greater select
//      (Q8), two-adjacent-bins early-out (Q9), miss pixel (Q10), G-buffer record (Q11).
// Integer only; output is the compact G-buffer (par_device.cuh).
#include <climits>

#include "par_kernels.cuh"

namespace par {

constexpr int kNoGroupZ = 0x7fffffff;
constexpr int kStagedSprites = 4;  // depth tables staged in shared memory when the atlas is this small

// Dynamic shared memory layout: int s_cnt[HL], int s_off[HL+1], int4 A[n], int4 B[n], int2 C[n]
// with n <= 7*HL, then (optionally) the depth tables.
template <bool kEmitGroups>
__global__ void __launch_bounds__(kTileThreads)
k_primary(PrimaryParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ViewDims& d = p.d;
    const int max_entries = (kSlots - 1) * d.HL;
    int4* sA = reinterpret_cast<int4*>(smem_raw);  // x0, x1 (exclusive), lo = py+pz, top = py+ey+pz+ez
    int4* sB = sA + max_entries;                   // key0 = py-pz, ey, pz, ybase = py+ey+ez
    int2* sC = reinterpret_cast<int2*>(sB + max_entries);  // entity, sprite<<2 | gap<<1 | first
    int* s_cnt = reinterpret_cast<int*>(sC + max_entries);
    int* s_off = s_cnt + d.HL;
    int* s_depth = s_off + d.HL + 1;
    const bool staged = p.n_sprites <= kStagedSprites;

    const int tid = threadIdx.x;
    const int bx = blockIdx.x % d.HW;
    const int ty = p.tile_row_first + (blockIdx.x / d.HW) * max(d.stripe_n, 1);

    // 1a. counts of the column (the reference's wrapping count is cnt & 7)
    for (int bz = tid; bz < d.HL; bz += blockDim.x)
        s_cnt[bz] = p.cnt[flat_bin(d, bx, ty, bz)] & (kSlots - 1);
    if (staged)
        for (int i = tid; i < p.n_sprites * kTexels; i += blockDim.x) s_depth[i] = p.atlas_depth[i];
    __syncthreads();
    // 1b. exclusive scan over bin_z by warp 0
    if (tid < 32) {
        int carry = 0;
        for (int base = 0; base < d.HL; base += 32) {
            int bz = base + tid;
            int v = bz < d.HL ? s_cnt[bz] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += t;
            }
            if (bz < d.HL) s_off[bz] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (tid == 0) s_off[d.HL] = carry;
    }
    __syncthreads();
    // 1c. gather entries in (bin_z ascending, slot ascending) order
    for (int i = tid; i < d.HL * kSlots; i += blockDim.x) {
        int bz = i >> 3, s = i & 7, c = s_cnt[bz];
        if (s >= c) continue;
        int f = flat_bin(d, bx, ty, bz);
        int ent = p.ids[f * kSlots + (c - 1 - s)];
        Box b = unpack_box(p.boxes[ent]);
        int pos = s_off[bz] + s;
        sA[pos] = make_int4(b.px, b.px + b.ex, b.py + b.pz, b.py + b.ey + b.pz + b.ez);
        sB[pos] = make_int4(b.py - b.pz, b.ey, b.pz, b.py + b.ey + b.ez);
        int first = (s == 0);
        int gap = first && (bz == 0 || s_cnt[bz - 1] == 0);  // an empty bin precedes this one
        sC[pos] = make_int2(ent, b.sprite << 2 | gap << 1 | first);
    }
    __syncthreads();
    const int n = s_off[d.HL];

    // 2. per-pixel walk.  A thread owns 5 pixels of ONE screen column (8 rows apart), so the
    //    x half of the hit test (quirk Q6) is shared by them: the entry list is traversed once per
    //    thread, and per pixel only (best key, winning entry, its depth, run state) is kept; the
    //    G-buffer record is rebuilt from the winning entry at the end.
    const int col = tid % kBin, rsub = tid / kBin;
    const int i = bx * kBin + col;
    const int ra = max(ty * kBin, d.row0), rb = min(ty * kBin + kBin, d.row1);
    int wj[kTileRowsPerThread], best[kTileRowsPerThread], win[kTileRowsPerThread], wdep[kTileRowsPerThread];
    int run[kTileRowsPerThread];
    unsigned any_bits = 0, live = 0;  // bit m: has_intersected in the current bin / pixel still marching
#pragma unroll
    for (int m = 0; m < kTileRowsPerThread; m++) {
        const int j = ty * kBin + rsub + 8 * m;
        wj[m] = (short)(d.H - j);  // world_j, alternative.cpp:280
        best[m] = INT_MIN;         // closest_entity_depth, alternative.cpp:289
        win[m] = -1;               // miss (quirk Q10)
        wdep[m] = 0;
        run[m] = 0;                // intersected_bin_count
        if (j >= ra && j < rb) live |= 1u << m;
    }
    for (int k = 0; k < n && live; k++) {
        const int2 c = sC[k];
        if (c.y & 1) {  // first entry of a bin: close the previous bin (alternative.cpp:368-374)
#pragma unroll
            for (int m = 0; m < kTileRowsPerThread; m++) {
                run[m] += (any_bits >> m) & 1;
                if (run[m] >= 2) live &= ~(1u << m);  // two adjacent hit bins end the march (quirk Q9)
                if (c.y & 2) run[m] = 0;              // an empty bin in between resets the run (298-300)
            }
            any_bits = 0;
            if (!live) break;
        }
        const int4 a = sA[k];
        if (i < a.x || i >= a.y) continue;  // x half of quirk Q6, the same for all 5 pixels
        const int4 b = sB[k];
        const int spr = c.y >> 2;
#pragma unroll
        for (int m = 0; m < kTileRowsPerThread; m++) {
            if (!((live >> m) & 1) || !(wj[m] > a.z && wj[m] <= a.w)) continue;  // y half of Q6
            const int row = a.w - wj[m];
            const int idx = row * kSpriteW + (i - a.x);  // quirk Q7
            const int dep = staged ? s_depth[spr * kTexels + idx] : __ldg(&p.atlas_depth[spr * kTexels + idx]);
            const int key = b.x + min(0, b.y - row) - dep;  // quirk Q8
            if (best[m] < key) {  // strict: ties keep the earlier (bin_z, slot)
                best[m] = key;
                win[m] = k;
                wdep[m] = dep;
                any_bits |= 1u << m;
            }
        }
    }
    int gz[kTileRowsPerThread], oy[kTileRowsPerThread], oz[kTileRowsPerThread];  // for step 3
#pragma unroll
    for (int m = 0; m < kTileRowsPerThread; m++) {
        gz[m] = kNoGroupZ;
        const int j = ty * kBin + rsub + 8 * m;
        if (j < ra || j >= rb) continue;
        int4 out = make_int4(0, 0, 0, -1);  // miss: entity 0, y 0, z 0 (quirk Q10)
        if (win[m] >= 0) {
            const int4 a = sA[win[m]], b = sB[win[m]];
            const int2 c = sC[win[m]];
            const int row = a.w - wj[m], idx = row * kSpriteW + (i - a.x);
            out = make_int4(c.x, b.w - row - wdep[m], b.z + wdep[m], idx | (c.y >> 2) << 10);  // Q11
        }
        p.gbuf[(size_t)j * d.W + i] = out;
        if (kEmitGroups && out.w >= 0) {
            gz[m] = out.z / kBin;  // ray_bin_z, alternative.cpp:727 (C division truncates toward zero)
            oy[m] = out.y;
            oz[m] = out.z;
        }
    }

    // 3. the tile's z-groups (ascending), their pixel counts and the integer bounds of their ray
    //    origins: the work descriptors of the shadow-walk kernel (walks.cu).
    if (!kEmitGroups) return;
    __shared__ int s_group, s_npix, s_min[3], s_max[3];
    const int tile = ty * d.HW + bx;
    const int lane = tid & 31;
    int last = -0x7fffffff - 1, n_groups = 0;
    for (;;) {
        if (tid == 0) {
            s_group = kNoGroupZ;
            s_npix = 0;
            s_min[0] = s_min[1] = s_min[2] = 0x7fffffff;
            s_max[0] = s_max[1] = s_max[2] = -0x7fffffff - 1;
        }
        __syncthreads();
        int mine = kNoGroupZ;
#pragma unroll
        for (int m = 0; m < kTileRowsPerThread; m++)
            if (gz[m] > last) mine = min(mine, gz[m]);
#pragma unroll
        for (int o = 16; o; o >>= 1) mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, o));
        if (lane == 0 && mine != kNoGroupZ) atomicMin(&s_group, mine);
        __syncthreads();
        const int group = s_group;
        if (group == kNoGroupZ) break;
        last = group;
        if (n_groups == kMaxGroups) {  // too many groups: the shade kernel walks this tile itself
            n_groups = -1;
            break;
        }
        int cnt = 0, lo3[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi3[3] = {-0x7fffffff - 1, -0x7fffffff - 1, -0x7fffffff - 1};
#pragma unroll
        for (int m = 0; m < kTileRowsPerThread; m++)
            if (gz[m] == group) {
                cnt++;
                const int o3[3] = {i, oy[m], oz[m]};
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    lo3[a] = min(lo3[a], o3[a]);
                    hi3[a] = max(hi3[a], o3[a]);
                }
            }
#pragma unroll
        for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
#pragma unroll
        for (int a = 0; a < 3; a++) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                lo3[a] = min(lo3[a], __shfl_xor_sync(0xffffffffu, lo3[a], o));
                hi3[a] = max(hi3[a], __shfl_xor_sync(0xffffffffu, hi3[a], o));
            }
        }
        if (lane == 0 && cnt) {
            atomicAdd(&s_npix, cnt);
#pragma unroll
            for (int a = 0; a < 3; a++) {
                atomicMin(&s_min[a], lo3[a]);
                atomicMax(&s_max[a], hi3[a]);
            }
        }
        __syncthreads();
        if (tid == 0) {
            GroupMeta g;
            g.z = group;
            g.npix = s_npix;
#pragma unroll
            for (int a = 0; a < 3; a++) {
                g.omin[a] = s_min[a];
                g.omax[a] = s_max[a];
            }
            p.groups[(size_t)tile * kMaxGroups + n_groups] = g;
        }
        n_groups++;
    }
    if (tid == 0) p.tile_ngroups[tile] = n_groups;
}

size_t primary_smem_bytes(const ViewDims& d, int n_sprites) {
    size_t n = (size_t)(kSlots - 1) * d.HL;
    size_t bytes = n * (16 + 16 + 8) + sizeof(int) * (2 * (size_t)d.HL + 1);
    if (n_sprites <= kStagedSprites) bytes += sizeof(int) * (size_t)n_sprites * kTexels;
    return bytes;
}

cudaError_t configure_primary(size_t smem) {
    if (smem <= 48 * 1024) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k_primary<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return e;
    return cudaFuncSetAttribute(k_primary<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

cudaError_t launch_primary(const PrimaryParams& p, cudaStream_t s) {
    const ViewDims& d = p.d;
    int first, tile_rows;
    owned_tile_rows(d, first, tile_rows);
    if (tile_rows <= 0) return cudaSuccess;
    size_t smem = primary_smem_bytes(d, p.n_sprites);
    if (p.tile_ngroups)
        k_primary<true><<<tile_rows * d.HW, kTileThreads, smem, s>>>(p);
    else
        k_primary<false><<<tile_rows * d.HW, kTileThreads, smem, s>>>(p);
    return cudaGetLastError();
}

}  // namespace par
