// par_api.cu — the extern "C" layer of libpar_b200.so (include/par/par.h): context,
// device memory, streams/events/graphs, and the launch sequence that replaces the reference frame
// loop body /root/reference/src/alternative.cpp:689-760.  No compute happens on the host and
// there is no CPU fallback: without a CUDA device every entry point fails.
#include <algorithm>
#include <numeric>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only: ranges cost a null check unless a profiler injected itself

#include "par/par.h"
#include "par_kernels.cuh"

namespace par {

// Expand a compact G-buffer record (entity, y, z, global texel index) into the reference's
// 28-byte Pixel (normal, colour, y, z, entity; sprites.hpp:53-58) + the sprite-local texel index,
// for the parity checkpoints of par_get_gbuffer / par_render(out_gbuf).
__global__ void __launch_bounds__(256)
k_expand_gbuf(const int4* __restrict__ gbuf, const float4* __restrict__ texel_tab,
              const int4* __restrict__ boxes, const int2* __restrict__ sprite_dims, size_t first, size_t count,
              int* __restrict__ out_pixel7, int* __restrict__ out_texel) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const size_t px = first + t;
    const int4 g = gbuf[px];
    float4 tx = make_float4(0.f, 0.f, 0.f, __uint_as_float(127u | 127u << 8 | 127u << 16));  // alternative.cpp:281
    int texel = -1;
    if (g.w >= 0) {
        tx = texel_tab[g.w];
        texel = g.w - sprite_dims[boxes[g.x].w].x;
    }
    if (out_pixel7) {
        int* o = out_pixel7 + px * 7;
        o[0] = __float_as_int(tx.x);
        o[1] = __float_as_int(tx.y);
        o[2] = __float_as_int(tx.z);
        o[3] = __float_as_int(tx.w);
        o[4] = g.y;
        o[5] = g.z;
        o[6] = g.x;
    }
    if (out_texel) out_texel[px] = texel;
}

// Stripe-major staging frame ([n][T][40][W] uchar4) -> raster frame, 16 bytes per thread.
__global__ void __launch_bounds__(256)
k_unstripe(const uint4* __restrict__ staging, uint4* __restrict__ raster, int W4, int H, int n, int T) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)W4 * H) return;
    const int j = (int)(t / W4), c = (int)(t % W4);
    const int tile = j / kBin;
    const int src_row = ((tile % n) * T + tile / n) * kBin + j % kBin;
    raster[t] = staging[(size_t)src_row * W4 + c];
}

// ---- flags of the multi-GPU frame exchange (par_render_resident with par_exchange_setup) ----
// Every context's frame allocation ends with a footer of two flag rows; rank r owns column r of
// the rows in EVERY other rank's footer (it is the only writer):
//     arrive[r]  frames rank r has finished storing into this frame
//     credit[r]  frames rank r (a consumer) has released: producers may overwrite its frame
// Sequence numbers live on the device (seq[0]: frames produced, seq[1]: credits sent), so the
// kernels take constant arguments and a captured graph can be replayed unchanged.
struct ExchangeFooter {
    unsigned arrive[8];
    unsigned credit[8];
};
struct FlagTargets {
    unsigned* slot[8];
    int n;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// *seq += 1, then publish *seq - lag to every target.  (The kernel boundary before this launch
// ordered the render kernel's peer stores; the release at system scope makes them visible before
// the flag.)  lag = 1 for credits: starting frame k releases frame k - 1.
__global__ void k_flag_signal(FlagTargets t, unsigned* seq, unsigned lag) {
    if (threadIdx.x != 0) return;
    const unsigned v = *seq + 1u;
    *seq = v;
    __threadfence_system();
    for (int k = 0; k < t.n; k++) st_release_sys(t.slot[k], v - lag);
}

// Spin until every own slot has reached *want + delta; advance == 1 also stores the new *want.
// A peer that never answers (a rank died) ends the wait after ~10 s and raises *timeout_flag
// (mapped host memory) instead of hanging the GPU.
__global__ void k_flag_wait(FlagTargets t, unsigned* want, unsigned delta, int advance, int* timeout_flag) {
    if (threadIdx.x != 0) return;
    const unsigned v = *want + delta;
    if (advance) *want = v;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int k = 0; k < t.n; k++)
        while ((int)(ld_acquire_sys(t.slot[k]) - v) < 0) {
            __nanosleep(100);
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > 10000000000ull) {
                *timeout_flag = 1;
                return;
            }
        }
}

}  // namespace par

using namespace par;

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(g_err, sizeof g_err, fmt, a, b);
    return code;
}

#define PAR_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fail(e_ == cudaErrorMemoryAllocation ? PAR_ERR_OUT_OF_MEMORY : PAR_ERR_CUDA, \
                        "%s: %s", #call, cudaGetErrorString(e_));                        \
    } while (0)

constexpr int kMaxChunks = 8;  // row chunks of the pipelined readback in par_render

// NVTX range over the host side of an entry point (shows up in Nsight Systems / ncu --nvtx timelines).
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

struct par_ctx {
    par_config cfg;
    ViewDims d;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // D2H of finished row chunks, overlapping the next chunk
    cudaEvent_t ev_build0 = nullptr, ev_build1 = nullptr, ev_f0 = nullptr, ev_f2 = nullptr, ev_copy = nullptr;
    cudaEvent_t ev_chunk[kMaxChunks][2] = {};  // before / after the render kernel of a chunk
    int n_chunks = 1;
    // scene: the upload buffers and two generations of the grid
    int4* d_raw = nullptr;
    int* d_sprite_ids = nullptr;
    GridBuffers gen[2] = {};
    int cur = 0;                 // generation the latest build went into
    int list_bound[2] = {0, 0};  // host-side upper bound of each generation's survivor list
    bool gen_dirty[2] = {false, false};  // generation holds a build that has not been cleared yet
    int cap_entities = 0, n_entities = 0;
    bool has_sprite_ids = false;
    LoaderCounters* h_ctr = nullptr;  // pinned: [0] latest build, [1 + slot] pipelined frames
    // atlas
    int* d_atlas_depth = nullptr;
    float4* d_texel_tab = nullptr;
    int2* d_sprite_dims = nullptr;
    int n_sprites = 0, n_palette = 0, atlas_texels = 0;
    // frame
    int4* d_gbuf = nullptr;     // compact G-buffer, allocated when a caller first asks for the checkpoint
    bool gbuf_valid = false;    // d_gbuf holds the latest frame's records
    unsigned char* d_frame_block = nullptr;  // frame 0 | footer | frame 1 (allocated on first pipelined use)
    uchar4* d_frame = nullptr;
    ExchangeFooter* d_footer = nullptr;
    size_t footer_offset = 0;
    int* d_expanded = nullptr;  // W*H*7 ints, lazily allocated
    int* d_texel = nullptr;     // W*H ints, lazily allocated
    unsigned* d_tile_cost = nullptr;  // cycles per tile of the latest frame
    int* d_tile_order = nullptr;      // longest-first order computed from them
    bool order_valid = false;   // d_tile_order holds an order of this context's tiles
    bool cost_valid = false;    // d_tile_cost holds the costs of the latest ordered frame
    bool order_fresh = false;   // ... and d_tile_order was computed from exactly those costs
    bool order_forked = false;  // the order kernel is in flight on aux_stream: join before the render kernel
    cudaStream_t aux_stream = nullptr;  // side branch of a frame: the tile-order kernel runs beside the scene loader
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_side = nullptr;
    // par_render_resident, overlapped form: the frame of grid generation p renders while generation p ^ 1 is rebuilt
    // on the side branch.  Tile costs / orders are kept per generation parity (a frame's order comes from the costs of
    // the previous frame of the same parity, sorted during the frame in between).
    unsigned* d_res_cost[2] = {};
    int* d_res_order[2] = {};
    bool res_cost_valid[2] = {false, false}, res_order_valid[2] = {false, false};
    unsigned long long* d_phase_cycles = nullptr;  // debug only
    bool scene_set = false, frame_valid = false, build_timed = false, frame_timed = false;
    int launches_build = 0, launches_frame = 0, last_n_lights = 0;
    float ambient = 0.25f;
    int cta_slots = 0;     // render-kernel CTAs the device holds at a time (SMs x CTAs per SM)
    size_t out_pitch = 0;  // row pitch of host frames (0 = packed)
    // fused multi-GPU frame exchange: raster frames of the other ranks, mapped into this process
    uchar4* peer_frame[8] = {};
    bool peer_is_ipc[8] = {};
    int n_peers = 0;  // entries of peer_frame in use (own rank's entry stays NULL)
    bool exchange_on = false;
    int exchange_root = -1;
    unsigned* d_seq = nullptr;  // [0] frames produced, [1] frames begun as a consumer, [2] arrivals awaited
    int* h_exchange_timeout = nullptr;  // pinned: raised by a flag wait that gave up
    // pipelined frames (par_submit_frame / par_wait_frame): two slots, each with its own device frame
    uchar4* d_frame_alt = nullptr;  // slot 1's frame, allocated on first use (slot 0 uses d_frame)
    cudaEvent_t ev_slot_begin[2] = {}, ev_slot_kernels[2] = {}, ev_slot_done[2] = {};
    int slots_in_flight = 0, slot_oldest = 0, slot_next = 0, slot_lights[2] = {};
    // cursor probe: [0] frames of the synchronous calls, [1 + slot] pipelined frames (pinned, 7 ints each)
    int cursor_x = -1, cursor_y = -1;
    int (*h_probe)[7] = nullptr;
    int (*probe_slot)[7] = nullptr;    // where the frame being submitted also stores its probe
    int (*probe_latest)[7] = nullptr;  // record of the frame par_cursor_pixel reports
    bool capturing = false;            // work is being captured into a graph: no timing events
    cudaGraphExec_t frame_exec = nullptr;  // upload + build + kernels of one pipelined frame as ONE graph launch
    // par_render_resident: one executable graph per grid generation, valid while the key matches
    cudaGraphExec_t resident_exec[2] = {};
    bool resident_ok[2] = {false, false};
    par_light resident_lights[PAR_MAX_LIGHTS];
    int resident_n_lights = -1;
    unsigned long long resident_epoch[2] = {0, 0};
    unsigned long long epoch = 1;  // bumped by everything that invalidates captured graphs
    float last_kernel_ms = 0.f;  // render kernel of the previous par_render (pipelining heuristic)
    int readback_chunks = 3;  // row chunks of the pipelined readback (PAR_READBACK_CHUNKS overrides)
    int debug_flags = 0;  // from the PAR_DEBUG_FLAGS environment variable (developer A/B switches)
    std::vector<int2> h_sprite_dims;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

// Timing events of the synchronous calls.  While work is captured into a CUDA graph they are
// skipped: an event recorded by a graph node can no longer be queried from the host.
cudaError_t record_timing(par_ctx* c, cudaEvent_t ev) {
    return c->capturing ? cudaSuccess : cudaEventRecord(ev, c->stream);
}

size_t frame_bytes(const par_ctx* c) { return sizeof(uchar4) * (size_t)c->d.W * c->d.H; }
size_t host_pitch(const par_ctx* c) { return c->out_pitch ? c->out_pitch : sizeof(par_color) * (size_t)c->d.W; }

// Build the grid from the resident upload buffers into the generation that is not current.
int run_loader(par_ctx* c, LoaderCounters* slot_ctr = nullptr) {
    c->launches_build = 0;
    const int nxt = c->cur ^ 1;
    LoaderParams lp;
    lp.d = c->d;
    lp.raw = c->d_raw;
    lp.sprite_ids = c->has_sprite_ids ? c->d_sprite_ids : nullptr;
    lp.sprite_dims = c->d_sprite_dims;
    lp.n = c->n_entities;
    lp.n_sprites = c->n_sprites;
    lp.cur = c->gen[nxt];
    lp.old = c->gen[c->cur];
    // (a captured build is replayed later, when the other generation WILL hold a build: always clear)
    lp.old_n_list_cap = c->capturing ? c->cap_entities : c->list_bound[c->cur];
    if (!c->gen_dirty[c->cur] && !c->capturing) lp.old.cnt = nullptr;  // nothing to clear
    lp.host_a = c->h_ctr;
    lp.host_b = slot_ctr;
    PAR_CUDA(record_timing(c, c->ev_build0));
    if (c->gen_dirty[nxt]) {  // left behind by an overlapped resident frame: the target generation is cleared on its own first
        PAR_CUDA(launch_clear_touched(c->d, c->gen[nxt], c->capturing ? c->cap_entities : c->list_bound[nxt], c->stream,
                                      &c->launches_build));
        c->gen_dirty[nxt] = false;
        c->list_bound[nxt] = 0;
    }
    PAR_CUDA(launch_scene_loader(lp, c->stream, &c->launches_build));
    PAR_CUDA(record_timing(c, c->ev_build1));
    c->gen_dirty[c->cur] = false;  // cleared by this launch (its counters are re-armed by k_occupancy)
    c->list_bound[c->cur] = 0;
    c->gen_dirty[nxt] = true;
    c->list_bound[nxt] = c->n_entities;
    c->cur = nxt;
    c->build_timed = !c->capturing;
    c->scene_set = true;
    c->frame_valid = false;
    c->gbuf_valid = false;
    return PAR_OK;
}

int bad_scene_error(const LoaderCounters& lc) {
    char who[64];
    snprintf(who, sizeof who, "%d", lc.bad_entity);
    return fail(PAR_ERR_BAD_SCENE,
                "scene: entity %s spans a bin but has an extent.x larger than its sprite's width, an "
                "extent.y+extent.z larger than its height, a negative extent or a sprite id outside the "
                "atlas (would index outside the sprite, alternative.cpp:330)%s", who);
}

int check_scene_flag(par_ctx* c) { return c->h_ctr->bad_scene ? bad_scene_error(*c->h_ctr) : PAR_OK; }

// Image rows the context renders (its band, restricted to its stripes).
// Pixels the context renders (its band, restricted to its stripes).
uint64_t owned_pixel_count(const par_ctx* c) {
    uint64_t px = 0;
    int first, count;
    owned_tile_rows(c->d, first, count);
    const int seg = stripe_segments(c->d);
    for (int v = first, q = 0; q < count; q++, v += c->d.stripe_n) {
        const int t = v / seg;
        const int a = t * kBin > c->d.row0 ? t * kBin : c->d.row0;
        const int b = (t + 1) * kBin < c->d.row1 ? (t + 1) * kBin : c->d.row1;
        px += (uint64_t)(b - a) * (uint64_t)(c->d.W / seg);
    }
    return px;
}

void free_grid(GridBuffers& g) {
    cudaFree(g.cnt);
    cudaFree(g.ids);
    cudaFree(g.occ4);
    cudaFree(g.ctr);
    g.cnt = g.ids = nullptr;
    g.occ4 = nullptr;
    g.ctr = nullptr;
}

void invalidate_graphs(par_ctx* c) { c->epoch++; }

}  // namespace

extern "C" {

const char* par_last_error(void) { return g_err; }
const char* par_version(void) { return "par_b200 0.2 (sm_100a)"; }

int par_create(par_ctx** out, const par_config* cfg) {
    if (!out || !cfg) return fail(PAR_ERR_INVALID_ARG, "par_create: null argument%s%s");
    *out = nullptr;
    const int B = kBin;
    if (cfg->width <= 0 || cfg->height <= 0 || cfg->length <= 0 || cfg->width % B ||
        cfg->height % B || cfg->length % B || cfg->width > PAR_MAX_VIEW ||
        cfg->height > PAR_MAX_VIEW || cfg->length > PAR_MAX_VIEW)
        return fail(PAR_ERR_INVALID_ARG,
                    "par_create: width/height/length must be positive multiples of 40 and <= 12800%s%s");
    int row0 = cfg->row_begin, row1 = cfg->row_end;
    if (row0 == 0 && row1 == 0) row1 = cfg->height;
    if (row0 < 0 || row1 > cfg->height || row0 >= row1)
        return fail(PAR_ERR_INVALID_ARG, "par_create: bad row band%s%s");
    if (cfg->stripe_count < 0 || cfg->stripe_count > 64 ||
        (cfg->stripe_count > 1 && (cfg->stripe_index < 0 || cfg->stripe_index >= cfg->stripe_count)))
        return fail(PAR_ERR_INVALID_ARG, "par_create: bad stripe_count / stripe_index%s%s");
    const int split = cfg->stripe_split > 1 ? cfg->stripe_split : 1;
    if (cfg->stripe_split < 0 || split > 8 || (cfg->width / B) % split != 0 || (split > 1 && cfg->stripe_count < 2))
        return fail(PAR_ERR_INVALID_ARG,
                    "par_create: stripe_split must be 0..8, divide width / 40 and go with stripe_count >= 2%s%s");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(PAR_ERR_NO_DEVICE, "par_create: no CUDA device (this library has no CPU fallback)%s%s");
    }
    if (cfg->device < 0 || cfg->device >= n_dev)
        return fail(PAR_ERR_INVALID_ARG, "par_create: device ordinal out of range%s%s");
    cudaDeviceProp prop;
    PAR_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(PAR_ERR_NO_DEVICE, "par_create: device %s is not sm_100 (kernels are built for sm_100a only)%s",
                    prop.name);
    if (tile_smem_bytes() > prop.sharedMemPerBlockOptin)
        return fail(PAR_ERR_NO_DEVICE, "par_create: device %s offers too little shared memory per block%s", prop.name);

    par_ctx* c = new (std::nothrow) par_ctx;
    if (!c) return fail(PAR_ERR_OUT_OF_MEMORY, "par_create: host allocation failed%s%s");
    c->cfg = *cfg;
    c->ambient = cfg->ambient == 0.f ? 0.25f : cfg->ambient;
    if (const char* e = getenv("PAR_DEBUG_FLAGS")) c->debug_flags = atoi(e);
    if (const char* e = getenv("PAR_READBACK_CHUNKS")) c->readback_chunks = atoi(e) < 1 ? 1 : atoi(e) > kMaxChunks ? kMaxChunks : atoi(e);
    ViewDims& d = c->d;
    d.W = cfg->width;
    d.H = cfg->height;
    d.L = cfg->length;
    d.HW = d.W / B;
    d.HH = d.H / B;
    d.HL = d.L / B;
    d.V = d.HW * d.HH * d.HL;
    d.row0 = row0;
    d.row1 = row1;
    d.stripe_n = cfg->stripe_count > 1 ? cfg->stripe_count : 1;
    d.stripe_i = cfg->stripe_count > 1 ? cfg->stripe_index : 0;
    d.stripe_s = split;
    d.stripe_rot = std::lcm(d.stripe_n, split);

    DeviceGuard guard(cfg->device);
    int rc = [&]() -> int {
        PAR_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        c->stream = c->own_stream;
        PAR_CUDA(cudaEventCreate(&c->ev_build0));
        PAR_CUDA(cudaEventCreate(&c->ev_build1));
        PAR_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        PAR_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
        PAR_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        PAR_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        PAR_CUDA(cudaEventCreateWithFlags(&c->ev_side, cudaEventDisableTiming));
        PAR_CUDA(cudaEventCreate(&c->ev_f0));
        PAR_CUDA(cudaEventCreate(&c->ev_f2));
        PAR_CUDA(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
        for (auto& ev : c->ev_chunk)
            for (cudaEvent_t& e : ev) PAR_CUDA(cudaEventCreate(&e));
        for (GridBuffers& g : c->gen) {
            PAR_CUDA(cudaMalloc(&g.cnt, sizeof(int) * (size_t)d.V));
            PAR_CUDA(cudaMalloc(&g.ids, sizeof(int) * (size_t)d.V * kSlots));
            PAR_CUDA(cudaMalloc(&g.occ4, sizeof(unsigned) * (((size_t)d.V + 7) / 8)));
            PAR_CUDA(cudaMalloc(&g.ctr, loader_counter_bytes()));
            PAR_CUDA(cudaMemsetAsync(g.ctr, 0, loader_counter_bytes(), c->stream));
            PAR_CUDA(launch_clear_grid(g, d.V, c->stream));
        }
        PAR_CUDA(cudaMallocHost(&c->h_ctr, 3 * sizeof(LoaderCounters)));
        memset(c->h_ctr, 0, 3 * sizeof(LoaderCounters));
        PAR_CUDA(cudaMallocHost(&c->h_probe, 3 * sizeof(int[7])));
        memset(c->h_probe, 0, 3 * sizeof(int[7]));
        for (int k = 0; k < 2; k++) {
            PAR_CUDA(cudaEventCreate(&c->ev_slot_begin[k]));
            PAR_CUDA(cudaEventCreate(&c->ev_slot_kernels[k]));
            PAR_CUDA(cudaEventCreate(&c->ev_slot_done[k]));
        }
        const size_t tiles = (size_t)d.HW * d.HH;
        PAR_CUDA(cudaMalloc(&c->d_tile_cost, sizeof(unsigned) * tiles));
        PAR_CUDA(cudaMalloc(&c->d_tile_order, sizeof(int) * tiles));
        PAR_CUDA(cudaMemsetAsync(c->d_tile_cost, 0, sizeof(unsigned) * tiles, c->stream));
        for (int k = 0; k < 2; k++) {
            PAR_CUDA(cudaMalloc(&c->d_res_cost[k], sizeof(unsigned) * tiles));
            PAR_CUDA(cudaMalloc(&c->d_res_order[k], sizeof(int) * tiles));
            PAR_CUDA(cudaMemsetAsync(c->d_res_cost[k], 0, sizeof(unsigned) * tiles, c->stream));
        }
        PAR_CUDA(cudaMalloc(&c->d_seq, 4 * sizeof(unsigned)));
        PAR_CUDA(cudaMemsetAsync(c->d_seq, 0, 4 * sizeof(unsigned), c->stream));
        PAR_CUDA(cudaMallocHost(&c->h_exchange_timeout, sizeof(int)));
        *c->h_exchange_timeout = 0;
        // the frame and, right behind it, the footer with the exchange flags (one allocation: a single
        // CUDA IPC handle gives a peer both)
        c->footer_offset = (frame_bytes(c) + 255) & ~(size_t)255;
        PAR_CUDA(cudaMalloc(&c->d_frame_block, c->footer_offset + 256));
        c->d_frame = reinterpret_cast<uchar4*>(c->d_frame_block);
        c->d_footer = reinterpret_cast<ExchangeFooter*>(c->d_frame_block + c->footer_offset);
        PAR_CUDA(cudaMemsetAsync(c->d_frame_block, 0, c->footer_offset + 256, c->stream));
        PAR_CUDA(configure_tile());
        c->cta_slots = prop.multiProcessorCount * tile_ctas_per_sm();
        return PAR_OK;
    }();
    if (rc != PAR_OK) {
        par_destroy(c);
        return rc;
    }
    *out = c;
    return PAR_OK;
}

void par_destroy(par_ctx* c) {
    if (!c) return;
    DeviceGuard guard(c->cfg.device);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    if (c->stream && c->stream != c->own_stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    if (c->aux_stream) cudaStreamSynchronize(c->aux_stream);
    cudaFree(c->d_raw);
    cudaFree(c->d_sprite_ids);
    for (GridBuffers& g : c->gen) {
        free_grid(g);
        cudaFree(g.boxes);
        cudaFree(g.survivors);
    }
    if (c->h_ctr) cudaFreeHost(c->h_ctr);
    if (c->h_probe) cudaFreeHost(c->h_probe);
    if (c->h_exchange_timeout) cudaFreeHost(c->h_exchange_timeout);
    cudaFree(c->d_atlas_depth);
    cudaFree(c->d_texel_tab);
    cudaFree(c->d_sprite_dims);
    for (int r = 0; r < 8; r++)
        if (c->peer_frame[r] && c->peer_is_ipc[r]) cudaIpcCloseMemHandle(c->peer_frame[r]);
    cudaFree(c->d_tile_cost);
    cudaFree(c->d_tile_order);
    for (int k = 0; k < 2; k++) {
        cudaFree(c->d_res_cost[k]);
        cudaFree(c->d_res_order[k]);
    }
    cudaFree(c->d_seq);
    cudaFree(c->d_gbuf);
    cudaFree(c->d_frame_block);
    cudaFree(c->d_frame_alt);
    if (c->frame_exec) cudaGraphExecDestroy(c->frame_exec);
    for (cudaGraphExec_t& e : c->resident_exec)
        if (e) cudaGraphExecDestroy(e);
    for (int k = 0; k < 2; k++) {
        if (c->ev_slot_begin[k]) cudaEventDestroy(c->ev_slot_begin[k]);
        if (c->ev_slot_kernels[k]) cudaEventDestroy(c->ev_slot_kernels[k]);
        if (c->ev_slot_done[k]) cudaEventDestroy(c->ev_slot_done[k]);
    }
    cudaFree(c->d_expanded);
    cudaFree(c->d_texel);
    cudaFree(c->d_phase_cycles);
    cudaEvent_t evs[] = {c->ev_build0, c->ev_build1, c->ev_f0, c->ev_f2, c->ev_copy};
    for (cudaEvent_t e : evs)
        if (e) cudaEventDestroy(e);
    for (auto& ev : c->ev_chunk)
        for (cudaEvent_t e : ev)
            if (e) cudaEventDestroy(e);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_side) cudaEventDestroy(c->ev_side);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int par_set_stream(par_ctx* c, void* cuda_stream) {
    if (!c) return fail(PAR_ERR_INVALID_ARG, "par_set_stream: null context%s%s");
    DeviceGuard guard(c->cfg.device);
    PAR_CUDA(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->own_stream;
    return PAR_OK;
}

void* par_get_stream(par_ctx* c) { return c ? static_cast<void*>(c->stream) : nullptr; }

int par_sync(par_ctx* c) {
    if (!c) return fail(PAR_ERR_INVALID_ARG, "par_sync: null context%s%s");
    DeviceGuard guard(c->cfg.device);
    PAR_CUDA(cudaStreamSynchronize(c->stream));
    if (*c->h_exchange_timeout) {
        *c->h_exchange_timeout = 0;
        return fail(PAR_ERR_CUDA, "par_sync: the multi-GPU frame exchange timed out waiting for a peer's flag%s%s");
    }
    return c->scene_set ? check_scene_flag(c) : PAR_OK;
}

void* par_alloc_host(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        fail(PAR_ERR_OUT_OF_MEMORY, "par_alloc_host: cudaMallocHost failed%s%s");
        return nullptr;
    }
    return p;
}

void par_free_host(void* p) {
    if (p) cudaFreeHost(p);
}

// ---- atlas ---------------------------------------------------------------------------------------------
// Compact device atlas: depth[texels], one float4 per texel (normal + the palette colour the texel
// selects — the palette is applied here, once, instead of per shaded pixel) and per sprite the
// texel base and size.
static int set_atlas_impl(par_ctx* c, int n_sprites, const int32_t* widths, const int32_t* heights,
                          const int32_t* color, const int32_t* depth, const float* normal,
                          const par_color* palette, int n_palette, const char* who) {
    if (!c || !widths || !heights || !color || !depth || !normal || !palette || n_sprites <= 0 || n_palette <= 0 ||
        n_palette > 256 || n_sprites > (1 << 20))
        return fail(PAR_ERR_INVALID_ARG, "%s: bad argument%s", who);
    size_t nt = 0;
    std::vector<int2> dims((size_t)n_sprites);
    for (int s = 0; s < n_sprites; s++) {
        if (widths[s] < 1 || heights[s] < 1 || widths[s] > PAR_MAX_SPRITE_DIM || heights[s] > PAR_MAX_SPRITE_DIM)
            return fail(PAR_ERR_INVALID_ARG, "%s: sprite width/height must be in 1..1024%s", who);
        dims[(size_t)s] = make_int2((int)nt, widths[s] | heights[s] << 16);
        nt += (size_t)widths[s] * heights[s];
        if (nt > (size_t)INT_MAX / 2) return fail(PAR_ERR_INVALID_ARG, "%s: atlas too large (>= 2^30 texels)%s", who);
    }
    std::vector<float4> tab(nt);
    for (size_t t = 0; t < nt; t++) {
        const int ci = color[t];
        if (ci < 0 || ci >= n_palette)
            return fail(PAR_ERR_INVALID_ARG, "%s: sprite colour index outside the palette%s", who);
        if (depth[t] < -PAR_MAX_SPRITE_DEPTH || depth[t] > PAR_MAX_SPRITE_DEPTH)
            return fail(PAR_ERR_INVALID_ARG, "%s: sprite depth outside [-4095, 4095]%s", who);
        unsigned rgba;
        memcpy(&rgba, &palette[ci], 4);
        float w;
        memcpy(&w, &rgba, 4);
        tab[t] = make_float4(normal[3 * t], normal[3 * t + 1], normal[3 * t + 2], w);
    }
    DeviceGuard guard(c->cfg.device);
    PAR_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_atlas_depth);
    cudaFree(c->d_texel_tab);
    cudaFree(c->d_sprite_dims);
    c->d_atlas_depth = nullptr;
    c->d_texel_tab = nullptr;
    c->d_sprite_dims = nullptr;
    c->n_sprites = 0;
    invalidate_graphs(c);
    PAR_CUDA(cudaMalloc(&c->d_atlas_depth, nt * sizeof(int)));
    PAR_CUDA(cudaMalloc(&c->d_texel_tab, nt * sizeof(float4)));
    PAR_CUDA(cudaMalloc(&c->d_sprite_dims, (size_t)n_sprites * sizeof(int2)));
    PAR_CUDA(cudaMemcpy(c->d_atlas_depth, depth, nt * sizeof(int), cudaMemcpyHostToDevice));
    PAR_CUDA(cudaMemcpy(c->d_texel_tab, tab.data(), nt * sizeof(float4), cudaMemcpyHostToDevice));
    PAR_CUDA(cudaMemcpy(c->d_sprite_dims, dims.data(), (size_t)n_sprites * sizeof(int2), cudaMemcpyHostToDevice));
    c->h_sprite_dims = dims;
    c->n_sprites = n_sprites;
    c->n_palette = n_palette;
    c->atlas_texels = (int)nt;
    c->frame_valid = false;
    c->gbuf_valid = false;
    return PAR_OK;
}

int par_set_atlas_sized(par_ctx* c, int n_sprites, const int32_t* widths, const int32_t* heights,
                        const int32_t* color, const int32_t* depth, const float* normal,
                        const par_color* palette, int n_palette) {
    try {
        return set_atlas_impl(c, n_sprites, widths, heights, color, depth, normal, palette, n_palette, "par_set_atlas_sized");
    } catch (const std::bad_alloc&) {
        return fail(PAR_ERR_OUT_OF_MEMORY, "par_set_atlas_sized: host allocation failed%s%s");
    }
}

int par_set_atlas(par_ctx* c, const par_sprite* sprites, int n_sprites, const par_color* palette,
                  int n_palette) {
    if (!c || !sprites || !palette || n_sprites <= 0 || n_sprites > (1 << 20))
        return fail(PAR_ERR_INVALID_ARG, "par_set_atlas: bad argument%s%s");
    try {  // the reference's 20x40 Sprite records (sprites.hpp:67-71) as a sized atlas
        const size_t nt = (size_t)n_sprites * PAR_SPRITE_TEXELS;
        std::vector<int32_t> w((size_t)n_sprites, PAR_SPRITE_W), h((size_t)n_sprites, PAR_SPRITE_H), color(nt), depth(nt);
        std::vector<float> normal(3 * nt);
        for (int s = 0; s < n_sprites; s++) {
            memcpy(&color[(size_t)s * PAR_SPRITE_TEXELS], sprites[s].color, sizeof sprites[s].color);
            memcpy(&depth[(size_t)s * PAR_SPRITE_TEXELS], sprites[s].depth, sizeof sprites[s].depth);
            memcpy(&normal[3 * (size_t)s * PAR_SPRITE_TEXELS], sprites[s].normal, sizeof sprites[s].normal);
        }
        return set_atlas_impl(c, n_sprites, w.data(), h.data(), color.data(), depth.data(), normal.data(), palette,
                              n_palette, "par_set_atlas");
    } catch (const std::bad_alloc&) {
        return fail(PAR_ERR_OUT_OF_MEMORY, "par_set_atlas: host allocation failed%s%s");
    }
}

// ---- scene ---------------------------------------------------------------------------------------------
// Room for n entities in the scene buffers (synchronises the stream when it has to reallocate).
static int reserve_entities(par_ctx* c, int n) {
    if (n <= c->cap_entities) return PAR_OK;
    PAR_CUDA(cudaStreamSynchronize(c->stream));
    const size_t cap = (size_t)n + (size_t)n / 8 + 64;
    int4* raw = nullptr;
    int* ids = nullptr;
    PAR_CUDA(cudaMalloc(&raw, sizeof(int4) * cap));
    PAR_CUDA(cudaMalloc(&ids, sizeof(int) * cap));
    if (c->n_entities > 0 && c->d_raw) {  // keep the resident scene (par_update_entities may follow)
        PAR_CUDA(cudaMemcpy(raw, c->d_raw, sizeof(int4) * (size_t)c->n_entities, cudaMemcpyDeviceToDevice));
        PAR_CUDA(cudaMemcpy(ids, c->d_sprite_ids, sizeof(int) * (size_t)c->n_entities, cudaMemcpyDeviceToDevice));
    }
    cudaFree(c->d_raw);
    cudaFree(c->d_sprite_ids);
    c->d_raw = raw;
    c->d_sprite_ids = ids;
    for (int k = 0; k < 2; k++) {
        GridBuffers& g = c->gen[k];
        int4* boxes = nullptr;
        int* surv = nullptr;
        PAR_CUDA(cudaMalloc(&boxes, sizeof(int4) * cap));
        PAR_CUDA(cudaMalloc(&surv, sizeof(int) * cap));
        if (c->cap_entities > 0) {  // a dirty generation is cleared through its box records and survivor list
            PAR_CUDA(cudaMemcpy(boxes, g.boxes, sizeof(int4) * (size_t)c->cap_entities, cudaMemcpyDeviceToDevice));
            PAR_CUDA(cudaMemcpy(surv, g.survivors, sizeof(int) * (size_t)c->cap_entities, cudaMemcpyDeviceToDevice));
        }
        cudaFree(g.boxes);
        cudaFree(g.survivors);
        g.boxes = boxes;
        g.survivors = surv;
    }
    c->cap_entities = (int)cap;
    invalidate_graphs(c);
    return PAR_OK;
}

static int set_scene_impl(par_ctx* c, const par_aabb* aabbs, const int32_t* sprite_ids, int n,
                          LoaderCounters* slot_ctr) {
    if (!c || n < 0 || n > (1 << 26) || (n > 0 && !aabbs))
        return fail(PAR_ERR_INVALID_ARG, "par_set_scene: bad argument (at most 2^26 entities)%s%s");
    if (c->n_sprites == 0) return fail(PAR_ERR_STATE, "par_set_scene: call par_set_atlas first%s%s");
    DeviceGuard guard(c->cfg.device);
    if (n > c->cap_entities) {
        if (c->capturing) return fail(PAR_ERR_STATE, "par_submit_frame: internal: capacity must grow before the capture%s%s");
        int rc = reserve_entities(c, n);
        if (rc != PAR_OK) return rc;
    }
    if (n != c->n_entities || c->has_sprite_ids != (sprite_ids != nullptr)) invalidate_graphs(c);
    c->n_entities = n;
    c->has_sprite_ids = sprite_ids != nullptr;
    if (n > 0) {
        PAR_CUDA(cudaMemcpyAsync(c->d_raw, aabbs, sizeof(par_aabb) * (size_t)n, cudaMemcpyHostToDevice,
                                 c->stream));
        if (sprite_ids)
            PAR_CUDA(cudaMemcpyAsync(c->d_sprite_ids, sprite_ids, sizeof(int) * (size_t)n,
                                     cudaMemcpyHostToDevice, c->stream));
    }
    return run_loader(c, slot_ctr);
}

int par_set_scene(par_ctx* c, const par_aabb* aabbs, const int32_t* sprite_ids, int n) {
    NvtxRange nvtx("par_set_scene");
    return set_scene_impl(c, aabbs, sprite_ids, n, nullptr);
}

int par_rebuild_grid(par_ctx* c) {
    NvtxRange nvtx("par_rebuild_grid");
    if (!c) return fail(PAR_ERR_INVALID_ARG, "par_rebuild_grid: null context%s%s");
    if (!c->scene_set) return fail(PAR_ERR_STATE, "par_rebuild_grid: no scene resident%s%s");
    DeviceGuard guard(c->cfg.device);
    return run_loader(c);
}

// Patch entities [first, first + count) of the resident scene (see par.h).
static int update_entities_impl(par_ctx* c, int first, int count, const par_aabb* aabbs, const int32_t* sprite_ids,
                                LoaderCounters* slot_ctr, const char* who) {
    if (!c || first < 0 || count < 0 || (count > 0 && !aabbs))
        return fail(PAR_ERR_INVALID_ARG, "%s: bad argument%s", who);
    if (!c->scene_set) return fail(PAR_ERR_STATE, "%s: no scene resident (call par_set_scene first)%s", who);
    if ((long long)first + count > c->n_entities)
        return fail(PAR_ERR_INVALID_ARG, "%s: range exceeds the resident scene%s", who);
    DeviceGuard guard(c->cfg.device);
    if (count == 0) {  // nothing moves: the frame still gets the counters of the grid it uses
        c->launches_build = 0;
        if (slot_ctr) {
            PAR_CUDA(launch_publish_counters(c->gen[c->cur], nullptr, slot_ctr, c->stream));
            c->launches_build = 1;
        }
        return PAR_OK;
    }
    if (sprite_ids && !c->has_sprite_ids) {  // the scene had none so far: materialise the implicit zeros
        PAR_CUDA(cudaMemsetAsync(c->d_sprite_ids, 0, sizeof(int) * (size_t)c->n_entities, c->stream));
        c->has_sprite_ids = true;
        invalidate_graphs(c);
    }
    const bool in_place = count <= kMaxUpdate && c->list_bound[c->cur] + count <= c->cap_entities;
    if (!in_place) {  // big update: upload the range, re-bin everything on the device
        PAR_CUDA(cudaMemcpyAsync(c->d_raw + first, aabbs, sizeof(par_aabb) * (size_t)count, cudaMemcpyHostToDevice, c->stream));
        if (sprite_ids)
            PAR_CUDA(cudaMemcpyAsync(c->d_sprite_ids + first, sprite_ids, sizeof(int) * (size_t)count, cudaMemcpyHostToDevice, c->stream));
        return run_loader(c, slot_ctr);
    }
    UpdateParams up;
    up.d = c->d;
    up.g = c->gen[c->cur];
    up.sprite_dims = c->d_sprite_dims;
    up.n = c->n_entities;
    up.n_sprites = c->n_sprites;
    up.count = count;
    up.first = first;
    for (int u = 0; u < count; u++) {
        memcpy(&up.fresh[u], &aabbs[u], sizeof(int4));
        up.fresh[u].w = sprite_ids ? sprite_ids[u] : 0;
    }
    up.keep_sprite_ids = sprite_ids == nullptr;
    up.raw = c->d_raw;
    up.raw_sprite_ids = c->has_sprite_ids ? c->d_sprite_ids : nullptr;
    up.host_a = c->h_ctr;
    up.host_b = slot_ctr;
    c->launches_build = 0;
    PAR_CUDA(record_timing(c, c->ev_build0));
    PAR_CUDA(launch_scene_update(up, c->stream, &c->launches_build));
    PAR_CUDA(record_timing(c, c->ev_build1));
    c->list_bound[c->cur] += count;
    c->build_timed = !c->capturing;
    c->frame_valid = false;
    c->gbuf_valid = false;
    return PAR_OK;
}

int par_update_entities(par_ctx* c, int first, int count, const par_aabb* aabbs, const int32_t* sprite_ids) {
    NvtxRange nvtx("par_update_entities");
    return update_entities_impl(c, first, count, aabbs, sprite_ids, nullptr, "par_update_entities");
}

// ---- frame ---------------------------------------------------------------------------------------------
// Which build of the render kernel a frame runs on (tile.cu): the one-light configuration (6 CTAs per SM, small
// lists) for one-light frames of at least 3 waves of CTAs (3840x2160: 7 waves, -7 %; its half on each of 2 GPUs,
// 3.5 waves: -6 %) — below that the longer latency of a CTA with 64 registers costs more than the sixth CTA
// hides (1920x1080 = 1.75 waves: +1 %; 480x320: +8 %).
// PAR_DEBUG_FLAGS 256: never, 512: always (the parity tests run many-light scenes on it).
static bool want_one_light_config(const par_ctx* c, int n_lights) {
    if (c->debug_flags & 256) return false;
    if (c->debug_flags & 512) return true;
    int first, rows;
    owned_tile_rows(c->d, first, rows);
    return n_lights == 1 && (long)rows * tiles_per_stripe(c->d) >= 3L * c->cta_slots;
}

static void fill_tile_params(par_ctx* c, TileParams& tp, const par_light* lights, int n_lights, uchar4* d_out) {
    const GridBuffers& g = c->gen[c->cur];
    memset(&tp, 0, sizeof tp);
    tp.d = c->d;
    tp.cnt = g.cnt;
    tp.ids = g.ids;
    tp.occ4 = g.occ4;
    tp.boxes = g.boxes;
    tp.atlas_depth = c->d_atlas_depth;
    tp.texel_tab = c->d_texel_tab;
    tp.sprite_dims = c->d_sprite_dims;
    tp.atlas_texels = c->atlas_texels;
    tp.out = d_out;
    tp.n_lights = n_lights;
    tp.ambient = c->ambient;
    tp.probe_x = tp.probe_y = -1;
    tp.debug_flags = c->debug_flags;
    tp.one_light_config = want_one_light_config(c, n_lights);
    tp.phase_cycles = c->d_phase_cycles;  // NULL unless par_debug_phase_timing enabled it
    tp.dbg_light = -1;
    for (int l = 0; l < n_lights; l++)
        tp.lights[l] = make_short4(lights[l].x, lights[l].y, lights[l].z, lights[l].radius);
}

static bool want_tile_order(const par_ctx* c, int n_lights) {
    if (c->debug_flags & 32) return false;
    if (c->debug_flags & 64) return true;
    if (c->cfg.tile_order != 0) return c->cfg.tile_order > 0;
    if (n_lights >= 2) return true;  // tile costs differ by an order of magnitude: always worth it
    // One light: costs vary mildly, so the order only pays through the tail of the last wave — when the owned tiles
    // make between one and three waves of resident CTAs (measured: -10 % at 1920x1080 = 1.75 waves, nothing at
    // 3840x2160 = 7 waves, a launch too many at 480x320 = 0.13 waves).
    int first, rows;
    owned_tile_rows(c->d, first, rows);
    const long tiles = (long)rows * tiles_per_stripe(c->d);
    return n_lights == 1 && tiles >= c->cta_slots && tiles <= 3L * c->cta_slots;
}

// Longest-tile-first order for the NEXT render kernel from the costs the previous one recorded.  The sort
// (one small block) runs on a side branch beside the scene loader, so it never sits on a frame's critical
// path; inside a graph capture the branch becomes a parallel node.  join_tile_order makes the main stream
// wait for it (render_impl does, right before the render kernel).
static int fork_tile_order(par_ctx* c, int n_lights) {
    if (!want_tile_order(c, n_lights) || !c->cost_valid || c->order_fresh || c->order_forked) return PAR_OK;
    PAR_CUDA(cudaEventRecord(c->ev_fork, c->stream));
    PAR_CUDA(cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
    PAR_CUDA(launch_tile_order(c->d_tile_cost, c->d_tile_order, c->d, c->aux_stream));
    PAR_CUDA(cudaEventRecord(c->ev_join, c->aux_stream));
    c->order_forked = true;
    return PAR_OK;
}

static int join_tile_order(par_ctx* c) {
    if (!c->order_forked) return PAR_OK;
    PAR_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    c->order_forked = false;
    c->order_valid = c->order_fresh = true;
    return PAR_OK;
}

// Launch the render kernel for the context's band into d_out.  With host_out the band is cut
// into up to kMaxChunks row chunks (whole tile rows) and the D2H copy of chunk k overlaps the
// rendering of chunk k+1 on a second stream — at 4K the 33 MB readback takes longer than the
// kernel, so the drop-in call is roughly max(render, copy) instead of their sum.
struct TileOrderIO {   // explicit tile cost / order buffers of a frame (the overlapped resident frame), instead of the
    unsigned* cost;    // context's generic ones
    const int* order;
};

static int render_impl(par_ctx* c, const par_light* lights, int n_lights, uchar4* d_out, par_color* host_out,
                       bool striped_out = false, bool to_peers = false, bool pipelined = false, bool want_gbuf = false,
                       const TileOrderIO* io = nullptr) {
    if (c && c->slots_in_flight && !pipelined)
        return fail(PAR_ERR_STATE, "par_render: pipelined frames in flight, call par_wait_frame first%s%s");
    if (!c || n_lights < 0 || n_lights > PAR_MAX_LIGHTS || (n_lights > 0 && !lights))
        return fail(PAR_ERR_INVALID_ARG, "par_render: bad argument (at most 64 lights)%s%s");
    if (!c->scene_set || c->n_sprites == 0)
        return fail(PAR_ERR_STATE, "par_render: set the atlas and the scene first%s%s");
    DeviceGuard guard(c->cfg.device);
    const ViewDims& d = c->d;
    if (want_gbuf && !c->d_gbuf) {
        if (c->capturing) return fail(PAR_ERR_STATE, "par_render: internal: G-buffer must exist before a capture%s%s");
        PAR_CUDA(cudaMalloc(&c->d_gbuf, sizeof(int4) * (size_t)d.W * d.H));
        PAR_CUDA(cudaMemsetAsync(c->d_gbuf, 0, sizeof(int4) * (size_t)d.W * d.H, c->stream));
    }
    TileParams tp;
    fill_tile_params(c, tp, lights, n_lights, d_out);
    tp.gbuf = want_gbuf ? c->d_gbuf : nullptr;
    tp.out_stripe_T = striped_out ? (d.HH + d.stripe_n - 1) / d.stripe_n : 0;
    for (int r = 0; r < 8 && to_peers; r++)
        if (c->peer_frame[r] && (!c->exchange_on || c->exchange_root < 0 || r == c->exchange_root))
            tp.peer_out[tp.n_peer_out++] = c->peer_frame[r];
    if (c->cursor_x >= 0) {  // the record under the cursor goes to the host with every frame
        tp.probe_x = c->cursor_x;
        tp.probe_y = c->cursor_y;
        tp.probe_a = c->h_probe[0];
        tp.probe_b = pipelined ? *c->probe_slot : nullptr;
        if (!pipelined) c->probe_latest = &c->h_probe[0];
    }

    const int tile0 = d.row0 / kBin, tile1 = (d.row1 + kBin - 1) / kBin;
    int n_chunks = 1;
    if (host_out && d.stripe_n == 1 && tile1 - tile0 >= 4 * c->readback_chunks) {
        // Pipelining costs kernel efficiency (partial waves per chunk), so it is used only when
        // the readback is not small next to the kernels (measured on the previous frame) and
        // the destination is page-locked (a pageable copy would block the launching thread).
        cudaPointerAttributes attr;
        const bool pinned = cudaPointerGetAttributes(&attr, host_out) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        cudaGetLastError();
        const float copy_ms = (float)(d.row1 - d.row0) * d.W * 4.f / 50e6f;  // ~50 GB/s PCIe gen5
        if (pinned && c->last_kernel_ms > 0.f && c->last_kernel_ms < 2.5f * copy_ms) n_chunks = c->readback_chunks;
    }
    const bool ordered = !io && n_chunks == 1 && want_tile_order(c, n_lights);
    int extra_launches = 0;
    {   // the CTA order of this frame: forked beside the loader by the frame-level calls, else computed here
        const bool was_forked = c->order_forked;
        int rc = join_tile_order(c);
        if (rc != PAR_OK) return rc;
        if (ordered && c->cost_valid && !c->order_fresh) {
            PAR_CUDA(launch_tile_order(c->d_tile_cost, c->d_tile_order, d, c->stream));
            c->order_valid = c->order_fresh = true;
            extra_launches = 1;
        } else if (was_forked) {
            extra_launches = 1;
        }
    }
    PAR_CUDA(record_timing(c, c->ev_f0));
    for (int k = 0; k < n_chunks; k++) {
        const int ta = tile0 + (tile1 - tile0) * k / n_chunks, tb = tile0 + (tile1 - tile0) * (k + 1) / n_chunks;
        const int ra = ta * kBin > d.row0 ? ta * kBin : d.row0, rb = tb * kBin < d.row1 ? tb * kBin : d.row1;
        tp.d.row0 = ra;
        tp.d.row1 = rb;
        int first_owned, n_owned;
        owned_tile_rows(tp.d, first_owned, n_owned);
        tp.tile_row_first = first_owned;
        tp.tile_cost = io ? io->cost : ordered ? c->d_tile_cost : nullptr;
        tp.tile_order = io ? io->order : ordered && c->order_valid ? c->d_tile_order : nullptr;
        PAR_CUDA(record_timing(c, c->ev_chunk[k][0]));
        PAR_CUDA(launch_tile(tp, c->stream));
        PAR_CUDA(record_timing(c, c->ev_chunk[k][1]));
        if (host_out) {  // rows of this chunk the context owns -> their place in the host frame
            cudaStream_t cs = n_chunks > 1 ? c->copy_stream : c->stream;
            if (n_chunks > 1) PAR_CUDA(cudaStreamWaitEvent(cs, c->ev_chunk[k][1], 0));
            const size_t row_bytes = sizeof(par_color) * (size_t)d.W, pitch = host_pitch(c);
            const int seg = stripe_segments(d);
            const size_t seg_bytes = row_bytes / (size_t)seg;  // bytes of a stripe's rows (the whole row unless stripe_s > 1)
            for (int v = first_owned, q = 0; q < (d.stripe_n == 1 ? 1 : n_owned); q++, v += d.stripe_n) {
                const int t = v / seg;
                const int sa = d.stripe_n == 1 ? ra : std::max(t * kBin, ra), sb = d.stripe_n == 1 ? rb : std::min((t + 1) * kBin, rb);
                if (sb <= sa) continue;
                if (seg > 1) {  // a stripe that is part of a tile row: its columns only
                    const size_t col = (size_t)stripe_column_segment(d, v) * seg_bytes;
                    PAR_CUDA(cudaMemcpy2DAsync(reinterpret_cast<char*>(host_out) + (size_t)sa * pitch + col, pitch,
                                               reinterpret_cast<const char*>(d_out) + (size_t)sa * row_bytes + col, row_bytes,
                                               seg_bytes, (size_t)(sb - sa), cudaMemcpyDeviceToHost, cs));
                } else if (pitch == row_bytes)  // packed rows: a plain linear copy
                    PAR_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(host_out) + (size_t)sa * pitch,
                                             reinterpret_cast<const char*>(d_out) + (size_t)sa * row_bytes,
                                             row_bytes * (size_t)(sb - sa), cudaMemcpyDeviceToHost, cs));
                else
                    PAR_CUDA(cudaMemcpy2DAsync(reinterpret_cast<char*>(host_out) + (size_t)sa * pitch, pitch,
                                               reinterpret_cast<const char*>(d_out) + (size_t)sa * row_bytes, row_bytes,
                                               row_bytes, (size_t)(sb - sa), cudaMemcpyDeviceToHost, cs));
            }
        }
    }
    PAR_CUDA(record_timing(c, c->ev_f2));
    if (ordered) {  // this frame recorded its tile costs: the next frame sorts its CTAs by them
        c->cost_valid = true;
        c->order_fresh = false;
    } else if (!io) {
        c->cost_valid = c->order_valid = c->order_fresh = false;
    }
    if (host_out && n_chunks > 1) {  // make the context's stream cover the copies too
        PAR_CUDA(cudaEventRecord(c->ev_copy, c->copy_stream));
        PAR_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
    }
    c->n_chunks = n_chunks;
    c->launches_frame = n_chunks + extra_launches;
    c->last_n_lights = n_lights;
    c->frame_valid = true;
    c->gbuf_valid = want_gbuf;
    c->frame_timed = !c->capturing;
    return PAR_OK;
}

int par_render_device(par_ctx* c, const par_light* lights, int n_lights, void* d_rgba) {
    NvtxRange nvtx("par_render_device");
    return render_impl(c, lights, n_lights, d_rgba ? static_cast<uchar4*>(d_rgba) : (c ? c->d_frame : nullptr),
                       nullptr);
}

size_t par_staging_bytes(const par_ctx* c) {
    if (!c) return 0;
    const size_t T = (size_t)(c->d.HH + c->d.stripe_n - 1) / c->d.stripe_n;
    return (size_t)c->d.stripe_n * T * kBin * c->d.W * sizeof(par_color);
}

int par_render_device_striped(par_ctx* c, const par_light* lights, int n_lights, void* d_staging) {
    if (!c || !d_staging) return fail(PAR_ERR_INVALID_ARG, "par_render_device_striped: null argument%s%s");
    if (c->d.row0 != 0 || c->d.row1 != c->d.H || c->d.stripe_s > 1)
        return fail(PAR_ERR_INVALID_ARG,
                    "par_render_device_striped: the context must cover the whole frame with whole-row stripes (stripe_split <= 1)%s%s");
    return render_impl(c, lights, n_lights, static_cast<uchar4*>(d_staging), nullptr, true);
}

// ---- fused frame exchange over peer memory -------------------------------------------------------
int par_peer_export(par_ctx* c, void* handle64) {
    if (!c || !handle64) return fail(PAR_ERR_INVALID_ARG, "par_peer_export: null argument%s%s");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard guard(c->cfg.device);
    cudaIpcMemHandle_t h;
    PAR_CUDA(cudaIpcGetMemHandle(&h, c->d_frame_block));
    memcpy(handle64, &h, sizeof h);
    return PAR_OK;
}

static int set_peer(par_ctx* c, int rank, uchar4* ptr, bool ipc) {
    if (rank < 0 || rank >= 8 || rank == c->d.stripe_i)
        return fail(PAR_ERR_INVALID_ARG, "par_peer: rank must be another stripe index below 8%s%s");
    if (c->peer_frame[rank] && c->peer_is_ipc[rank]) cudaIpcCloseMemHandle(c->peer_frame[rank]);
    if (!c->peer_frame[rank]) c->n_peers++;
    c->peer_frame[rank] = ptr;
    c->peer_is_ipc[rank] = ipc;
    invalidate_graphs(c);
    return PAR_OK;
}

int par_peer_import(par_ctx* c, int rank, const void* handle64) {
    if (!c || !handle64) return fail(PAR_ERR_INVALID_ARG, "par_peer_import: null argument%s%s");
    DeviceGuard guard(c->cfg.device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    void* ptr = nullptr;
    PAR_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return set_peer(c, rank, static_cast<uchar4*>(ptr), true);
}

int par_peer_set(par_ctx* c, int rank, void* d_peer_frame) {
    if (!c || !d_peer_frame) return fail(PAR_ERR_INVALID_ARG, "par_peer_set: null argument%s%s");
    DeviceGuard guard(c->cfg.device);
    cudaPointerAttributes attr;
    PAR_CUDA(cudaPointerGetAttributes(&attr, d_peer_frame));
    if (attr.type != cudaMemoryTypeDevice)
        return fail(PAR_ERR_INVALID_ARG, "par_peer_set: not a device pointer%s%s");
    if (attr.device != c->cfg.device) {
        int can = 0;
        PAR_CUDA(cudaDeviceCanAccessPeer(&can, c->cfg.device, attr.device));
        if (!can) return fail(PAR_ERR_NO_DEVICE, "par_peer_set: no peer access between the two devices%s%s");
        cudaError_t e = cudaDeviceEnablePeerAccess(attr.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) PAR_CUDA(e);
        cudaGetLastError();
    }
    return set_peer(c, rank, static_cast<uchar4*>(d_peer_frame), false);
}

int par_exchange_setup(par_ctx* c, int root) {
    if (!c || root < -1 || root >= c->d.stripe_n)
        return fail(PAR_ERR_INVALID_ARG, "par_exchange_setup: root must be -1 or a rank below the stripe count%s%s");
    if (c->d.stripe_n < 2 || c->d.stripe_n > 8)
        return fail(PAR_ERR_INVALID_ARG, "par_exchange_setup: needs a striped context (2..8 ranks)%s%s");
    for (int r = 0; r < c->d.stripe_n; r++)
        if (r != c->d.stripe_i && !c->peer_frame[r])
            return fail(PAR_ERR_STATE, "par_exchange_setup: import every other rank first (par_peer_import / par_peer_set)%s%s");
    c->exchange_on = true;
    c->exchange_root = root;
    invalidate_graphs(c);
    return PAR_OK;
}

// Footer of rank r's frame as seen from this context.
static ExchangeFooter* footer_of(par_ctx* c, int r) {
    if (r == c->d.stripe_i) return c->d_footer;
    return reinterpret_cast<ExchangeFooter*>(reinterpret_cast<unsigned char*>(c->peer_frame[r]) + c->footer_offset);
}

// One frame from the resident scene on c->stream: [credit: release the previous frame] -> loader ->
// [wait for the consumers' credits] -> render kernel (+ peer stores) -> [signal arrival] ->
// [consumers: wait for all arrivals].
// Overlapped form (overlap == true): the frame renders from the current grid generation p — always the grid of the
// resident scene: every scene change rebuilds or patches it at once — while the side branch clears and rebuilds
// generation p ^ 1 from the same scene for the next frame (and sorts the next frame's tile order).  Every frame
// still runs one scene loader and one render kernel; the loader just no longer sits in front of the kernel.
static int enqueue_resident_frame(par_ctx* c, const par_light* lights, int n_lights, bool overlap) {
    const int gp = c->cur, gq = c->cur ^ 1;
    const bool ordered = overlap && want_tile_order(c, n_lights);
    int side_launches = 0;
    if (overlap) {
        PAR_CUDA(cudaEventRecord(c->ev_fork, c->stream));
        PAR_CUDA(cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
        if (ordered && c->res_cost_valid[gq]) {  // the next frame's CTA order, from the costs of the previous frame of its parity
            PAR_CUDA(launch_tile_order(c->d_res_cost[gq], c->d_res_order[gq], c->d, c->aux_stream));
            c->res_order_valid[gq] = true;
            side_launches++;
        }
        if (c->capturing || c->gen_dirty[gq])
            PAR_CUDA(launch_clear_touched(c->d, c->gen[gq], c->capturing ? c->cap_entities : c->list_bound[gq], c->aux_stream,
                                          &side_launches));
        LoaderParams lp;
        lp.d = c->d;
        lp.raw = c->d_raw;
        lp.sprite_ids = c->has_sprite_ids ? c->d_sprite_ids : nullptr;
        lp.sprite_dims = c->d_sprite_dims;
        lp.n = c->n_entities;
        lp.n_sprites = c->n_sprites;
        lp.cur = c->gen[gq];
        lp.old = GridBuffers{};  // nothing to clear in the same launch
        lp.old_n_list_cap = 0;
        lp.host_a = c->h_ctr;
        lp.host_b = nullptr;
        PAR_CUDA(launch_scene_loader(lp, c->aux_stream, &side_launches));
        PAR_CUDA(cudaEventRecord(c->ev_side, c->aux_stream));
    }
    const int me = c->d.stripe_i, n = c->d.stripe_n;
    const bool ex = c->exchange_on;
    const bool consumer = ex && (c->exchange_root < 0 || c->exchange_root == me);
    FlagTargets mine_credit{}, mine_arrive{}, their_arrive{}, their_credit{};
    if (ex) {
        for (int r = 0; r < n; r++) {
            if (r == me) continue;
            const bool r_consumes = c->exchange_root < 0 || c->exchange_root == r;
            if (consumer) {  // I consume: r's arrivals land in my footer, my credits go to r's footer
                mine_arrive.slot[mine_arrive.n++] = &c->d_footer->arrive[r];
                their_credit.slot[their_credit.n++] = &footer_of(c, r)->credit[me];
            }
            if (r_consumes) {  // r consumes what I produce
                their_arrive.slot[their_arrive.n++] = &footer_of(c, r)->arrive[me];
                mine_credit.slot[mine_credit.n++] = &c->d_footer->credit[r];
            }
        }
        // starting frame k releases frame k - 1: producers may overwrite my frame again
        if (their_credit.n) k_flag_signal<<<1, 32, 0, c->stream>>>(their_credit, c->d_seq + 1, 1u);
    }
    int rc = PAR_OK;
    if (!overlap) {
        if ((rc = fork_tile_order(c, n_lights)) != PAR_OK) return rc;  // beside the loader
        if ((rc = run_loader(c)) != PAR_OK) return rc;
    }
    // frame k (= produced + 1) may be stored once every consumer has released frame k - 1
    if (ex && mine_credit.n) k_flag_wait<<<1, 32, 0, c->stream>>>(mine_credit, c->d_seq + 0, 0u, 0, c->h_exchange_timeout);
    const TileOrderIO io{ordered ? c->d_res_cost[gp] : nullptr, ordered && c->res_order_valid[gp] ? c->d_res_order[gp] : nullptr};
    if ((rc = render_impl(c, lights, n_lights, c->d_frame, nullptr, false, ex, false, false, overlap ? &io : nullptr)) != PAR_OK)
        return rc;
    if (ex) {
        if (their_arrive.n) k_flag_signal<<<1, 32, 0, c->stream>>>(their_arrive, c->d_seq + 0, 0u);
        if (mine_arrive.n) k_flag_wait<<<1, 32, 0, c->stream>>>(mine_arrive, c->d_seq + 2, 1u, 1, c->h_exchange_timeout);
        PAR_CUDA(cudaGetLastError());
        c->launches_frame += (their_credit.n > 0) + (mine_credit.n > 0) + (their_arrive.n > 0) + (mine_arrive.n > 0);
    }
    if (overlap) {  // join the side branch; the rebuilt generation becomes the current one
        PAR_CUDA(cudaStreamWaitEvent(c->stream, c->ev_side, 0));
        c->res_cost_valid[gp] = ordered;
        c->gen_dirty[gq] = true;  // (gen[gp] stays dirty: it is cleared when it is the side branch's target, next frame)
        c->list_bound[gq] = c->n_entities;
        c->cur = gq;
        c->launches_build = side_launches;
        c->build_timed = false;
        c->gbuf_valid = false;
    }
    return PAR_OK;
}

int par_render_resident(par_ctx* c, const par_light* lights, int n_lights) {
    NvtxRange nvtx("par_render_resident");
    if (!c || n_lights < 0 || n_lights > PAR_MAX_LIGHTS || (n_lights > 0 && !lights))
        return fail(PAR_ERR_INVALID_ARG, "par_render_resident: bad argument (at most 64 lights)%s%s");
    if (!c->scene_set) return fail(PAR_ERR_STATE, "par_render_resident: no scene resident%s%s");
    if (c->slots_in_flight) return fail(PAR_ERR_STATE, "par_render_resident: pipelined frames in flight%s%s");
    DeviceGuard guard(c->cfg.device);
    const bool graph_ok = !(c->debug_flags & 16) && c->stream != nullptr && c->stream != cudaStreamLegacy &&
                          c->stream != cudaStreamPerThread && !c->d_phase_cycles;
    // Overlapped form (default): the scene loader rebuilds the other grid generation on a side branch while this
    // frame renders.  PAR_DEBUG_FLAGS & 128 (and streams that cannot fork) keep the loader in front of the kernel.
    const bool overlap = !(c->debug_flags & 128) && c->stream != nullptr && c->stream != cudaStreamLegacy &&
                         c->stream != cudaStreamPerThread;
    if (!graph_ok) return enqueue_resident_frame(c, lights, n_lights, overlap);
    // cached graphs are valid for one (lights, configuration) key and one grid generation each
    const bool same_key = c->resident_n_lights == n_lights &&
                          (n_lights == 0 || memcmp(c->resident_lights, lights, sizeof(par_light) * (size_t)n_lights) == 0);
    if (!same_key) {
        c->resident_ok[0] = c->resident_ok[1] = false;
        c->resident_n_lights = n_lights;
        if (n_lights) memcpy(c->resident_lights, lights, sizeof(par_light) * (size_t)n_lights);
    }
    const int parity = c->cur;  // the graph for "current generation = parity" builds into parity ^ 1
    const int nxt = parity ^ 1;
    const bool ordered = want_tile_order(c, n_lights);
    // A graph bakes in whether the frame uses a tile order and whether one is being sorted: capture only in the steady state.
    const bool steady = !ordered || (overlap ? c->res_order_valid[parity] && c->res_cost_valid[nxt] : c->cost_valid);
    if (c->resident_ok[parity] && c->resident_epoch[parity] == c->epoch && steady) {
        PAR_CUDA(cudaGraphLaunch(c->resident_exec[parity], c->stream));
        // host-side state the captured calls would have updated
        if (overlap) {
            c->gen_dirty[nxt] = true;  // (the generation rendered from stays dirty until it is the side branch's target)
        } else {
            if (ordered) {
                c->order_valid = true;
                c->order_fresh = false;
            }
            c->gen_dirty[parity] = false;
            c->list_bound[parity] = 0;
            c->gen_dirty[nxt] = true;
        }
        c->list_bound[nxt] = c->n_entities;
        c->cur = nxt;
        c->frame_valid = true;
        c->gbuf_valid = false;
        c->build_timed = c->frame_timed = false;
        c->last_n_lights = n_lights;
        return PAR_OK;
    }
    // the first ordered frames have no costs / order yet: their graph would bake "no order" in, so run them plainly
    if (!steady) return enqueue_resident_frame(c, lights, n_lights, overlap);
    PAR_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    c->capturing = true;
    int rc = enqueue_resident_frame(c, lights, n_lights, overlap);
    c->capturing = false;
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
    if (rc != PAR_OK) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        c->order_forked = false;  // (a side branch of the abandoned capture is gone with it)
        return rc;
    }
    PAR_CUDA(ce);
    if (c->resident_exec[parity]) {
        cudaGraphExecDestroy(c->resident_exec[parity]);
        c->resident_exec[parity] = nullptr;
    }
    ce = cudaGraphInstantiate(&c->resident_exec[parity], graph, 0);
    cudaGraphDestroy(graph);
    PAR_CUDA(ce);
    c->resident_ok[parity] = true;
    c->resident_epoch[parity] = c->epoch;
    PAR_CUDA(cudaGraphLaunch(c->resident_exec[parity], c->stream));
    return PAR_OK;
}

int par_set_output_pitch(par_ctx* c, size_t pitch_bytes) {
    if (!c) return fail(PAR_ERR_INVALID_ARG, "par_set_output_pitch: null context%s%s");
    if (pitch_bytes != 0 && (pitch_bytes < sizeof(par_color) * (size_t)c->d.W || pitch_bytes % 4 != 0))
        return fail(PAR_ERR_INVALID_ARG, "par_set_output_pitch: pitch must be 0 or a multiple of 4 that is >= width * 4%s%s");
    c->out_pitch = pitch_bytes;
    invalidate_graphs(c);
    return PAR_OK;
}

int par_read_frame_pitched(par_ctx* c, void* dst, size_t pitch_bytes) {
    if (!c || !dst) return fail(PAR_ERR_INVALID_ARG, "par_read_frame_pitched: null argument%s%s");
    const size_t row_bytes = sizeof(par_color) * (size_t)c->d.W;
    if (pitch_bytes < row_bytes) return fail(PAR_ERR_INVALID_ARG, "par_read_frame_pitched: pitch smaller than a row%s%s");
    DeviceGuard guard(c->cfg.device);
    if (pitch_bytes == row_bytes)
        PAR_CUDA(cudaMemcpyAsync(dst, c->d_frame, row_bytes * (size_t)c->d.H, cudaMemcpyDeviceToHost, c->stream));
    else
        PAR_CUDA(cudaMemcpy2DAsync(dst, pitch_bytes, c->d_frame, row_bytes, row_bytes, (size_t)c->d.H,
                                   cudaMemcpyDeviceToHost, c->stream));
    return PAR_OK;
}

int par_read_frame(par_ctx* c, par_color* out_rgba) {
    if (!c || !out_rgba) return fail(PAR_ERR_INVALID_ARG, "par_read_frame: null argument%s%s");
    return par_read_frame_pitched(c, out_rgba, host_pitch(c));
}

// D2H of the rows the context owns, from a raster frame in HBM into the same rows of a host frame
// (rows host_pitch apart).
static int enqueue_owned_rows_d2h(par_ctx* c, const uchar4* d_src, par_color* host_frame, cudaStream_t st) {
    const ViewDims& d = c->d;
    const size_t row_bytes = sizeof(par_color) * (size_t)d.W, pitch = host_pitch(c);
    int first, count;
    owned_tile_rows(d, first, count);
    if (count <= 0) return PAR_OK;
    const int n = d.stripe_n > 1 ? d.stripe_n : 1;
    const int last = first + (count - 1) * n;
    const int seg = stripe_segments(d);
    const bool whole_tiles = (first / seg) * kBin >= d.row0 && (last / seg + 1) * kBin <= d.row1;
    char* dst = reinterpret_cast<char*>(host_frame);
    const char* src = reinterpret_cast<const char*>(d_src);
    if (n == 1) {  // a band (or the whole frame): one block of rows — a plain linear copy when the rows are packed
        if (pitch == row_bytes)
            PAR_CUDA(cudaMemcpyAsync(dst + d.row0 * pitch, src + d.row0 * row_bytes, row_bytes * (size_t)(d.row1 - d.row0),
                                     cudaMemcpyDeviceToHost, st));
        else
            PAR_CUDA(cudaMemcpy2DAsync(dst + d.row0 * pitch, pitch, src + d.row0 * row_bytes, row_bytes, row_bytes,
                                       (size_t)(d.row1 - d.row0), cudaMemcpyDeviceToHost, st));
        return PAR_OK;
    }
    if (seg > 1) {  // stripes that are parts of tile rows: one 2-D copy each (its columns of its 40 rows)
        const size_t seg_bytes = row_bytes / (size_t)seg;
        if (whole_tiles && n % seg == 0) {
            // the column segment of the rank's stripes advances by one per stripe (stripe_column_segment: the
            // rotation every lcm(n, seg) = n stripes), so every seg-th of them has the same columns, n tile rows
            // apart: seg 3-D DMAs (x = the stripe's bytes of a row, y = 40 rows, z = the stripes, a slice = n
            // tile rows)
            const size_t slice_rows = (size_t)kBin * (size_t)n;
            for (int j = 0; j < seg && j < count; j++) {
                const int v = first + j * n;
                const size_t col = (size_t)stripe_column_segment(d, v) * seg_bytes, r0 = (size_t)(v / seg) * kBin;
                cudaMemcpy3DParms p3 = {};
                p3.srcPtr = make_cudaPitchedPtr(const_cast<char*>(src) + r0 * row_bytes + col, row_bytes, row_bytes, slice_rows);
                p3.dstPtr = make_cudaPitchedPtr(dst + r0 * pitch + col, pitch, row_bytes, slice_rows);
                p3.extent = make_cudaExtent(seg_bytes, kBin, (size_t)((count - j + seg - 1) / seg));
                p3.kind = cudaMemcpyDeviceToHost;
                PAR_CUDA(cudaMemcpy3DAsync(&p3, st));
            }
            return PAR_OK;
        }
        for (int v = first; v <= last; v += n) {
            const int t = v / seg;
            const int r0 = std::max(t * kBin, d.row0), r1 = std::min((t + 1) * kBin, d.row1);
            if (r1 <= r0) continue;
            const size_t col = (size_t)stripe_column_segment(d, v) * seg_bytes;
            PAR_CUDA(cudaMemcpy2DAsync(dst + r0 * pitch + col, pitch, src + r0 * row_bytes + col, row_bytes, seg_bytes,
                                       (size_t)(r1 - r0), cudaMemcpyDeviceToHost, st));
        }
        return PAR_OK;
    }
    if (whole_tiles && pitch == row_bytes) {  // the owned stripes lie at a regular pitch: one strided DMA
        const size_t stride = row_bytes * kBin * n, off = (size_t)first * kBin * row_bytes;
        PAR_CUDA(cudaMemcpy2DAsync(dst + off, stride, src + off, stride, row_bytes * kBin, count,
                                   cudaMemcpyDeviceToHost, st));
        return PAR_OK;
    }
    for (int t = first; t <= last; t += n) {  // band edges inside a tile row, or pitched rows: one copy per stripe
        const int r0 = std::max(t * kBin, d.row0), r1 = std::min((t + 1) * kBin, d.row1);
        if (r1 <= r0) continue;
        PAR_CUDA(cudaMemcpy2DAsync(dst + r0 * pitch, pitch, src + r0 * row_bytes, row_bytes, row_bytes,
                                   (size_t)(r1 - r0), cudaMemcpyDeviceToHost, st));
    }
    return PAR_OK;
}

int par_read_stripes(par_ctx* c, par_color* host_frame) {
    if (!c || !host_frame) return fail(PAR_ERR_INVALID_ARG, "par_read_stripes: null argument%s%s");
    DeviceGuard guard(c->cfg.device);
    return enqueue_owned_rows_d2h(c, c->d_frame, host_frame, c->stream);
}

// ---- pipelined frames ------------------------------------------------------------------------------
// Main stream:  [H2D scene k+1][loader][render kernel] ...      copy stream:  [D2H frame k]
// PCIe is full duplex and the copy engines run beside the SMs, so in steady state a frame costs
// max(D2H, H2D + kernels) instead of their sum.  Two slots: each has its own device frame (the
// copy of frame k reads slot k&1 while frame k+1 is shaded into the other) and its own copy of
// the loader's counters.  `update` selects par_submit_update: no upload, the resident scene is
// patched (first, n = count) instead.
static int submit_impl(par_ctx* c, bool update, int first, const par_aabb* aabbs, const int32_t* sprite_ids, int n,
                       const par_light* lights, int n_lights, par_color* out_rgba, const char* who) {
    if (!c || !out_rgba) return fail(PAR_ERR_INVALID_ARG, "%s: null argument%s", who);
    if (n < 0 || n > (1 << 26) || (n > 0 && !aabbs) || n_lights < 0 || n_lights > PAR_MAX_LIGHTS || (n_lights > 0 && !lights))
        return fail(PAR_ERR_INVALID_ARG, "%s: bad argument (at most 2^26 entities, 64 lights)%s", who);
    if (c->n_sprites == 0) return fail(PAR_ERR_STATE, "%s: call par_set_atlas first%s", who);
    if (update && !c->scene_set) return fail(PAR_ERR_STATE, "%s: no scene resident (call par_set_scene first)%s", who);
    if (update && (first < 0 || (long long)first + n > c->n_entities))
        return fail(PAR_ERR_INVALID_ARG, "%s: range exceeds the resident scene%s", who);
    if (c->slots_in_flight == 2)
        return fail(PAR_ERR_STATE, "%s: two frames in flight, call par_wait_frame first%s", who);
    DeviceGuard guard(c->cfg.device);
    const int slot = c->slot_next;
    if (slot == 1 && !c->d_frame_alt) {
        PAR_CUDA(cudaMalloc(&c->d_frame_alt, frame_bytes(c)));
        PAR_CUDA(cudaMemsetAsync(c->d_frame_alt, 0, frame_bytes(c), c->stream));
    }
    uchar4* d_out = slot ? c->d_frame_alt : c->d_frame;
    c->probe_slot = &c->h_probe[1 + slot];
    PAR_CUDA(cudaEventRecord(c->ev_slot_begin[slot], c->stream));
    // Upload, grid build and the render kernel go to the GPU as ONE graph launch: while the previous
    // frame's readback saturates PCIe, every separate launch costs ~25 us of command fetch latency
    // (measured: 0.36 ms of kernels stretch to 0.66 ms beside a running D2H).  The graph is
    // re-captured each frame (pointers, light values and grid generation change) and the executable
    // is updated in place.  Needs page-locked inputs (a pageable copy cannot be captured); an update
    // carries its records as kernel arguments, so it always qualifies unless it is a big one.
    bool use_graph = !(c->debug_flags & 16) && c->stream != nullptr && c->stream != cudaStreamLegacy &&
                     c->stream != cudaStreamPerThread;
    const bool big_update = update && (n > kMaxUpdate || c->list_bound[c->cur] + n > c->cap_entities);
    if (use_graph && (!update || big_update)) {
        cudaPointerAttributes attr;
        use_graph = n == 0 || (cudaPointerGetAttributes(&attr, aabbs) == cudaSuccess && attr.type == cudaMemoryTypeHost);
        if (use_graph && sprite_ids)
            use_graph = cudaPointerGetAttributes(&attr, sprite_ids) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        cudaGetLastError();
    }
    LoaderCounters* slot_ctr = &c->h_ctr[1 + slot];
    auto enqueue = [&]() -> int {
        int rc = fork_tile_order(c, n_lights);  // beside the upload and the loader
        if (rc == PAR_OK)
            rc = update ? update_entities_impl(c, first, n, aabbs, sprite_ids, slot_ctr, who)
                        : set_scene_impl(c, aabbs, sprite_ids, n, slot_ctr);
        if (rc == PAR_OK) rc = render_impl(c, lights, n_lights, d_out, nullptr, false, false, true);
        return rc;
    };
    int rc = PAR_OK;
    if (use_graph) {
        if (!update && (rc = reserve_entities(c, n)) != PAR_OK) return rc;  // may reallocate: not allowed inside a capture
        if (update && n > 0 && sprite_ids && !c->has_sprite_ids) {          // likewise: a memset of the whole id array
            PAR_CUDA(cudaMemsetAsync(c->d_sprite_ids, 0, sizeof(int) * (size_t)c->n_entities, c->stream));
            c->has_sprite_ids = true;
        }
        PAR_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        c->capturing = true;
        rc = enqueue();
        c->capturing = false;
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
        if (rc != PAR_OK) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            c->order_forked = false;  // (a side branch of the abandoned capture is gone with it)
            return rc;
        }
        PAR_CUDA(ce);
        if (c->frame_exec) {
            cudaGraphExecUpdateResultInfo info;
            if (cudaGraphExecUpdate(c->frame_exec, graph, &info) != cudaSuccess) {  // topology changed: rebuild
                cudaGetLastError();
                cudaGraphExecDestroy(c->frame_exec);
                c->frame_exec = nullptr;
            }
        }
        if (!c->frame_exec) {
            ce = cudaGraphInstantiate(&c->frame_exec, graph, 0);
            if (ce != cudaSuccess) {
                cudaGraphDestroy(graph);
                PAR_CUDA(ce);
            }
        }
        cudaGraphDestroy(graph);
        PAR_CUDA(cudaGraphLaunch(c->frame_exec, c->stream));
    } else if ((rc = enqueue()) != PAR_OK) {
        return rc;
    }
    PAR_CUDA(cudaEventRecord(c->ev_slot_kernels[slot], c->stream));
    PAR_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_slot_kernels[slot], 0));
    if ((rc = enqueue_owned_rows_d2h(c, d_out, out_rgba, c->copy_stream)) != PAR_OK) return rc;
    PAR_CUDA(cudaEventRecord(c->ev_slot_done[slot], c->copy_stream));
    c->slot_lights[slot] = n_lights;
    c->slot_next = slot ^ 1;
    c->slots_in_flight++;
    return PAR_OK;
}

int par_submit_frame(par_ctx* c, const par_aabb* aabbs, const int32_t* sprite_ids, int n, const par_light* lights,
                     int n_lights, par_color* out_rgba) {
    NvtxRange nvtx("par_submit_frame");
    return submit_impl(c, false, 0, aabbs, sprite_ids, n, lights, n_lights, out_rgba, "par_submit_frame");
}

int par_submit_update(par_ctx* c, int first, int count, const par_aabb* aabbs, const int32_t* sprite_ids,
                      const par_light* lights, int n_lights, par_color* out_rgba) {
    NvtxRange nvtx("par_submit_update");
    return submit_impl(c, true, first, aabbs, sprite_ids, count, lights, n_lights, out_rgba, "par_submit_update");
}

int par_wait_frame(par_ctx* c, par_stats* stats) {
    NvtxRange nvtx("par_wait_frame");
    if (!c) return fail(PAR_ERR_INVALID_ARG, "par_wait_frame: null context%s%s");
    if (c->slots_in_flight == 0) return fail(PAR_ERR_STATE, "par_wait_frame: no frame in flight%s%s");
    DeviceGuard guard(c->cfg.device);
    const int slot = c->slot_oldest;
    c->slot_oldest = slot ^ 1;
    c->slots_in_flight--;
    PAR_CUDA(cudaEventSynchronize(c->ev_slot_done[slot]));
    c->probe_latest = &c->h_probe[1 + slot];
    const LoaderCounters& lc = c->h_ctr[1 + slot];
    if (stats) {
        memset(stats, 0, sizeof *stats);
        PAR_CUDA(cudaEventElapsedTime(&stats->ms_total, c->ev_slot_begin[slot], c->ev_slot_done[slot]));
        PAR_CUDA(cudaEventElapsedTime(&stats->ms_readback, c->ev_slot_kernels[slot], c->ev_slot_done[slot]));
        stats->kernel_launches = c->launches_build + c->launches_frame;
        stats->n_entities = c->n_entities;
        stats->n_survivors = lc.n_survivors;
        stats->n_inserts = lc.n_inserts;
        stats->rays = owned_pixel_count(c) * (1 + (uint64_t)c->slot_lights[slot]);
    }
    return lc.bad_scene ? bad_scene_error(lc) : PAR_OK;
}

int par_set_cursor(par_ctx* c, int x, int y) {
    if (!c) return fail(PAR_ERR_INVALID_ARG, "par_set_cursor: null context%s%s");
    invalidate_graphs(c);
    if (x < 0 || y < 0) {
        c->cursor_x = c->cursor_y = -1;
        c->probe_latest = nullptr;
        return PAR_OK;
    }
    const ViewDims& d = c->d;
    const int n = d.stripe_n > 1 ? d.stripe_n : 1;
    const int stripe_of_pixel = stripe_of_segment(d, y / kBin, (x / kBin) / tiles_per_stripe(d));
    if (x >= d.W || y < d.row0 || y >= d.row1 || stripe_of_pixel % n != (n > 1 ? d.stripe_i : 0))
        return fail(PAR_ERR_INVALID_ARG, "par_set_cursor: the pixel is not one this context renders%s%s");
    c->cursor_x = x;
    c->cursor_y = y;
    c->probe_latest = nullptr;
    return PAR_OK;
}

int par_cursor_pixel(par_ctx* c, par_pixel* out) {
    if (!c || !out) return fail(PAR_ERR_INVALID_ARG, "par_cursor_pixel: null argument%s%s");
    if (!c->probe_latest)
        return fail(PAR_ERR_STATE, "par_cursor_pixel: no frame rendered since par_set_cursor%s%s");
    if (c->probe_latest == &c->h_probe[0]) {  // frame of a synchronous / device call: it may still be running
        DeviceGuard guard(c->cfg.device);
        PAR_CUDA(cudaStreamSynchronize(c->stream));
    }
    static_assert(sizeof(par_pixel) == sizeof(int[7]), "Pixel is 7 words");
    memcpy(out, *c->probe_latest, sizeof(par_pixel));
    return PAR_OK;
}

int par_register_host(void* p, size_t bytes) {
    if (!p || !bytes) return fail(PAR_ERR_INVALID_ARG, "par_register_host: null argument%s%s");
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(PAR_ERR_CUDA, "par_register_host: %s%s", cudaGetErrorString(e));
    }
    return PAR_OK;
}

int par_unregister_host(void* p) {
    if (p && cudaHostUnregister(p) != cudaSuccess) cudaGetLastError();
    return PAR_OK;
}

int par_render_device_peers(par_ctx* c, const par_light* lights, int n_lights) {
    if (!c) return fail(PAR_ERR_INVALID_ARG, "par_render_device_peers: null context%s%s");
    return render_impl(c, lights, n_lights, c->d_frame, nullptr, false, true);
}

int par_unstripe_device(par_ctx* c, const void* d_staging, void* d_rgba) {
    if (!c || !d_staging || !d_rgba) return fail(PAR_ERR_INVALID_ARG, "par_unstripe_device: null argument%s%s");
    DeviceGuard guard(c->cfg.device);
    const ViewDims& d = c->d;
    const int W4 = d.W / 4, T = (d.HH + d.stripe_n - 1) / d.stripe_n;
    const size_t n = (size_t)W4 * d.H;
    k_unstripe<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(static_cast<const uint4*>(d_staging),
                                                                    static_cast<uint4*>(d_rgba), W4, d.H, d.stripe_n, T);
    PAR_CUDA(cudaGetLastError());
    return PAR_OK;
}

void* par_device_frame(par_ctx* c) { return c ? c->d_frame : nullptr; }

// The parity checkpoints need the compact G-buffer in HBM; production frames do not write it.  When
// the latest frame was rendered without it, the primary-ray phase alone is run again (same grid,
// same atlas: same records).
static int ensure_gbuffer(par_ctx* c) {
    if (c->gbuf_valid) return PAR_OK;
    const ViewDims& d = c->d;
    if (!c->d_gbuf) {
        PAR_CUDA(cudaMalloc(&c->d_gbuf, sizeof(int4) * (size_t)d.W * d.H));
        PAR_CUDA(cudaMemsetAsync(c->d_gbuf, 0, sizeof(int4) * (size_t)d.W * d.H, c->stream));
    }
    TileParams tp;
    fill_tile_params(c, tp, nullptr, 0, c->d_frame);
    tp.gbuf = c->d_gbuf;
    tp.gbuf_only = 1;
    int first_owned, n_owned;
    owned_tile_rows(tp.d, first_owned, n_owned);
    tp.tile_row_first = first_owned;
    PAR_CUDA(launch_tile(tp, c->stream));
    c->gbuf_valid = true;
    return PAR_OK;
}

static int expand_gbuffer(par_ctx* c, par_pixel* gbuf, int32_t* texel) {
    const ViewDims& d = c->d;
    int rc = ensure_gbuffer(c);
    if (rc != PAR_OK) return rc;
    size_t px = (size_t)d.W * d.H;
    size_t first = (size_t)d.row0 * d.W, count = (size_t)(d.row1 - d.row0) * d.W;
    if (gbuf && !c->d_expanded) PAR_CUDA(cudaMalloc(&c->d_expanded, sizeof(int) * 7 * px));
    if (texel && !c->d_texel) PAR_CUDA(cudaMalloc(&c->d_texel, sizeof(int) * px));
    k_expand_gbuf<<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(
        c->d_gbuf, c->d_texel_tab, c->gen[c->cur].boxes, c->d_sprite_dims, first, count,
        gbuf ? c->d_expanded : nullptr, texel ? c->d_texel : nullptr);
    PAR_CUDA(cudaGetLastError());
    if (gbuf)
        PAR_CUDA(cudaMemcpyAsync(gbuf + first, c->d_expanded + first * 7, sizeof(par_pixel) * count,
                                 cudaMemcpyDeviceToHost, c->stream));
    if (texel)
        PAR_CUDA(cudaMemcpyAsync(texel + first, c->d_texel + first, sizeof(int) * count,
                                 cudaMemcpyDeviceToHost, c->stream));
    return PAR_OK;
}

int par_get_stats(par_ctx* c, par_stats* st) {
    if (!c || !st) return fail(PAR_ERR_INVALID_ARG, "par_get_stats: null argument%s%s");
    DeviceGuard guard(c->cfg.device);
    PAR_CUDA(cudaStreamSynchronize(c->stream));
    memset(st, 0, sizeof *st);
    if (c->build_timed) PAR_CUDA(cudaEventElapsedTime(&st->ms_grid_build, c->ev_build0, c->ev_build1));
    if (c->frame_timed) {
        for (int k = 0; k < c->n_chunks; k++) {
            float a = 0.f;
            PAR_CUDA(cudaEventElapsedTime(&a, c->ev_chunk[k][0], c->ev_chunk[k][1]));
            st->ms_render += a;
        }
        PAR_CUDA(cudaEventElapsedTime(&st->ms_total, c->ev_f0, c->ev_f2));
    }
    st->kernel_launches = c->launches_build + c->launches_frame;
    st->n_entities = c->n_entities;
    st->n_survivors = c->h_ctr->n_survivors;
    st->n_inserts = c->h_ctr->n_inserts;
    st->rays = owned_pixel_count(c) * (1 + (uint64_t)c->last_n_lights);
    return PAR_OK;
}

int par_render(par_ctx* c, const par_light* lights, int n_lights, par_color* out_rgba,
               par_pixel* out_gbuf, par_stats* stats) {
    NvtxRange nvtx("par_render");
    if (!c || !out_rgba) return fail(PAR_ERR_INVALID_ARG, "par_render: null argument%s%s");
    int rc = render_impl(c, lights, n_lights, c->d_frame, out_rgba, false, false, false, out_gbuf != nullptr);
    if (rc != PAR_OK) return rc;
    DeviceGuard guard(c->cfg.device);
    if (out_gbuf && (rc = expand_gbuffer(c, out_gbuf, nullptr)) != PAR_OK) return rc;
    PAR_CUDA(cudaStreamSynchronize(c->stream));
    if ((rc = check_scene_flag(c)) != PAR_OK) return rc;
    par_stats st;
    if ((rc = par_get_stats(c, &st)) != PAR_OK) return rc;
    c->last_kernel_ms = st.ms_render;
    if (stats) *stats = st;
    return PAR_OK;
}

int par_get_gbuffer(par_ctx* c, par_pixel* gbuf, int32_t* texel) {
    if (!c) return fail(PAR_ERR_INVALID_ARG, "par_get_gbuffer: null context%s%s");
    if (!c->frame_valid) return fail(PAR_ERR_STATE, "par_get_gbuffer: no frame rendered%s%s");
    DeviceGuard guard(c->cfg.device);
    int rc = expand_gbuffer(c, gbuf, texel);
    if (rc != PAR_OK) return rc;
    PAR_CUDA(cudaStreamSynchronize(c->stream));
    return PAR_OK;
}

int par_debug_intermediates(par_ctx* c, const par_light* lights, int n_lights, int light, float* t_lam,
                            float* factor) {
    if (!c || n_lights < 0 || n_lights > PAR_MAX_LIGHTS || (n_lights > 0 && !lights) || light < 0 || light >= std::max(n_lights, 1))
        return fail(PAR_ERR_INVALID_ARG, "par_debug_intermediates: bad argument%s%s");
    if (!c->frame_valid) return fail(PAR_ERR_STATE, "par_debug_intermediates: no frame rendered%s%s");
    DeviceGuard guard(c->cfg.device);
    const ViewDims& d = c->d;
    const size_t px = (size_t)d.W * d.H;
    float4* d_t = nullptr;
    float* d_f = nullptr;
    uchar4* d_scratch = nullptr;
    int rc = [&]() -> int {
        PAR_CUDA(cudaMalloc(&d_t, sizeof(float4) * px));
        PAR_CUDA(cudaMalloc(&d_f, sizeof(float) * px));
        PAR_CUDA(cudaMalloc(&d_scratch, sizeof(uchar4) * px));
        PAR_CUDA(cudaMemsetAsync(d_t, 0, sizeof(float4) * px, c->stream));
        PAR_CUDA(cudaMemsetAsync(d_f, 0, sizeof(float) * px, c->stream));
        TileParams tp;
        fill_tile_params(c, tp, lights, n_lights, d_scratch);
        tp.dbg_light = light;
        tp.dbg_t = d_t;
        tp.dbg_factor = d_f;
        int first_owned, n_owned;
        owned_tile_rows(tp.d, first_owned, n_owned);
        tp.tile_row_first = first_owned;
        PAR_CUDA(launch_tile(tp, c->stream));
        if (t_lam) PAR_CUDA(cudaMemcpyAsync(t_lam, d_t, sizeof(float4) * px, cudaMemcpyDeviceToHost, c->stream));
        if (factor) PAR_CUDA(cudaMemcpyAsync(factor, d_f, sizeof(float) * px, cudaMemcpyDeviceToHost, c->stream));
        PAR_CUDA(cudaStreamSynchronize(c->stream));
        return PAR_OK;
    }();
    cudaFree(d_t);
    cudaFree(d_f);
    cudaFree(d_scratch);
    return rc;
}

int par_grid_volume(const par_ctx* c) { return c ? c->d.V : 0; }

// Debug: barrier-to-barrier cycle totals of the render kernel's phases, summed over CTAs.  enable != 0
// switches the instrumentation on (and zeroes the counters); out (16 values) may be NULL.
int par_debug_phase_timing(par_ctx* c, int enable, uint64_t* out) {
    if (!c) return fail(PAR_ERR_INVALID_ARG, "par_debug_phase_timing: null context%s%s");
    DeviceGuard guard(c->cfg.device);
    PAR_CUDA(cudaStreamSynchronize(c->stream));
    invalidate_graphs(c);
    if (out) {
        memset(out, 0, 16 * sizeof(uint64_t));
        if (c->d_phase_cycles)
            PAR_CUDA(cudaMemcpy(out, c->d_phase_cycles, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    }
    if (enable) {
        if (!c->d_phase_cycles) PAR_CUDA(cudaMalloc(&c->d_phase_cycles, 16 * sizeof(uint64_t)));
        PAR_CUDA(cudaMemset(c->d_phase_cycles, 0, 16 * sizeof(uint64_t)));
    } else if (c->d_phase_cycles) {
        cudaFree(c->d_phase_cycles);
        c->d_phase_cycles = nullptr;
    }
    return PAR_OK;
}

int par_get_grid(par_ctx* c, int32_t* count, int32_t* ids) {
    if (!c) return fail(PAR_ERR_INVALID_ARG, "par_get_grid: null context%s%s");
    if (!c->scene_set) return fail(PAR_ERR_STATE, "par_get_grid: no scene resident%s%s");
    DeviceGuard guard(c->cfg.device);
    const GridBuffers& g = c->gen[c->cur];
    size_t V = (size_t)c->d.V;
    int* cnt = (int*)malloc(sizeof(int) * V);
    if (!cnt) return fail(PAR_ERR_OUT_OF_MEMORY, "par_get_grid: host allocation failed%s%s");
    cudaError_t e = cudaMemcpyAsync(cnt, g.cnt, sizeof(int) * V, cudaMemcpyDeviceToHost, c->stream);
    if (!e && ids)
        e = cudaMemcpyAsync(ids, g.ids, sizeof(int) * V * kSlots, cudaMemcpyDeviceToHost, c->stream);
    if (!e) e = cudaStreamSynchronize(c->stream);
    if (e) {
        free(cnt);
        return fail(PAR_ERR_CUDA, "par_get_grid: %s%s", cudaGetErrorString(e));
    }
    // Layout conversion only: device keeps insert totals and descending winners; the
    // reference's view is count = total & 7 and slots in ascending entity order (quirk Q2).
    for (size_t f = 0; f < V; f++) {
        int keep = cnt[f] & (kSlots - 1);
        if (ids) {
            int* row = ids + f * kSlots;
            for (int a = 0, b = keep - 1; a < b; a++, b--) {
                int t = row[a];
                row[a] = row[b];
                row[b] = t;
            }
            for (int s = keep; s < kSlots; s++) row[s] = -1;
        }
        if (count) count[f] = keep;
    }
    free(cnt);
    return check_scene_flag(c);
}

}  // extern "C"
