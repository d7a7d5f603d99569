// par_api.cu — the extern "C" layer of libpar_b200.so (include/par/par.h): context,
// device memory, streams/events, and the launch sequence that replaces the reference frame
// loop body /root/reference/src/alternative.cpp:689-760.  No compute happens on the host and
// there is no CPU fallback: without a CUDA device every entry point fails.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "par/par.h"
#include "par_kernels.cuh"

namespace par {

// Expand the compact G-buffer into the reference's Pixel records (sprites.hpp:53-58) and a
// texel-index plane, for the parity checkpoints of par_get_gbuffer / par_render(out_gbuf).
// Compact G-buffer record -> the reference's 28-byte Pixel (normal, colour, y, z, entity) + texel index.
__device__ __forceinline__ int expand_pixel(int4 g, const float* __restrict__ atlas_normal,
                                            const unsigned char* __restrict__ atlas_color,
                                            const uchar4* __restrict__ palette, int o[7]) {
    float nx = 0.f, ny = 0.f, nz = 0.f;
    uchar4 c = make_uchar4(127, 127, 127, 0);  // alternative.cpp:281
    int texel = -1;
    if (g.w >= 0) {
        int spr = g.w >> 10;
        texel = g.w & 1023;
        const float* nrm = atlas_normal + (spr * kTexels + texel) * 3;
        nx = nrm[0];
        ny = nrm[1];
        nz = nrm[2];
        c = palette[atlas_color[spr * kTexels + texel]];
    }
    o[0] = __float_as_int(nx);
    o[1] = __float_as_int(ny);
    o[2] = __float_as_int(nz);
    o[3] = (int)((unsigned)c.x | (unsigned)c.y << 8 | (unsigned)c.z << 16 | (unsigned)c.w << 24);
    o[4] = g.y;
    o[5] = g.z;
    o[6] = g.x;
    return texel;
}

__global__ void __launch_bounds__(256)
k_expand_gbuf(const int4* __restrict__ gbuf, const float* __restrict__ atlas_normal,
              const unsigned char* __restrict__ atlas_color, const uchar4* __restrict__ palette,
              size_t first, size_t count, int* __restrict__ out_pixel7, int* __restrict__ out_texel) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    size_t px = first + t;
    int o[7];
    const int texel = expand_pixel(gbuf[px], atlas_normal, atlas_color, palette, o);
    if (out_pixel7)
        for (int k = 0; k < 7; k++) out_pixel7[px * 7 + k] = o[k];
    if (out_texel) out_texel[px] = texel;
}

// The reference's cursor probe (mouse_pixel, alternative.cpp:380-382): the record of ONE pixel,
// stored straight into mapped pinned host memory (no copy-engine traffic, see k_publish_counters).
__global__ void k_probe_pixel(const int4* __restrict__ gbuf, const float* __restrict__ atlas_normal,
                              const unsigned char* __restrict__ atlas_color, const uchar4* __restrict__ palette,
                              size_t px, int* host_a, int* host_b) {
    int o[7];
    expand_pixel(gbuf[px], atlas_normal, atlas_color, palette, o);
    for (int k = 0; k < 7; k++) {
        if (host_a) host_a[k] = o[k];
        if (host_b) host_b[k] = o[k];
    }
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

struct par_ctx {
    par_config cfg;
    ViewDims d;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // D2H of finished row chunks, overlapping the next chunk
    cudaEvent_t ev_build0 = nullptr, ev_build1 = nullptr, ev_f0 = nullptr, ev_f2 = nullptr, ev_copy = nullptr;
    cudaEvent_t ev_chunk[kMaxChunks][4] = {};  // before primary / after primary / after walks / after shade
    int n_chunks = 1;
    // scene
    int4* d_raw = nullptr;
    int4* d_boxes = nullptr;
    int* d_sprite_ids = nullptr;
    int* d_survivors = nullptr;
    int cap_entities = 0, n_entities = 0;
    bool has_sprite_ids = false;
    // grid
    int* d_cnt = nullptr;
    int* d_ids = nullptr;
    unsigned* d_occ4 = nullptr;
    LoaderCounters* d_ctr = nullptr;
    LoaderCounters* h_ctr = nullptr;  // pinned
    // atlas
    int* d_atlas_depth = nullptr;
    float* d_atlas_normal = nullptr;
    unsigned char* d_atlas_color = nullptr;
    uchar4* d_palette = nullptr;
    int n_sprites = 0, n_palette = 0;
    // shadow-walk work descriptors and results (primary -> walks -> shade)
    int* d_tile_ngroups = nullptr;
    GroupMeta* d_groups = nullptr;
    int2* d_table = nullptr;
    int table_lights = 0;  // lights the table is sized for
    int4* d_pool = nullptr;
    int pool_cap = 0;
    int* d_pool_cursor = nullptr;
    // frame
    int4* d_gbuf = nullptr;
    uchar4* d_frame = nullptr;
    int* d_expanded = nullptr;  // W*H*7 ints, lazily allocated
    int* d_texel = nullptr;     // W*H ints, lazily allocated
    unsigned long long* d_phase_cycles = nullptr;  // debug only
    bool scene_set = false, frame_valid = false, build_timed = false, frame_timed = false;
    int launches_build = 0, launches_frame = 0, last_n_lights = 0;
    float ambient = 0.25f;
    // fused multi-GPU frame exchange: raster frames of the other ranks, mapped into this process
    uchar4* peer_frame[8] = {};
    bool peer_is_ipc[8] = {};
    int n_peers = 0;  // entries of peer_frame in use (own rank's entry stays NULL)
    // pipelined frames (par_submit_frame / par_wait_frame): two slots, each with its own device frame
    uchar4* d_frame_alt = nullptr;  // slot 1's frame, allocated on first use (slot 0 uses d_frame)
    cudaEvent_t ev_slot_begin[2] = {}, ev_slot_kernels[2] = {}, ev_slot_done[2] = {};
    int slots_in_flight = 0, slot_oldest = 0, slot_next = 0, slot_lights[2] = {};
    // cursor probe: [0] frames of the synchronous calls, [1 + slot] pipelined frames (pinned, 7 ints each)
    int cursor_x = -1, cursor_y = -1;
    int (*h_probe)[7] = nullptr;
    int (*probe_slot)[7] = nullptr;    // where the frame being submitted also stores its probe
    int (*probe_latest)[7] = nullptr;  // record of the frame par_cursor_pixel reports
    bool capturing = false;            // a pipelined frame is being captured into frame_exec
    cudaGraphExec_t frame_exec = nullptr;  // upload + build + kernels of one pipelined frame as ONE graph launch
    float last_kernel_ms = 0.f;  // primary + shade of the previous par_render (pipelining heuristic)
    int readback_chunks = 3;  // row chunks of the pipelined readback (PAR_READBACK_CHUNKS overrides)
    int debug_flags = 0;  // from the PAR_DEBUG_FLAGS environment variable (developer A/B switches)
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

// Timing events of the synchronous calls.  While a pipelined frame is captured into a CUDA graph
// they are skipped: an event recorded by a graph node can no longer be queried from the host.
cudaError_t record_timing(par_ctx* c, cudaEvent_t ev) {
    return c->capturing ? cudaSuccess : cudaEventRecord(ev, c->stream);
}

int run_loader(par_ctx* c, LoaderCounters* slot_ctr = nullptr) {
    c->launches_build = 0;
    PAR_CUDA(record_timing(c, c->ev_build0));
    PAR_CUDA(launch_scene_loader(c->d_raw, c->has_sprite_ids ? c->d_sprite_ids : nullptr,
                                 c->n_entities, c->n_sprites, c->d, c->d_boxes, c->d_cnt, c->d_ids,
                                 c->d_occ4, c->d_survivors, c->d_ctr, c->h_ctr, slot_ctr, c->stream,
                                 &c->launches_build));
    PAR_CUDA(record_timing(c, c->ev_build1));
    c->build_timed = !c->capturing;
    c->scene_set = true;
    c->frame_valid = false;
    return PAR_OK;
}

int bad_scene_error() {
    return fail(PAR_ERR_BAD_SCENE,
                "scene has an AABB with extent.x outside [0,20], extent.y+extent.z outside "
                "[0,40] or a sprite id outside the atlas (would index outside the 20x40 "
                "sprite, alternative.cpp:330)%s%s");
}

int check_scene_flag(par_ctx* c) { return c->h_ctr->bad_scene ? bad_scene_error() : PAR_OK; }

// Image rows the context renders (its band, restricted to its stripes).
uint64_t owned_row_count(const par_ctx* c) {
    uint64_t rows = 0;
    int first, count;
    owned_tile_rows(c->d, first, count);
    for (int t = first, q = 0; q < count; q++, t += c->d.stripe_n) {
        const int a = t * kBin > c->d.row0 ? t * kBin : c->d.row0;
        const int b = (t + 1) * kBin < c->d.row1 ? (t + 1) * kBin : c->d.row1;
        rows += (uint64_t)(b - a);
    }
    return rows;
}

}  // namespace

extern "C" {

const char* par_last_error(void) { return g_err; }
const char* par_version(void) { return "par_b200 0.1 (sm_100a)"; }

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

    DeviceGuard guard(cfg->device);
    int rc = [&]() -> int {
        PAR_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        c->stream = c->own_stream;
        PAR_CUDA(cudaEventCreate(&c->ev_build0));
        PAR_CUDA(cudaEventCreate(&c->ev_build1));
        PAR_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        PAR_CUDA(cudaEventCreate(&c->ev_f0));
        PAR_CUDA(cudaEventCreate(&c->ev_f2));
        PAR_CUDA(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
        for (auto& ev : c->ev_chunk)
            for (cudaEvent_t& e : ev) PAR_CUDA(cudaEventCreate(&e));
        size_t px = (size_t)d.W * d.H;
        PAR_CUDA(cudaMalloc(&c->d_cnt, sizeof(int) * (size_t)d.V));
        PAR_CUDA(cudaMalloc(&c->d_ids, sizeof(int) * (size_t)d.V * kSlots));
        PAR_CUDA(cudaMalloc(&c->d_occ4, sizeof(unsigned) * (((size_t)d.V + 7) / 8)));
        PAR_CUDA(cudaMalloc(&c->d_ctr, sizeof(LoaderCounters)));
        PAR_CUDA(cudaMallocHost(&c->h_ctr, 3 * sizeof(LoaderCounters)));  // [0] latest build, [1 + slot] pipelined frames
        memset(c->h_ctr, 0, 3 * sizeof(LoaderCounters));
        PAR_CUDA(cudaMallocHost(&c->h_probe, 3 * sizeof(int[7])));
        memset(c->h_probe, 0, 3 * sizeof(int[7]));
        for (int k = 0; k < 2; k++) {
            PAR_CUDA(cudaEventCreate(&c->ev_slot_begin[k]));
            PAR_CUDA(cudaEventCreate(&c->ev_slot_kernels[k]));
            PAR_CUDA(cudaEventCreate(&c->ev_slot_done[k]));
        }
        const size_t tiles = (size_t)d.HW * d.HH;
        PAR_CUDA(cudaMalloc(&c->d_tile_ngroups, sizeof(int) * tiles));
        PAR_CUDA(cudaMalloc(&c->d_groups, sizeof(GroupMeta) * tiles * kMaxGroups));
        PAR_CUDA(cudaMalloc(&c->d_pool_cursor, sizeof(int)));
        PAR_CUDA(cudaMalloc(&c->d_gbuf, sizeof(int4) * px));
        PAR_CUDA(cudaMalloc(&c->d_frame, sizeof(uchar4) * px));
        PAR_CUDA(cudaMemsetAsync(c->d_frame, 0, sizeof(uchar4) * px, c->stream));
        PAR_CUDA(cudaMemsetAsync(c->d_gbuf, 0, sizeof(int4) * px, c->stream));
        PAR_CUDA(configure_primary(primary_smem_bytes(d, 1)));
        PAR_CUDA(configure_shade());
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
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    cudaFree(c->d_raw);
    cudaFree(c->d_boxes);
    cudaFree(c->d_sprite_ids);
    cudaFree(c->d_survivors);
    cudaFree(c->d_cnt);
    cudaFree(c->d_ids);
    cudaFree(c->d_occ4);
    cudaFree(c->d_ctr);
    if (c->h_ctr) cudaFreeHost(c->h_ctr);
    if (c->h_probe) cudaFreeHost(c->h_probe);
    cudaFree(c->d_atlas_depth);
    cudaFree(c->d_atlas_normal);
    cudaFree(c->d_atlas_color);
    cudaFree(c->d_palette);
    for (int r = 0; r < 8; r++)
        if (c->peer_frame[r] && c->peer_is_ipc[r]) cudaIpcCloseMemHandle(c->peer_frame[r]);
    cudaFree(c->d_tile_ngroups);
    cudaFree(c->d_groups);
    cudaFree(c->d_table);
    cudaFree(c->d_pool);
    cudaFree(c->d_pool_cursor);
    cudaFree(c->d_gbuf);
    cudaFree(c->d_frame);
    cudaFree(c->d_frame_alt);
    if (c->frame_exec) cudaGraphExecDestroy(c->frame_exec);
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

int par_set_atlas(par_ctx* c, const par_sprite* sprites, int n_sprites, const par_color* palette,
                  int n_palette) {
    if (!c || !sprites || !palette || n_sprites <= 0 || n_palette <= 0 || n_palette > 256 ||
        n_sprites > (1 << 20))
        return fail(PAR_ERR_INVALID_ARG, "par_set_atlas: bad argument%s%s");
    // Compact atlas: the 16000-byte reference Sprite becomes three dense tables.
    size_t nt = (size_t)n_sprites * kTexels;
    int* depth = (int*)malloc(nt * sizeof(int));
    float* normal = (float*)malloc(nt * 3 * sizeof(float));
    unsigned char* color = (unsigned char*)malloc(nt);
    if (!depth || !normal || !color) {
        free(depth);
        free(normal);
        free(color);
        return fail(PAR_ERR_OUT_OF_MEMORY, "par_set_atlas: host allocation failed%s%s");
    }
    bool ok = true;
    for (int s = 0; s < n_sprites && ok; s++)
        for (int t = 0; t < kTexels; t++) {
            int ci = sprites[s].color[t];
            if (ci < 0 || ci >= n_palette) {
                ok = false;
                break;
            }
            color[(size_t)s * kTexels + t] = (unsigned char)ci;
            depth[(size_t)s * kTexels + t] = sprites[s].depth[t];
            memcpy(&normal[((size_t)s * kTexels + t) * 3], sprites[s].normal[t], 3 * sizeof(float));
        }
    int rc = PAR_OK;
    if (!ok) {
        rc = fail(PAR_ERR_INVALID_ARG, "par_set_atlas: sprite colour index outside the palette%s%s");
    } else {
        DeviceGuard guard(c->cfg.device);
        rc = [&]() -> int {
            PAR_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(c->d_atlas_depth);
            cudaFree(c->d_atlas_normal);
            cudaFree(c->d_atlas_color);
            cudaFree(c->d_palette);
            c->d_atlas_depth = nullptr;
            c->d_atlas_normal = nullptr;
            c->d_atlas_color = nullptr;
            c->d_palette = nullptr;
            c->n_sprites = 0;
            PAR_CUDA(cudaMalloc(&c->d_atlas_depth, nt * sizeof(int)));
            PAR_CUDA(cudaMalloc(&c->d_atlas_normal, nt * 3 * sizeof(float)));
            PAR_CUDA(cudaMalloc(&c->d_atlas_color, nt));
            PAR_CUDA(cudaMalloc(&c->d_palette, sizeof(uchar4) * 256));
            PAR_CUDA(cudaMemcpy(c->d_atlas_depth, depth, nt * sizeof(int), cudaMemcpyHostToDevice));
            PAR_CUDA(cudaMemcpy(c->d_atlas_normal, normal, nt * 3 * sizeof(float), cudaMemcpyHostToDevice));
            PAR_CUDA(cudaMemcpy(c->d_atlas_color, color, nt, cudaMemcpyHostToDevice));
            PAR_CUDA(cudaMemset(c->d_palette, 0, sizeof(uchar4) * 256));
            PAR_CUDA(cudaMemcpy(c->d_palette, palette, sizeof(par_color) * n_palette, cudaMemcpyHostToDevice));
            PAR_CUDA(configure_primary(primary_smem_bytes(c->d, n_sprites)));
            c->n_sprites = n_sprites;
            c->n_palette = n_palette;
            c->frame_valid = false;
            return PAR_OK;
        }();
    }
    free(depth);
    free(normal);
    free(color);
    return rc;
}

// Room for n entities in the scene buffers (synchronises the stream when it has to reallocate).
static int reserve_entities(par_ctx* c, int n) {
    if (n <= c->cap_entities) return PAR_OK;
    PAR_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_raw);
    cudaFree(c->d_boxes);
    cudaFree(c->d_sprite_ids);
    cudaFree(c->d_survivors);
    c->d_raw = c->d_boxes = nullptr;
    c->d_sprite_ids = c->d_survivors = nullptr;
    c->cap_entities = 0;
    size_t cap = (size_t)n + (size_t)n / 8 + 64;
    PAR_CUDA(cudaMalloc(&c->d_raw, sizeof(int4) * cap));
    PAR_CUDA(cudaMalloc(&c->d_boxes, sizeof(int4) * cap));
    PAR_CUDA(cudaMalloc(&c->d_sprite_ids, sizeof(int) * cap));
    PAR_CUDA(cudaMalloc(&c->d_survivors, sizeof(int) * cap));
    c->cap_entities = (int)cap;
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
    return set_scene_impl(c, aabbs, sprite_ids, n, nullptr);
}

int par_rebuild_grid(par_ctx* c) {
    if (!c) return fail(PAR_ERR_INVALID_ARG, "par_rebuild_grid: null context%s%s");
    if (!c->scene_set) return fail(PAR_ERR_STATE, "par_rebuild_grid: no scene resident%s%s");
    DeviceGuard guard(c->cfg.device);
    return run_loader(c);
}

// Launch primary + shade for the context's band into d_out.  With host_out the band is cut
// into up to kMaxChunks row chunks (whole tile rows) and the D2H copy of chunk k overlaps the
// rendering of chunk k+1 on a second stream — at 4K the 33 MB readback takes about as long as
// the kernels, so the drop-in call is roughly max(render, copy) instead of their sum.
static int render_impl(par_ctx* c, const par_light* lights, int n_lights, uchar4* d_out, par_color* host_out,
                       bool striped_out = false, bool to_peers = false, bool pipelined = false) {
    if (c && c->slots_in_flight && !pipelined)
        return fail(PAR_ERR_STATE, "par_render: pipelined frames in flight, call par_wait_frame first%s%s");
    if (!c || n_lights < 0 || n_lights > PAR_MAX_LIGHTS || (n_lights > 0 && !lights))
        return fail(PAR_ERR_INVALID_ARG, "par_render: bad argument (at most 64 lights)%s%s");
    if (!c->scene_set || c->n_sprites == 0)
        return fail(PAR_ERR_STATE, "par_render: set the atlas and the scene first%s%s");
    DeviceGuard guard(c->cfg.device);
    const ViewDims& d = c->d;
    PrimaryParams pp;
    pp.d = d;
    pp.cnt = c->d_cnt;
    pp.ids = c->d_ids;
    pp.boxes = c->d_boxes;
    pp.atlas_depth = c->d_atlas_depth;
    pp.n_sprites = c->n_sprites;
    pp.gbuf = c->d_gbuf;
    // Shadow walks run as their own kernel between primary and shade (walks.cu); size its table
    // for this light count and its pool for ~128 kept boxes per (tile, light).
    // EXPERIMENTAL, off by default (PAR_DEBUG_FLAGS bit 3): the in-kernel walks of k_shade are the
    // product path; the separate kernel only pays off for many-light scenes once its per-thread
    // loads are batched.
    const bool use_walks = n_lights > 0 && (c->debug_flags & 8) && !(c->debug_flags & 4);
    if (use_walks) {
        const size_t tiles = (size_t)d.HW * d.HH;
        if (n_lights > c->table_lights) {
            PAR_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(c->d_table);
            c->d_table = nullptr;
            c->table_lights = 0;
            PAR_CUDA(cudaMalloc(&c->d_table, sizeof(int2) * tiles * kMaxGroups * n_lights));
            c->table_lights = n_lights;
        }
        const size_t want = tiles * (size_t)n_lights * 128;
        const int cap = (int)(want < (1u << 20) ? (1u << 20) : want > 0x7fffffffu / 2 ? 0x7fffffffu / 2 : want);
        if (cap > c->pool_cap) {
            PAR_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(c->d_pool);
            c->d_pool = nullptr;
            c->pool_cap = 0;
            PAR_CUDA(cudaMalloc(&c->d_pool, sizeof(int4) * (size_t)cap));
            c->pool_cap = cap;
        }
        PAR_CUDA(cudaMemsetAsync(c->d_pool_cursor, 0, sizeof(int), c->stream));
    }
    pp.tile_ngroups = use_walks ? c->d_tile_ngroups : nullptr;
    pp.groups = c->d_groups;
    WalkParams wp;
    wp.d = d;
    wp.ids = c->d_ids;
    wp.occ4 = c->d_occ4;
    wp.boxes = c->d_boxes;
    wp.tile_ngroups = c->d_tile_ngroups;
    wp.groups = c->d_groups;
    wp.table = c->d_table;
    wp.pool = c->d_pool;
    wp.pool_cursor = c->d_pool_cursor;
    wp.pool_cap = c->pool_cap;
    wp.n_lights = n_lights;
    wp.debug_flags = c->debug_flags;
    ShadeParams sp;
    sp.d = d;
    sp.cnt = c->d_cnt;
    sp.ids = c->d_ids;
    sp.occ4 = c->d_occ4;
    sp.boxes = c->d_boxes;
    sp.gbuf = c->d_gbuf;
    sp.atlas_normal = c->d_atlas_normal;
    sp.atlas_color = c->d_atlas_color;
    sp.palette = c->d_palette;
    sp.out = d_out;
    sp.n_lights = n_lights;
    sp.ambient = c->ambient;
    sp.phase_cycles = c->d_phase_cycles;  // NULL unless par_debug_phase_timing enabled it
    sp.debug_flags = c->debug_flags;
    sp.tile_ngroups = use_walks ? c->d_tile_ngroups : nullptr;
    sp.table = use_walks ? c->d_table : nullptr;
    sp.pool = c->d_pool;
    sp.out_stripe_T = striped_out ? (d.HH + d.stripe_n - 1) / d.stripe_n : 0;
    sp.n_peer_out = 0;
    for (int r = 0; r < 8 && to_peers; r++)
        if (c->peer_frame[r]) sp.peer_out[sp.n_peer_out++] = c->peer_frame[r];
    memset(sp.lights, 0, sizeof sp.lights);
    for (int l = 0; l < n_lights; l++)
        sp.lights[l] = make_short4(lights[l].x, lights[l].y, lights[l].z, lights[l].radius);
    memcpy(wp.lights, sp.lights, sizeof wp.lights);

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
    PAR_CUDA(record_timing(c, c->ev_f0));
    for (int k = 0; k < n_chunks; k++) {
        const int ta = tile0 + (tile1 - tile0) * k / n_chunks, tb = tile0 + (tile1 - tile0) * (k + 1) / n_chunks;
        const int ra = ta * kBin > d.row0 ? ta * kBin : d.row0, rb = tb * kBin < d.row1 ? tb * kBin : d.row1;
        pp.d.row0 = sp.d.row0 = wp.d.row0 = ra;
        pp.d.row1 = sp.d.row1 = wp.d.row1 = rb;
        int first_owned, n_owned;
        owned_tile_rows(pp.d, first_owned, n_owned);
        pp.tile_row_first = sp.tile_row_first = wp.tile_row_first = first_owned;
        PAR_CUDA(record_timing(c, c->ev_chunk[k][0]));
        PAR_CUDA(launch_primary(pp, c->stream));
        PAR_CUDA(record_timing(c, c->ev_chunk[k][1]));
        if (use_walks) PAR_CUDA(launch_walks(wp, c->stream));
        PAR_CUDA(record_timing(c, c->ev_chunk[k][2]));
        PAR_CUDA(launch_shade(sp, c->stream));
        PAR_CUDA(record_timing(c, c->ev_chunk[k][3]));
        if (host_out && d.stripe_n == 1) {
            const size_t first = (size_t)ra * d.W, count = (size_t)(rb - ra) * d.W;
            cudaStream_t cs = n_chunks > 1 ? c->copy_stream : c->stream;
            if (n_chunks > 1) PAR_CUDA(cudaStreamWaitEvent(cs, c->ev_chunk[k][3], 0));
            PAR_CUDA(cudaMemcpyAsync(host_out + first, d_out + first, sizeof(par_color) * count,
                                     cudaMemcpyDeviceToHost, cs));
        } else if (host_out) {  // only the owned stripes, each clipped to the band
            for (int t = first_owned, q = 0; q < n_owned; q++, t += d.stripe_n) {
                const int sa = t * kBin > ra ? t * kBin : ra, sb = (t + 1) * kBin < rb ? (t + 1) * kBin : rb;
                const size_t first = (size_t)sa * d.W, count = (size_t)(sb - sa) * d.W;
                PAR_CUDA(cudaMemcpyAsync(host_out + first, d_out + first, sizeof(par_color) * count,
                                         cudaMemcpyDeviceToHost, c->stream));
            }
        }
    }
    int probe_launch = 0;
    if (c->cursor_x >= 0) {  // the record under the cursor goes to the host with every frame
        k_probe_pixel<<<1, 1, 0, c->stream>>>(c->d_gbuf, c->d_atlas_normal, c->d_atlas_color, c->d_palette,
                                              (size_t)c->cursor_y * d.W + c->cursor_x, c->h_probe[0],
                                              pipelined ? *c->probe_slot : nullptr);
        PAR_CUDA(cudaGetLastError());
        probe_launch = 1;
        if (!pipelined) c->probe_latest = &c->h_probe[0];
    }
    PAR_CUDA(record_timing(c, c->ev_f2));
    if (host_out && n_chunks > 1) {  // make the context's stream cover the copies too
        PAR_CUDA(cudaEventRecord(c->ev_copy, c->copy_stream));
        PAR_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
    }
    c->n_chunks = n_chunks;
    c->launches_frame = (use_walks ? 3 : 2) * n_chunks + probe_launch;
    c->last_n_lights = n_lights;
    c->frame_valid = true;
    c->frame_timed = !c->capturing;
    return PAR_OK;
}

int par_render_device(par_ctx* c, const par_light* lights, int n_lights, void* d_rgba) {
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
    if (c->d.row0 != 0 || c->d.row1 != c->d.H)
        return fail(PAR_ERR_INVALID_ARG, "par_render_device_striped: the context must cover the whole frame%s%s");
    return render_impl(c, lights, n_lights, static_cast<uchar4*>(d_staging), nullptr, true);
}

// ---- fused frame exchange over peer memory -------------------------------------------------------
int par_peer_export(par_ctx* c, void* handle64) {
    if (!c || !handle64) return fail(PAR_ERR_INVALID_ARG, "par_peer_export: null argument%s%s");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard guard(c->cfg.device);
    cudaIpcMemHandle_t h;
    PAR_CUDA(cudaIpcGetMemHandle(&h, c->d_frame));
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

int par_read_frame(par_ctx* c, par_color* out_rgba) {
    if (!c || !out_rgba) return fail(PAR_ERR_INVALID_ARG, "par_read_frame: null argument%s%s");
    DeviceGuard guard(c->cfg.device);
    PAR_CUDA(cudaMemcpyAsync(out_rgba, c->d_frame, sizeof(par_color) * (size_t)c->d.W * c->d.H,
                             cudaMemcpyDeviceToHost, c->stream));
    return PAR_OK;
}

// D2H of the rows the context owns, from a raster frame in HBM into the same rows of a host frame.
static int enqueue_owned_rows_d2h(par_ctx* c, const uchar4* d_src, par_color* host_frame, cudaStream_t st) {
    const ViewDims& d = c->d;
    const size_t row_bytes = sizeof(par_color) * (size_t)d.W;
    int first, count;
    owned_tile_rows(d, first, count);
    if (count <= 0) return PAR_OK;
    const int n = d.stripe_n > 1 ? d.stripe_n : 1;
    const int last = first + (count - 1) * n;
    const bool whole_tiles = first * kBin >= d.row0 && (last + 1) * kBin <= d.row1;
    char* dst = reinterpret_cast<char*>(host_frame);
    const char* src = reinterpret_cast<const char*>(d_src);
    if (n == 1) {  // a band (or the whole frame): one contiguous block
        PAR_CUDA(cudaMemcpyAsync(dst + d.row0 * row_bytes, src + d.row0 * row_bytes,
                                 row_bytes * (size_t)(d.row1 - d.row0), cudaMemcpyDeviceToHost, st));
        return PAR_OK;
    }
    if (whole_tiles) {  // the owned stripes lie at a regular pitch: one strided DMA
        const size_t pitch = row_bytes * kBin * n, off = (size_t)first * kBin * row_bytes;
        PAR_CUDA(cudaMemcpy2DAsync(dst + off, pitch, src + off, pitch, row_bytes * kBin, count,
                                   cudaMemcpyDeviceToHost, st));
        return PAR_OK;
    }
    for (int t = first; t <= last; t += n) {  // band edges inside a tile row: one copy per clipped stripe
        const int r0 = std::max(t * kBin, d.row0), r1 = std::min((t + 1) * kBin, d.row1);
        if (r1 <= r0) continue;
        PAR_CUDA(cudaMemcpyAsync(dst + r0 * row_bytes, src + r0 * row_bytes, row_bytes * (size_t)(r1 - r0),
                                 cudaMemcpyDeviceToHost, st));
    }
    return PAR_OK;
}

int par_read_stripes(par_ctx* c, par_color* host_frame) {
    if (!c || !host_frame) return fail(PAR_ERR_INVALID_ARG, "par_read_stripes: null argument%s%s");
    DeviceGuard guard(c->cfg.device);
    return enqueue_owned_rows_d2h(c, c->d_frame, host_frame, c->stream);
}

// ---- pipelined frames ------------------------------------------------------------------------------
// Main stream:  [H2D scene k+1][loader][primary][shade] ...      copy stream:  [D2H frame k]
// PCIe is full duplex and the copy engines run beside the SMs, so in steady state a frame costs
// max(D2H, H2D + kernels) instead of their sum.  Two slots: each has its own device frame (the
// copy of frame k reads slot k&1 while frame k+1 is shaded into the other) and its own copy of
// the loader's counters.
int par_submit_frame(par_ctx* c, const par_aabb* aabbs, const int32_t* sprite_ids, int n, const par_light* lights,
                     int n_lights, par_color* out_rgba) {
    if (!c || !out_rgba) return fail(PAR_ERR_INVALID_ARG, "par_submit_frame: null argument%s%s");
    if (n < 0 || n > (1 << 26) || (n > 0 && !aabbs) || n_lights < 0 || n_lights > PAR_MAX_LIGHTS || (n_lights > 0 && !lights))
        return fail(PAR_ERR_INVALID_ARG, "par_submit_frame: bad argument (at most 2^26 entities, 64 lights)%s%s");
    if (c->n_sprites == 0) return fail(PAR_ERR_STATE, "par_submit_frame: call par_set_atlas first%s%s");
    if (c->slots_in_flight == 2)
        return fail(PAR_ERR_STATE, "par_submit_frame: two frames in flight, call par_wait_frame first%s%s");
    DeviceGuard guard(c->cfg.device);
    const int slot = c->slot_next;
    if (slot == 1 && !c->d_frame_alt) {
        const size_t px = (size_t)c->d.W * c->d.H;
        PAR_CUDA(cudaMalloc(&c->d_frame_alt, sizeof(uchar4) * px));
        PAR_CUDA(cudaMemsetAsync(c->d_frame_alt, 0, sizeof(uchar4) * px, c->stream));
    }
    uchar4* d_out = slot ? c->d_frame_alt : c->d_frame;
    c->probe_slot = &c->h_probe[1 + slot];
    PAR_CUDA(cudaEventRecord(c->ev_slot_begin[slot], c->stream));
    // Upload, grid build and both kernels go to the GPU as ONE graph launch: while the previous
    // frame's readback saturates PCIe, every separate launch costs ~25 us of command fetch latency
    // (measured: 0.36 ms of kernels stretch to 0.66 ms beside a running D2H).  The graph is
    // re-captured each frame (pointers, light values and grid sizes change) and the executable is
    // updated in place.  Needs page-locked inputs (a pageable copy cannot be captured).
    bool use_graph = !(c->debug_flags & (8 | 16)) && c->stream != nullptr && c->stream != cudaStreamLegacy &&
                     c->stream != cudaStreamPerThread;
    if (use_graph) {
        cudaPointerAttributes attr;
        use_graph = n == 0 || (cudaPointerGetAttributes(&attr, aabbs) == cudaSuccess && attr.type == cudaMemoryTypeHost);
        if (use_graph && sprite_ids)
            use_graph = cudaPointerGetAttributes(&attr, sprite_ids) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        cudaGetLastError();
    }
    int rc = PAR_OK;
    if (use_graph) {
        if ((rc = reserve_entities(c, n)) != PAR_OK) return rc;  // may reallocate: not allowed inside a capture
        PAR_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        c->capturing = true;
        rc = set_scene_impl(c, aabbs, sprite_ids, n, &c->h_ctr[1 + slot]);
        if (rc == PAR_OK) rc = render_impl(c, lights, n_lights, d_out, nullptr, false, false, true);
        c->capturing = false;
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
        if (rc != PAR_OK) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
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
    } else {
        rc = set_scene_impl(c, aabbs, sprite_ids, n, &c->h_ctr[1 + slot]);
        if (rc != PAR_OK) return rc;
        if ((rc = render_impl(c, lights, n_lights, d_out, nullptr, false, false, true)) != PAR_OK) return rc;
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

int par_wait_frame(par_ctx* c, par_stats* stats) {
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
        stats->rays = owned_row_count(c) * c->d.W * (1 + (uint64_t)c->slot_lights[slot]);
    }
    return lc.bad_scene ? bad_scene_error() : PAR_OK;
}

int par_set_cursor(par_ctx* c, int x, int y) {
    if (!c) return fail(PAR_ERR_INVALID_ARG, "par_set_cursor: null context%s%s");
    if (x < 0 || y < 0) {
        c->cursor_x = c->cursor_y = -1;
        c->probe_latest = nullptr;
        return PAR_OK;
    }
    const ViewDims& d = c->d;
    const int n = d.stripe_n > 1 ? d.stripe_n : 1;
    if (x >= d.W || y < d.row0 || y >= d.row1 || (y / kBin) % n != (n > 1 ? d.stripe_i : 0))
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

static int expand_gbuffer(par_ctx* c, par_pixel* gbuf, int32_t* texel) {
    const ViewDims& d = c->d;
    size_t px = (size_t)d.W * d.H;
    size_t first = (size_t)d.row0 * d.W, count = (size_t)(d.row1 - d.row0) * d.W;
    if (gbuf && !c->d_expanded) PAR_CUDA(cudaMalloc(&c->d_expanded, sizeof(int) * 7 * px));
    if (texel && !c->d_texel) PAR_CUDA(cudaMalloc(&c->d_texel, sizeof(int) * px));
    k_expand_gbuf<<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(
        c->d_gbuf, c->d_atlas_normal, c->d_atlas_color, c->d_palette, first, count,
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
            float a = 0.f, w = 0.f, b = 0.f;
            PAR_CUDA(cudaEventElapsedTime(&a, c->ev_chunk[k][0], c->ev_chunk[k][1]));
            PAR_CUDA(cudaEventElapsedTime(&w, c->ev_chunk[k][1], c->ev_chunk[k][2]));
            PAR_CUDA(cudaEventElapsedTime(&b, c->ev_chunk[k][2], c->ev_chunk[k][3]));
            st->ms_primary += a;
            st->ms_walks += w;
            st->ms_shade += b;
        }
        PAR_CUDA(cudaEventElapsedTime(&st->ms_total, c->ev_f0, c->ev_f2));
    }
    st->kernel_launches = c->launches_build + c->launches_frame;
    st->n_entities = c->n_entities;
    st->n_survivors = c->h_ctr->n_survivors;
    st->n_inserts = c->h_ctr->n_inserts;
    st->rays = owned_row_count(c) * c->d.W * (1 + (uint64_t)c->last_n_lights);
    return PAR_OK;
}

int par_render(par_ctx* c, const par_light* lights, int n_lights, par_color* out_rgba,
               par_pixel* out_gbuf, par_stats* stats) {
    if (!c || !out_rgba) return fail(PAR_ERR_INVALID_ARG, "par_render: null argument%s%s");
    int rc = render_impl(c, lights, n_lights, c->d_frame, out_rgba);
    if (rc != PAR_OK) return rc;
    DeviceGuard guard(c->cfg.device);
    if (out_gbuf && (rc = expand_gbuffer(c, out_gbuf, nullptr)) != PAR_OK) return rc;
    PAR_CUDA(cudaStreamSynchronize(c->stream));
    if ((rc = check_scene_flag(c)) != PAR_OK) return rc;
    par_stats st;
    if ((rc = par_get_stats(c, &st)) != PAR_OK) return rc;
    c->last_kernel_ms = st.ms_primary + st.ms_shade;
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

int par_grid_volume(const par_ctx* c) { return c ? c->d.V : 0; }

// Debug: barrier-to-barrier cycle totals of k_shade's phases, summed over CTAs.  enable != 0
// switches the instrumentation on (and zeroes the counters); out (16 values) may be NULL.
int par_debug_phase_timing(par_ctx* c, int enable, uint64_t* out) {
    if (!c) return fail(PAR_ERR_INVALID_ARG, "par_debug_phase_timing: null context%s%s");
    DeviceGuard guard(c->cfg.device);
    PAR_CUDA(cudaStreamSynchronize(c->stream));
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
    size_t V = (size_t)c->d.V;
    int* cnt = (int*)malloc(sizeof(int) * V);
    if (!cnt) return fail(PAR_ERR_OUT_OF_MEMORY, "par_get_grid: host allocation failed%s%s");
    cudaError_t e = cudaMemcpyAsync(cnt, c->d_cnt, sizeof(int) * V, cudaMemcpyDeviceToHost, c->stream);
    if (!e && ids)
        e = cudaMemcpyAsync(ids, c->d_ids, sizeof(int) * V * kSlots, cudaMemcpyDeviceToHost, c->stream);
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
