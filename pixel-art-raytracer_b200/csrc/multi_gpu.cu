// multi_gpu.cu — single-process multi-GPU mode of the C ABI (par_multi_*): the image is split
// into interleaved 40-row stripes (tile row t belongs to device t % n — the per-row cost of the
// shadow walks varies strongly over the image, so contiguous bands scale badly); the scene/grid
// is replicated (a shadow ray may visit any bin).  Frame exchange:
//  * preferred — FUSED over peer memory: every device maps the raster frames of all others and
//    its shade kernel stores each finished 16-byte chunk into all of them (posted writes over
//    NVLink / NVSwitch); cross-device events make the frames complete — no pass over the frame;
//  * fallback (no peer access, or PAR_MULTI_EXCHANGE=nccl) — every device renders its stripes
//    STRIPE-MAJOR into a staging frame (its output is one contiguous block), one in-place
//    ncclAllGather completes the staging frames and a copy kernel turns them into raster frames.
//
// NCCL is loaded lazily with dlopen("libnccl.so.2") so that libpar_b200.so itself has no
// link-time NCCL dependency and the one-GPU path never touches it.  (bench.py's torchrun mode
// does the same exchange with torch.distributed's NCCL, one process per GPU.)
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "par/par.h"
#include "par_kernels.cuh"

namespace {

struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;

    bool load() {
        if (handle) return true;
        handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!handle) return false;
#define PAR_SYM(name) name = reinterpret_cast<decltype(name)>(dlsym(handle, "nccl" #name))
        PAR_SYM(CommInitAll);
        PAR_SYM(CommDestroy);
        PAR_SYM(AllGather);
        PAR_SYM(GroupStart);
        PAR_SYM(GroupEnd);
        PAR_SYM(GetErrorString);
#undef PAR_SYM
        return CommInitAll && CommDestroy && AllGather && GroupStart && GroupEnd && GetErrorString;
    }
};

Nccl g_nccl;
thread_local char g_multi_err[256] = "";

// Restores the caller's current device on every exit path.
struct DeviceScope {
    int prev = -1;
    DeviceScope() { cudaGetDevice(&prev); }
    ~DeviceScope() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace

struct par_multi {
    int n = 0;
    int W = 0, H = 0;
    par_ctx* ctx[8] = {};
    int device[8] = {};
    void* staging[8] = {};  // stripe-major frames, one per device (NCCL path only)
    bool peer = false;      // fused exchange: every device writes its stripes straight into all frames
    cudaEvent_t done[8] = {};
    ncclComm_t comm[8] = {};
    bool have_comm = false;
};

extern "C" {

const char* par_multi_last_error(void) { return g_multi_err[0] ? g_multi_err : par_last_error(); }

void par_multi_destroy(par_multi* m) {
    if (!m) return;
    for (int i = 0; i < m->n; i++) {
        if (m->have_comm && m->comm[i]) g_nccl.CommDestroy(m->comm[i]);
        if (m->done[i]) {
            int prev = 0;
            cudaGetDevice(&prev);
            cudaSetDevice(m->device[i]);
            cudaEventDestroy(m->done[i]);
            cudaSetDevice(prev);
        }
        if (m->staging[i]) {
            int prev = 0;
            cudaGetDevice(&prev);
            cudaSetDevice(m->device[i]);
            cudaFree(m->staging[i]);
            cudaSetDevice(prev);
        }
        par_destroy(m->ctx[i]);
    }
    delete m;
}

int par_multi_create(par_multi** out, const par_config* cfg, const int* devices, int n_devices) {
    g_multi_err[0] = 0;
    if (!out || !cfg || !devices || n_devices < 1 || n_devices > 8 || cfg->height < n_devices) {
        snprintf(g_multi_err, sizeof g_multi_err, "par_multi_create: bad argument (1..8 devices)");
        return PAR_ERR_INVALID_ARG;
    }
    *out = nullptr;
    par_multi* m = new (std::nothrow) par_multi;
    if (!m) return PAR_ERR_OUT_OF_MEMORY;
    m->n = n_devices;
    m->W = cfg->width;
    m->H = cfg->height;
    for (int i = 0; i < n_devices; i++) {
        par_config c = *cfg;
        c.device = devices[i];
        c.row_begin = c.row_end = 0;  // whole frame, of which this context owns every n-th tile row
        c.stripe_count = n_devices;
        c.stripe_index = i;
        m->device[i] = devices[i];
        int rc = par_create(&m->ctx[i], &c);
        if (rc != PAR_OK) {
            m->n = i + 1;
            par_multi_destroy(m);
            return rc;
        }
    }
    // Preferred exchange: peer memory.  Every device maps the raster frames of all others and its
    // shade kernel stores finished pixels into all of them (par_render_device_peers).
    m->peer = n_devices > 1;
    if (const char* e = getenv("PAR_MULTI_EXCHANGE"))  // "nccl" forces the all-gather path (tests, A/B)
        if (!strcmp(e, "nccl")) m->peer = false;
    for (int i = 0; i < n_devices && m->peer; i++)
        for (int j = 0; j < n_devices && m->peer; j++) {
            int can = 1;
            if (i != j && (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) != cudaSuccess || !can)) m->peer = false;
        }
    for (int i = 0; i < n_devices && m->peer; i++) {
        for (int j = 0; j < n_devices && m->peer; j++)
            if (i != j && par_peer_set(m->ctx[i], j, par_device_frame(m->ctx[j])) != PAR_OK) m->peer = false;
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(devices[i]);
        if (cudaEventCreateWithFlags(&m->done[i], cudaEventDisableTiming) != cudaSuccess) m->peer = false;
        cudaSetDevice(prev);
    }
    for (int i = 0; i < n_devices; i++) {
        int rc = PAR_OK;
        if (n_devices > 1 && !m->peer) {
            int prev = 0;
            cudaGetDevice(&prev);
            cudaSetDevice(devices[i]);
            if (cudaMalloc(&m->staging[i], par_staging_bytes(m->ctx[i])) != cudaSuccess) {
                cudaGetLastError();
                snprintf(g_multi_err, sizeof g_multi_err, "par_multi_create: staging frame allocation failed");
                rc = PAR_ERR_OUT_OF_MEMORY;
            }
            cudaSetDevice(prev);
        }
        if (rc != PAR_OK) {
            par_multi_destroy(m);
            return rc;
        }
    }
    if (n_devices > 1 && !m->peer) {
        if (!g_nccl.load()) {
            snprintf(g_multi_err, sizeof g_multi_err, "par_multi_create: cannot load libnccl.so.2: %s", dlerror());
            par_multi_destroy(m);
            return PAR_ERR_NCCL;
        }
        ncclResult_t r = g_nccl.CommInitAll(m->comm, n_devices, devices);
        if (r != ncclSuccess) {
            snprintf(g_multi_err, sizeof g_multi_err, "ncclCommInitAll: %s", g_nccl.GetErrorString(r));
            par_multi_destroy(m);
            return PAR_ERR_NCCL;
        }
        m->have_comm = true;
    }
    *out = m;
    return PAR_OK;
}

int par_multi_size(const par_multi* m) { return m ? m->n : 0; }
par_ctx* par_multi_context(par_multi* m, int i) { return (m && i >= 0 && i < m->n) ? m->ctx[i] : nullptr; }

// On an error every device is drained before returning, so that no device keeps work in flight on
// state the others never received (the caller re-sends atlas / scene after an error).
static int drain_and_return(par_multi* m, int rc) {
    for (int i = 0; i < m->n; i++) par_sync(m->ctx[i]);
    return rc;
}

int par_multi_set_atlas(par_multi* m, const par_sprite* sprites, int n_sprites, const par_color* palette,
                        int n_palette) {
    if (!m) return PAR_ERR_INVALID_ARG;
    for (int i = 0; i < m->n; i++) {
        int rc = par_set_atlas(m->ctx[i], sprites, n_sprites, palette, n_palette);
        if (rc != PAR_OK) return drain_and_return(m, rc);
    }
    return PAR_OK;
}

int par_multi_set_scene(par_multi* m, const par_aabb* aabbs, const int32_t* sprite_ids, int n) {
    if (!m) return PAR_ERR_INVALID_ARG;
    for (int i = 0; i < m->n; i++) {  // asynchronous per device: uploads and loaders overlap
        int rc = par_set_scene(m->ctx[i], aabbs, sprite_ids, n);
        if (rc != PAR_OK) return drain_and_return(m, rc);
    }
    return PAR_OK;
}

// Render all bands, gather, and (out_rgba != NULL) read the finished frame back from device 0.
int par_multi_render(par_multi* m, const par_light* lights, int n_lights, par_color* out_rgba, par_stats* stats) {
    g_multi_err[0] = 0;
    if (!m) return PAR_ERR_INVALID_ARG;
    DeviceScope scope;  // the caller's current device is restored on every exit
    if (m->n == 1) {
        int rc = par_render_device(m->ctx[0], lights, n_lights, nullptr);
        if (rc != PAR_OK) return rc;
        if (out_rgba && (rc = par_read_stripes(m->ctx[0], out_rgba)) != PAR_OK) return rc;
    } else if (out_rgba) {
        // host consumer: no GPU-to-GPU exchange at all — every device renders its stripes into its
        // own frame and ships them into the one host frame over its own PCIe link
        for (int i = 0; i < m->n; i++) {
            int rc = par_render_device(m->ctx[i], lights, n_lights, nullptr);
            if (rc == PAR_OK) rc = par_read_stripes(m->ctx[i], out_rgba);
            if (rc != PAR_OK) return drain_and_return(m, rc);
        }
    } else if (m->peer) {
        for (int i = 0; i < m->n; i++) {  // asynchronous: all devices render and scatter concurrently
            int rc = par_render_device_peers(m->ctx[i], lights, n_lights);
            if (rc != PAR_OK) return drain_and_return(m, rc);
            cudaError_t e = cudaSetDevice(m->device[i]);
            if (e == cudaSuccess) e = cudaEventRecord(m->done[i], static_cast<cudaStream_t>(par_get_stream(m->ctx[i])));
            if (e != cudaSuccess) {
                snprintf(g_multi_err, sizeof g_multi_err, "par_multi_render: %s", cudaGetErrorString(e));
                return drain_and_return(m, PAR_ERR_CUDA);
            }
        }
        for (int i = 0; i < m->n; i++) {  // a frame is complete once EVERY device has finished writing into it
            cudaError_t e = cudaSetDevice(m->device[i]);
            for (int j = 0; j < m->n && e == cudaSuccess; j++)
                if (j != i) e = cudaStreamWaitEvent(static_cast<cudaStream_t>(par_get_stream(m->ctx[i])), m->done[j], 0);
            if (e != cudaSuccess) {
                snprintf(g_multi_err, sizeof g_multi_err, "par_multi_render: %s", cudaGetErrorString(e));
                return drain_and_return(m, PAR_ERR_CUDA);
            }
        }
    } else {
        for (int i = 0; i < m->n; i++) {  // asynchronous: all devices render their stripes concurrently
            int rc = par_render_device_striped(m->ctx[i], lights, n_lights, m->staging[i]);
            if (rc != PAR_OK) return drain_and_return(m, rc);
        }
        const size_t block = par_staging_bytes(m->ctx[0]) / m->n;  // one rank's contiguous stripes
        ncclResult_t r = g_nccl.GroupStart();
        for (int i = 0; i < m->n && r == ncclSuccess; i++) {  // in place: rank i's block sits at offset i
            unsigned char* st = static_cast<unsigned char*>(m->staging[i]);
            r = g_nccl.AllGather(st + i * block, st, block, ncclUint8, m->comm[i],
                                 static_cast<cudaStream_t>(par_get_stream(m->ctx[i])));
        }
        ncclResult_t e = g_nccl.GroupEnd();
        if (r == ncclSuccess) r = e;
        if (r != ncclSuccess) {
            snprintf(g_multi_err, sizeof g_multi_err, "NCCL frame gather: %s", g_nccl.GetErrorString(r));
            return drain_and_return(m, PAR_ERR_NCCL);
        }
        for (int i = 0; i < m->n; i++) {  // staging -> raster frame on every device
            int rc = par_unstripe_device(m->ctx[i], m->staging[i], par_device_frame(m->ctx[i]));
            if (rc != PAR_OK) return rc;
        }
    }
    for (int i = 0; i < m->n; i++) {
        int rc = par_sync(m->ctx[i]);
        if (rc != PAR_OK) return rc;
    }
    if (stats) {
        int rc = par_get_stats(m->ctx[0], stats);
        if (rc != PAR_OK) return rc;
        for (int i = 1; i < m->n; i++) {  // the frame is done when the slowest band is
            par_stats s;
            if (par_get_stats(m->ctx[i], &s) != PAR_OK) continue;
            if (s.ms_total > stats->ms_total) stats->ms_total = s.ms_total;
            if (s.ms_render > stats->ms_render) stats->ms_render = s.ms_render;
            stats->kernel_launches += s.kernel_launches;
            stats->rays += s.rays;
        }
    }
    return PAR_OK;
}

}  // extern "C"
