"""par_b200 — thin ctypes binding of libpar_b200.so (include/par/par.h).

The library is the product: hand-written sm_100a CUDA kernels behind a C ABI that replaces
the frame-loop body of Cons-Cat/Pixel-Art-Raytracer (/root/reference/src/alternative.cpp:689-760).
This module only marshals numpy arrays whose dtypes mirror the reference PODs.  There is no
CPU fallback: importing works anywhere (so symbols can be checked), but creating a Renderer
without the built library or without a B200 raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PAR_B200_LIB") or os.path.join(HERE, "libpar_b200.so")  # (override: A/B builds)

# numpy mirrors of the reference PODs (alternative.cpp:35-38, 619-622; sprites.hpp:5-6, 53-58, 67-71)
AABB = np.dtype([("px", "<i2"), ("py", "<i2"), ("pz", "<i2"), ("ex", "<i2"), ("ey", "<i2"),
                 ("ez", "<i2"), ("pad", "<i2", (2,))])
COLOR = np.dtype([("r", "u1"), ("g", "u1"), ("b", "u1"), ("a", "u1")])
PIXEL = np.dtype([("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"), ("color", COLOR), ("y", "<i4"),
                  ("z", "<i4"), ("entity", "<i4")])
LIGHT = np.dtype([("x", "<i2"), ("y", "<i2"), ("z", "<i2"), ("radius", "<i2")])
SPRITE = np.dtype([("color", "<i4", (800,)), ("depth", "<i4", (800,)),
                   ("normal", "<f4", (800, 3))])

PAR_OK = 0
STATUS_NAMES = {0: "PAR_OK", -1: "PAR_ERR_INVALID_ARG", -2: "PAR_ERR_NO_DEVICE",
                -3: "PAR_ERR_CUDA", -4: "PAR_ERR_OUT_OF_MEMORY", -5: "PAR_ERR_BAD_SCENE",
                -6: "PAR_ERR_STATE", -7: "PAR_ERR_NCCL"}

# every symbol include/par/par.h declares (checked by tests/test_abi.py)
EXPORTS = ["par_create", "par_destroy", "par_last_error", "par_version", "par_set_stream",
           "par_get_stream", "par_multi_create", "par_multi_destroy", "par_multi_size",
           "par_multi_context", "par_multi_set_atlas", "par_multi_set_scene", "par_multi_render",
           "par_multi_last_error", "par_sync", "par_alloc_host", "par_free_host", "par_set_atlas", "par_set_scene",
           "par_rebuild_grid", "par_render", "par_render_device", "par_device_frame",
           "par_render_device_striped", "par_staging_bytes", "par_unstripe_device",
           "par_peer_export", "par_peer_import", "par_peer_set", "par_render_device_peers", "par_read_frame",
           "par_read_stripes", "par_register_host", "par_unregister_host", "par_submit_frame", "par_wait_frame",
           "par_get_gbuffer", "par_get_grid", "par_get_stats", "par_grid_volume",
           "par_debug_phase_timing", "par_debug_intermediates", "par_set_atlas_sized", "par_update_entities",
           "par_submit_update", "par_set_output_pitch", "par_read_frame_pitched", "par_render_resident",
           "par_exchange_setup",
           "par_sprite_tile_floor", "par_palette_default", "par_scene_default",
           "par_light_default", "par_scene_synthetic", "par_apply_key", "par_draw_overlay",
           "par_draw_overlay_at", "par_set_cursor", "par_cursor_pixel", "par_fnv1a64"]


class Config(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("length", C.c_int32),
                ("device", C.c_int32), ("row_begin", C.c_int32), ("row_end", C.c_int32),
                ("ambient", C.c_float), ("stripe_count", C.c_int32), ("stripe_index", C.c_int32),
                ("tile_order", C.c_int32), ("stripe_split", C.c_int32), ("reserved", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("ms_grid_build", C.c_float), ("ms_render", C.c_float), ("ms_reserved", C.c_float),
                ("ms_total", C.c_float), ("kernel_launches", C.c_int32),
                ("n_entities", C.c_int32), ("n_survivors", C.c_int32), ("n_inserts", C.c_int32),
                ("rays", C.c_uint64), ("slab_tests", C.c_uint64), ("ms_reserved2", C.c_float),
                ("ms_readback", C.c_float), ("reserved", C.c_int32 * 2)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if "reserved" not in k}


class ParError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load libpar_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run pixel-art-raytracer_b200/build_native.sh "
                              "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, i32 = C.c_void_p, C.c_int
        L.par_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
        L.par_destroy.argtypes = [vp]
        L.par_destroy.restype = None
        L.par_last_error.restype = C.c_char_p
        L.par_version.restype = C.c_char_p
        L.par_set_stream.argtypes = [vp, vp]
        L.par_sync.argtypes = [vp]
        L.par_get_stream.argtypes = [vp]
        L.par_get_stream.restype = vp
        L.par_multi_create.argtypes = [C.POINTER(vp), C.POINTER(Config), C.POINTER(C.c_int), i32]
        L.par_multi_destroy.argtypes = [vp]
        L.par_multi_destroy.restype = None
        L.par_multi_size.argtypes = [vp]
        L.par_multi_context.argtypes = [vp, i32]
        L.par_multi_context.restype = vp
        L.par_multi_set_atlas.argtypes = [vp, vp, i32, vp, i32]
        L.par_multi_set_scene.argtypes = [vp, vp, vp, i32]
        L.par_multi_render.argtypes = [vp, vp, i32, vp, C.POINTER(Stats)]
        L.par_multi_last_error.restype = C.c_char_p
        L.par_alloc_host.argtypes = [C.c_size_t]
        L.par_alloc_host.restype = vp
        L.par_free_host.argtypes = [vp]
        L.par_free_host.restype = None
        L.par_set_atlas.argtypes = [vp, vp, i32, vp, i32]
        L.par_set_scene.argtypes = [vp, vp, vp, i32]
        L.par_rebuild_grid.argtypes = [vp]
        L.par_render.argtypes = [vp, vp, i32, vp, vp, C.POINTER(Stats)]
        L.par_render_device.argtypes = [vp, vp, i32, vp]
        L.par_render_device_striped.argtypes = [vp, vp, i32, vp]
        L.par_staging_bytes.argtypes = [vp]
        L.par_staging_bytes.restype = C.c_size_t
        L.par_unstripe_device.argtypes = [vp, vp, vp]
        L.par_peer_export.argtypes = [vp, vp]
        L.par_peer_import.argtypes = [vp, i32, vp]
        L.par_peer_set.argtypes = [vp, i32, vp]
        L.par_render_device_peers.argtypes = [vp, vp, i32]
        L.par_read_frame.argtypes = [vp, vp]
        L.par_read_stripes.argtypes = [vp, vp]
        L.par_submit_frame.argtypes = [vp, vp, vp, i32, vp, i32, vp]
        L.par_wait_frame.argtypes = [vp, vp]
        L.par_register_host.argtypes = [vp, C.c_size_t]
        L.par_unregister_host.argtypes = [vp]
        L.par_device_frame.argtypes = [vp]
        L.par_device_frame.restype = vp
        L.par_get_gbuffer.argtypes = [vp, vp, vp]
        L.par_get_grid.argtypes = [vp, vp, vp]
        L.par_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.par_grid_volume.argtypes = [vp]
        L.par_debug_phase_timing.argtypes = [vp, i32, vp]
        L.par_debug_intermediates.argtypes = [vp, vp, i32, i32, vp, vp]
        L.par_set_atlas_sized.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, i32]
        L.par_update_entities.argtypes = [vp, i32, i32, vp, vp]
        L.par_submit_update.argtypes = [vp, i32, i32, vp, vp, vp, i32, vp]
        L.par_set_output_pitch.argtypes = [vp, C.c_size_t]
        L.par_read_frame_pitched.argtypes = [vp, vp, C.c_size_t]
        L.par_render_resident.argtypes = [vp, vp, i32]
        L.par_exchange_setup.argtypes = [vp, i32]
        L.par_sprite_tile_floor.argtypes = [vp]
        L.par_sprite_tile_floor.restype = None
        L.par_palette_default.argtypes = [vp]
        L.par_palette_default.restype = None
        L.par_scene_default.argtypes = [vp, i32]
        L.par_light_default.argtypes = [vp]
        L.par_light_default.restype = None
        L.par_scene_synthetic.argtypes = [i32, i32, i32, C.c_uint64, i32, vp, i32, vp]
        L.par_scene_synthetic.restype = None
        L.par_apply_key.argtypes = [i32, vp, vp]
        L.par_apply_key.restype = None
        L.par_draw_overlay.argtypes = [i32, i32, vp, vp, i32, i32, vp]
        L.par_draw_overlay.restype = None
        L.par_draw_overlay_at.argtypes = [i32, i32, vp, vp, i32, vp]
        L.par_draw_overlay_at.restype = None
        L.par_fnv1a64.argtypes = [vp, C.c_size_t]
        L.par_fnv1a64.restype = C.c_uint64
        L.par_set_cursor.argtypes = [vp, i32, i32]
        L.par_cursor_pixel.argtypes = [vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _check(rc):
    if rc != PAR_OK:
        raise ParError(rc, lib().par_last_error().decode())


# ---- host-side pieces of the reference (inputs of the path) ---------------------------------

def tile_floor() -> np.ndarray:
    """make_tile_floor(), sprites.hpp:73-364."""
    s = np.zeros(1, SPRITE)
    lib().par_sprite_tile_floor(_p(s))
    return s


def default_palette() -> np.ndarray:
    p = np.zeros(4, COLOR)
    lib().par_palette_default(_p(p))
    return p


def scene_default() -> np.ndarray:
    """The 162 308-entity default scene, alternative.cpp:519-599."""
    n = lib().par_scene_default(None, 0)
    a = np.zeros(n, AABB)
    lib().par_scene_default(_p(a), n)
    return a


def light_default() -> np.ndarray:
    l = np.zeros(1, LIGHT)
    lib().par_light_default(_p(l))
    return l


def scene_synthetic(W, H, L, n=10000, n_lights=16, seed=0xB200):
    a = np.zeros(n, AABB)
    l = np.zeros(n_lights, LIGHT)
    lib().par_scene_synthetic(W, H, L, seed, n, _p(a), n_lights, _p(l))
    return a, l


def apply_key(key: str, boxes: np.ndarray, lights: np.ndarray) -> None:
    lib().par_apply_key(ord(key), _p(boxes), _p(lights))


def draw_overlay(W, H, gbuf, lights, frame, cx=0, cy=0) -> None:
    lib().par_draw_overlay(W, H, _p(gbuf), _p(lights), cx, cy, _p(frame))


def draw_overlay_at(W, H, under_cursor, lights, frame, cx=0) -> None:
    """Overlay from the single record under the cursor (Renderer.cursor_pixel())."""
    under = np.ascontiguousarray(under_cursor, PIXEL).reshape(1)
    lib().par_draw_overlay_at(W, H, _p(under), _p(lights), cx, _p(frame))


def fnv1a64(arr) -> int:
    """FNV-1a-64 of an array's bytes (the per-frame hash of the 240-frame goldens)."""
    a = np.ascontiguousarray(arr)
    return int(lib().par_fnv1a64(a.ctypes.data, a.nbytes))


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array over page-locked host memory from par_alloc_host (kept alive by the array)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    ptr = lib().par_alloc_host(max(n, 1))
    if not ptr:
        raise MemoryError(lib().par_last_error().decode())
    buf = (C.c_uint8 * n).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    _PINNED[arr.__array_interface__["data"][0]] = ptr
    return arr


_PINNED: dict[int, int] = {}


def shared_host_frame(path: str, H: int, W: int, create: bool, frames: int = 1) -> np.ndarray:
    """(H, W) COLOR frame — or (frames, H, W) of them — in a shared-memory file, page-locked in THIS
    process (par_register_host): several one-GPU processes DMA their stripes into the same host
    frame (par_read_stripes, par_submit_frame)."""
    n = frames * H * W * COLOR.itemsize
    if create:
        with open(path, "wb") as f:
            f.truncate(n)
    arr = np.memmap(path, dtype=COLOR, mode="r+", shape=(H, W) if frames == 1 else (frames, H, W))
    _check(lib().par_register_host(arr.ctypes.data, n))
    return arr


def release_shared_host_frame(arr: np.ndarray):
    lib().par_unregister_host(arr.ctypes.data)


# ---- the render path ------------------------------------------------------------------------

class Renderer:
    """One par_ctx: the device-resident state behind the reference's frame loop body.

    Mirrors the call shape of alternative.cpp:689-760:
        set_scene(aabbs)      <- memset + count_entities_in_bins
        render(lights)        <- trace_hash_for_pixel + shading loop, returns Color[H][W]
    """

    def __init__(self, W, H, L, device=0, row_begin=0, row_end=0, ambient=0.0, stripe_count=0,
                 stripe_index=0, tile_order=0, stripe_split=0):
        self.W, self.H, self.L = W, H, L
        self._h = C.c_void_p()
        cfg = Config(W, H, L, device, row_begin, row_end, ambient, stripe_count, stripe_index, tile_order, stripe_split)
        self.row_begin = row_begin
        self.row_end = row_end if (row_begin or row_end) else H
        _check(lib().par_create(C.byref(self._h), C.byref(cfg)))

    def close(self):
        if self._h:
            lib().par_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream: int | None):
        _check(lib().par_set_stream(self._h, cuda_stream))

    def sync(self):
        _check(lib().par_sync(self._h))

    def stream(self) -> int:
        """cudaStream_t (as an integer) all work of the context is ordered on."""
        return lib().par_get_stream(self._h)

    def set_atlas(self, sprites=None, palette=None):
        sprites = tile_floor() if sprites is None else np.ascontiguousarray(sprites, SPRITE)
        palette = default_palette() if palette is None else np.ascontiguousarray(palette, COLOR)
        _check(lib().par_set_atlas(self._h, _p(sprites), len(sprites), _p(palette), len(palette)))

    def set_atlas_sized(self, widths, heights, color, depth, normal, palette=None):
        """par_set_atlas_sized: sprites of their own width x height (tables concatenated, row-major)."""
        widths = np.ascontiguousarray(widths, np.int32)
        heights = np.ascontiguousarray(heights, np.int32)
        color = np.ascontiguousarray(color, np.int32)
        depth = np.ascontiguousarray(depth, np.int32)
        normal = np.ascontiguousarray(normal, np.float32)
        palette = default_palette() if palette is None else np.ascontiguousarray(palette, COLOR)
        _check(lib().par_set_atlas_sized(self._h, len(widths), _p(widths), _p(heights), _p(color), _p(depth),
                                         _p(normal), _p(palette), len(palette)))

    def set_scene(self, aabbs, sprite_ids=None):
        aabbs = np.ascontiguousarray(aabbs, AABB)
        if sprite_ids is not None:
            sprite_ids = np.ascontiguousarray(sprite_ids, np.int32)
        _check(lib().par_set_scene(self._h, _p(aabbs), _p(sprite_ids), len(aabbs)))

    def update_entities(self, first, aabbs, sprite_ids=None):
        """par_update_entities: entities [first, first + len(aabbs)) of the resident scene get new boxes."""
        aabbs = np.ascontiguousarray(aabbs, AABB)
        if sprite_ids is not None:
            sprite_ids = np.ascontiguousarray(sprite_ids, np.int32)
        _check(lib().par_update_entities(self._h, first, len(aabbs), _p(aabbs), _p(sprite_ids)))

    def submit_update(self, first, aabbs, lights, out, sprite_ids=None):
        """par_submit_update: pipelined frame from the resident scene with entities [first, ...) replaced."""
        aabbs = np.ascontiguousarray(aabbs, AABB)
        lights = np.ascontiguousarray(lights, LIGHT)
        if sprite_ids is not None:
            sprite_ids = np.ascontiguousarray(sprite_ids, np.int32)
        _check(lib().par_submit_update(self._h, first, len(aabbs), _p(aabbs), _p(sprite_ids), _p(lights),
                                       len(lights), _p(out)))

    def render_resident(self, lights):
        """par_render_resident: loader + render kernel (+ multi-GPU exchange) from the resident scene, async."""
        lights = np.ascontiguousarray(lights, LIGHT)
        _check(lib().par_render_resident(self._h, _p(lights), len(lights)))

    def exchange_setup(self, root=-1):
        _check(lib().par_exchange_setup(self._h, root))

    def set_output_pitch(self, pitch_bytes):
        _check(lib().par_set_output_pitch(self._h, pitch_bytes))

    def read_frame_pitched(self, out, pitch_bytes):
        """D2H of the whole frame into `out` (any array of >= H * pitch bytes), rows pitch_bytes apart; async."""
        _check(lib().par_read_frame_pitched(self._h, _p(out), pitch_bytes))
        return out

    def intermediates(self, lights, light):
        """par_debug_intermediates: (t_lam (H,W,4) float32, factor (H,W) float32) of the latest frame."""
        lights = np.ascontiguousarray(lights, LIGHT)
        t = np.zeros((self.H, self.W, 4), np.float32)
        f = np.zeros((self.H, self.W), np.float32)
        _check(lib().par_debug_intermediates(self._h, _p(lights), len(lights), light, _p(t), _p(f)))
        return t, f

    def rebuild_grid(self):
        _check(lib().par_rebuild_grid(self._h))

    def render(self, lights, out=None, want_gbuf=False):
        """par_render: returns rgba (H,W) COLOR [, gbuf (H,W) PIXEL], stats dict."""
        lights = np.ascontiguousarray(lights, LIGHT)
        rgba = np.zeros((self.H, self.W), COLOR) if out is None else out
        gbuf = np.zeros((self.H, self.W), PIXEL) if want_gbuf else None
        st = Stats()
        _check(lib().par_render(self._h, _p(lights), len(lights), _p(rgba), _p(gbuf), C.byref(st)))
        return (rgba, gbuf, st.as_dict()) if want_gbuf else (rgba, st.as_dict())

    def render_device(self, lights, d_rgba: int | None = None):
        """par_render_device: asynchronous, into HBM (own frame buffer when d_rgba is None)."""
        lights = np.ascontiguousarray(lights, LIGHT)
        _check(lib().par_render_device(self._h, _p(lights), len(lights), d_rgba))

    def render_device_striped(self, lights, d_staging: int):
        """This context's stripes, stripe-major, into a staging frame (see par.h)."""
        lights = np.ascontiguousarray(lights, LIGHT)
        _check(lib().par_render_device_striped(self._h, _p(lights), len(lights), d_staging))

    def staging_bytes(self) -> int:
        return int(lib().par_staging_bytes(self._h))

    def unstripe_device(self, d_staging: int, d_rgba: int):
        _check(lib().par_unstripe_device(self._h, d_staging, d_rgba))

    def peer_export(self) -> bytes:
        """64-byte CUDA IPC handle of this context's raster frame."""
        buf = C.create_string_buffer(64)
        _check(lib().par_peer_export(self._h, buf))
        return buf.raw

    def peer_import(self, rank: int, handle: bytes):
        _check(lib().par_peer_import(self._h, rank, C.create_string_buffer(handle, 64)))

    def peer_set(self, rank: int, d_frame: int):
        _check(lib().par_peer_set(self._h, rank, d_frame))

    def render_device_peers(self, lights):
        """Render own stripes into the own frame and, in place, into every imported peer frame."""
        lights = np.ascontiguousarray(lights, LIGHT)
        _check(lib().par_render_device_peers(self._h, _p(lights), len(lights)))

    def read_frame(self, out=None):
        """D2H of the whole raster frame, enqueued on the context's stream; call sync() before use."""
        out = np.zeros((self.H, self.W), COLOR) if out is None else out
        _check(lib().par_read_frame(self._h, _p(out)))
        return out

    def set_cursor(self, x: int, y: int):
        """Select the pixel whose G-buffer record every frame also delivers (mouse_pixel); x < 0: off."""
        _check(lib().par_set_cursor(self._h, x, y))

    def cursor_pixel(self) -> np.ndarray:
        """PIXEL record under the cursor of the frame most recently completed."""
        out = np.zeros(1, PIXEL)
        _check(lib().par_cursor_pixel(self._h, _p(out)))
        return out[0]

    def submit_frame(self, aabbs, lights, out, sprite_ids=None):
        """par_submit_frame: upload + build + render + readback of one frame, without waiting (<= 2 in
        flight).  `aabbs` and `out` should come from pinned_empty() and stay untouched until wait_frame()."""
        assert aabbs.dtype == AABB and aabbs.flags.c_contiguous
        assert out.shape == (self.H, self.W) and out.dtype == COLOR and out.flags.c_contiguous
        lights = np.ascontiguousarray(lights, LIGHT)
        if sprite_ids is not None:
            assert sprite_ids.dtype == np.int32 and sprite_ids.flags.c_contiguous
        _check(lib().par_submit_frame(self._h, _p(aabbs), _p(sprite_ids), len(aabbs), _p(lights), len(lights), _p(out)))

    def wait_frame(self):
        """par_wait_frame: block until the oldest submitted frame is complete in its `out`; returns its stats."""
        st = Stats()
        _check(lib().par_wait_frame(self._h, C.byref(st)))
        return st.as_dict()

    def read_stripes(self, host_frame):
        """D2H of the rows this context owns into their place in a full (H, W) host frame (async)."""
        assert host_frame.shape == (self.H, self.W) and host_frame.dtype == COLOR and host_frame.flags.c_contiguous
        _check(lib().par_read_stripes(self._h, _p(host_frame)))
        return host_frame

    def device_frame(self) -> int:
        return lib().par_device_frame(self._h)

    def gbuffer(self, want_texel=True):
        gbuf = np.zeros((self.H, self.W), PIXEL)
        texel = np.full((self.H, self.W), -1, np.int32) if want_texel else None
        _check(lib().par_get_gbuffer(self._h, _p(gbuf), _p(texel)))
        return gbuf, texel

    def grid(self):
        V = lib().par_grid_volume(self._h)
        count = np.zeros(V, np.int32)
        ids = np.zeros(V * 8, np.int32)
        _check(lib().par_get_grid(self._h, _p(count), _p(ids)))
        return count, ids.reshape(V, 8)

    PHASES = ["primary", "group", "setup", "walk", "gather", "shade", "tail", "p7", "p8"]

    def phase_timing(self, enable=True):
        """Debug: per-phase cycle totals of k_shade since the last call (dict), then (re)arm."""
        out = np.zeros(16, np.uint64)
        _check(lib().par_debug_phase_timing(self._h, int(enable), _p(out)))
        d = {n: int(out[i]) for i, n in enumerate(self.PHASES)}
        d.update(boxes_found=int(out[10]), boxes_kept=int(out[11]), rounds=int(out[12]),
                 retries_walk=int(out[13]), retries_gather=int(out[14]), retries_occ=int(out[15]), boxes_unique=int(out[9]))
        return d

    def stats(self):
        st = Stats()
        _check(lib().par_get_stats(self._h, C.byref(st)))
        return st.as_dict()


class MultiRenderer:
    """par_multi_*: one process, N GPUs, interleaved stripes (see par.h)."""

    def __init__(self, W, H, L, devices):
        self.W, self.H, self.L = W, H, L
        self._h = C.c_void_p()
        cfg = Config(W, H, L, 0, 0, 0, 0.0)
        devs = (C.c_int * len(devices))(*devices)
        rc = lib().par_multi_create(C.byref(self._h), C.byref(cfg), devs, len(devices))
        if rc != PAR_OK:
            raise ParError(rc, lib().par_multi_last_error().decode())

    def _check(self, rc):
        if rc != PAR_OK:
            raise ParError(rc, lib().par_multi_last_error().decode())

    def close(self):
        if self._h:
            lib().par_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_atlas(self, sprites=None, palette=None):
        sprites = tile_floor() if sprites is None else np.ascontiguousarray(sprites, SPRITE)
        palette = default_palette() if palette is None else np.ascontiguousarray(palette, COLOR)
        self._check(lib().par_multi_set_atlas(self._h, _p(sprites), len(sprites), _p(palette), len(palette)))

    def set_scene(self, aabbs, sprite_ids=None):
        aabbs = np.ascontiguousarray(aabbs, AABB)
        if sprite_ids is not None:
            sprite_ids = np.ascontiguousarray(sprite_ids, np.int32)
        self._check(lib().par_multi_set_scene(self._h, _p(aabbs), _p(sprite_ids), len(aabbs)))

    def render(self, lights, out=None):
        lights = np.ascontiguousarray(lights, LIGHT)
        rgba = np.zeros((self.H, self.W), COLOR) if out is None else out
        st = Stats()
        self._check(lib().par_multi_render(self._h, _p(lights), len(lights), _p(rgba), C.byref(st)))
        return rgba, st.as_dict()

    def render_resident(self, lights):
        """Device consumer: leave the finished frame in HBM, complete on every device."""
        lights = np.ascontiguousarray(lights, LIGHT)
        st = Stats()
        self._check(lib().par_multi_render(self._h, _p(lights), len(lights), None, C.byref(st)))
        return st.as_dict()

    def device_frame_copy(self, i) -> np.ndarray:
        """Synchronous copy of device i's raster frame (after render_resident: the whole frame)."""
        ctx = lib().par_multi_context(self._h, i)
        out = np.zeros((self.H, self.W), COLOR)
        _check(lib().par_read_frame(ctx, _p(out)))
        _check(lib().par_sync(ctx))
        return out
