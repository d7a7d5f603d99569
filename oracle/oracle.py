"""ctypes front end of the CPU oracle (oracle/oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (pixel-art-raytracer_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")

# numpy mirrors of the reference PODs (alternative.cpp:35-38, sprites.hpp:5-6,53-58,67-71)
AABB = np.dtype([("px", "<i2"), ("py", "<i2"), ("pz", "<i2"), ("ex", "<i2"), ("ey", "<i2"),
                 ("ez", "<i2"), ("pad", "<i2", (2,))])
COLOR = np.dtype([("r", "u1"), ("g", "u1"), ("b", "u1"), ("a", "u1")])
PIXEL = np.dtype([("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"), ("color", COLOR), ("y", "<i4"),
                  ("z", "<i4"), ("entity", "<i4")])
LIGHT = np.dtype([("x", "<i2"), ("y", "<i2"), ("z", "<i2"), ("radius", "<i2")])
SPRITE = np.dtype([("color", "<i4", (800,)), ("depth", "<i4", (800,)),
                   ("normal", "<f4", (800, 3))])
assert AABB.itemsize == 16 and PIXEL.itemsize == 28 and LIGHT.itemsize == 8
assert SPRITE.itemsize == 16000 and COLOR.itemsize == 4

COUNTER_FIELDS = ["pixels", "primary_bins", "primary_slot_tests", "primary_passed",
                  "primary_accepts", "shaded_px_lights", "lit_px_lights", "shadow_probes",
                  "shadow_slot_entries", "slab_tests", "pixels_hit"]


class View(C.Structure):
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("L", C.c_int32)]


class Atlas(C.Structure):
    """orc_atlas: sprites of per-entry width x height, concatenated tables."""
    _fields_ = [("n", C.c_int32), ("w", C.c_void_p), ("h", C.c_void_p), ("off", C.c_void_p),
                ("color", C.c_void_p), ("depth", C.c_void_p), ("normal", C.c_void_p)]


class Counters(C.Structure):
    _fields_ = [(f, C.c_uint64) for f in COUNTER_FIELDS]

    def as_dict(self):
        return {f: int(getattr(self, f)) for f in COUNTER_FIELDS}


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc, no GPU needed)."""
    src = os.path.join(HERE, "oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.orc_grid_volume.argtypes = [C.POINTER(View)]
        L.orc_grid_volume.restype = C.c_int
        L.orc_make_tile_floor.argtypes = [vp]
        L.orc_default_palette.argtypes = [vp]
        L.orc_scene_default.argtypes = [vp, C.c_int]
        L.orc_scene_default.restype = C.c_int
        L.orc_light_default.argtypes = [vp]
        L.orc_scene_synthetic.argtypes = [C.POINTER(View), C.c_uint64, C.c_int, vp, C.c_int, vp]
        L.orc_apply_key.argtypes = [C.c_int, vp, vp]
        L.orc_script_c_key.argtypes = [C.c_int]
        L.orc_script_c_key.restype = C.c_int
        L.orc_grid_build.argtypes = [C.POINTER(View), vp, C.c_int, vp, vp, vp]
        L.orc_trace_primary.argtypes = [C.POINTER(View), vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, C.c_int,
                                        C.c_int, C.POINTER(Counters)]
        L.orc_trace_primary.restype = C.c_int
        L.orc_render_frame_atlas.argtypes = [C.POINTER(View), vp, vp, C.c_int, C.POINTER(Atlas), vp, vp, C.c_int,
                                             vp, vp, vp, C.c_int, C.c_int, C.POINTER(Counters), C.c_int, vp, vp]
        L.orc_render_frame_atlas.restype = C.c_int
        L.orc_shade.argtypes = [C.POINTER(View), vp, vp, vp, vp, vp, C.c_int, vp, C.c_int,
                                C.c_int, C.POINTER(Counters)]
        L.orc_draw_overlay.argtypes = [C.POINTER(View), vp, vp, C.c_int, C.c_int, vp]
        L.orc_render_frame.argtypes = [C.POINTER(View), vp, vp, C.c_int, vp, vp, vp, C.c_int, vp,
                                       vp, vp, C.c_int, C.c_int, C.POINTER(Counters)]
        L.orc_render_frame.restype = C.c_int
        L.orc_fnv1a64.argtypes = [vp, C.c_size_t]
        L.orc_fnv1a64.restype = C.c_uint64
        L.orc_algorithmic_ops.argtypes = [C.POINTER(Counters)]
        L.orc_algorithmic_ops.restype = C.c_double
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def tile_floor() -> np.ndarray:
    s = np.zeros(1, SPRITE)
    lib().orc_make_tile_floor(_p(s))
    return s


def default_palette() -> np.ndarray:
    p = np.zeros(4, COLOR)
    lib().orc_default_palette(_p(p))
    return p


def scene_default() -> np.ndarray:
    n = lib().orc_scene_default(None, 0)
    a = np.zeros(n, AABB)
    lib().orc_scene_default(_p(a), n)
    return a


def light_default() -> np.ndarray:
    l = np.zeros(1, LIGHT)
    lib().orc_light_default(_p(l))
    return l


def scene_synthetic(W, H, L, n=10000, n_lights=16, seed=0xB200):
    v = View(W, H, L)
    a = np.zeros(n, AABB)
    l = np.zeros(n_lights, LIGHT)
    lib().orc_scene_synthetic(C.byref(v), seed, n, _p(a), n_lights, _p(l))
    return a, l


def apply_key(key: str, boxes: np.ndarray, lights: np.ndarray) -> None:
    """alternative.cpp:641-681 on entity 0 / light 0, in place."""
    lib().orc_apply_key(ord(key), _p(boxes), _p(lights))


def script_keys(script: str, frame: int) -> list[str]:
    """Keys delivered before `frame` by script 'C' or 'D' (SURVEY.md §8d)."""
    keys = []
    k = lib().orc_script_c_key(frame)
    if k:
        keys.append(chr(k))
    if script == "D" and frame >= 1:
        keys.append("o")
    return keys


def grid_build(W, H, L, boxes):
    v = View(W, H, L)
    V = lib().orc_grid_volume(C.byref(v))
    count = np.zeros(V, np.int32)
    bin_box = np.zeros(V * 8, AABB)
    bin_ent = np.zeros(V * 8, np.int32)
    lib().orc_grid_build(C.byref(v), _p(boxes), len(boxes), _p(count), _p(bin_box), _p(bin_ent))
    return count, bin_box, bin_ent


def ragged_atlas(sprites):
    """[(color (h,w) int, depth (h,w) int, normal (h,w,3) float), ...] -> the five arrays of an
    orc_atlas / par_set_atlas_sized: w, h, color, depth, normal (concatenated, row-major)."""
    w = np.array([s[0].shape[1] for s in sprites], np.int32)
    h = np.array([s[0].shape[0] for s in sprites], np.int32)
    color = np.concatenate([np.asarray(s[0], np.int32).reshape(-1) for s in sprites])
    depth = np.concatenate([np.asarray(s[1], np.int32).reshape(-1) for s in sprites])
    normal = np.concatenate([np.asarray(s[2], np.float32).reshape(-1, 3) for s in sprites])
    return w, h, np.ascontiguousarray(color), np.ascontiguousarray(depth), np.ascontiguousarray(normal)


def render(W, H, L, boxes, lights, atlas=None, palette=None, sprite_ids=None, row0=0, row1=None,
           want_rgba=True, want_gbuf=True, want_texel=True, threads=None, sized_atlas=None,
           dbg_light=None):
    """One frame (alternative.cpp:689-760, no overlay).  Returns a dict with rgba (H,W) COLOR,
    gbuf (H,W) PIXEL, texel (H,W) int32 and the §8(d) counters; only rows [row0,row1) are
    rendered (the rest stay zero).  sized_atlas = (w, h, color, depth, normal) from ragged_atlas()
    replaces the fixed 20x40 `atlas`; dbg_light = l adds 't' (H,W,4: t.xyz, Lambert term of light l)
    and 'factor' (H,W: acc + ambient) — the fp32 intermediates."""
    if sized_atlas is None:
        atlas = tile_floor() if atlas is None else atlas
        sized_atlas = ragged_atlas([(sp["color"].reshape(40, 20), sp["depth"].reshape(40, 20),
                                     sp["normal"].reshape(40, 20, 3)) for sp in atlas])
    aw, ah, acolor, adepth, anormal = sized_atlas
    aoff = np.concatenate([[0], np.cumsum(aw.astype(np.int64) * ah)[:-1]]).astype(np.int32)
    at = Atlas(len(aw), aw.ctypes.data, ah.ctypes.data, aoff.ctypes.data, acolor.ctypes.data,
               adepth.ctypes.data, anormal.ctypes.data)
    palette = default_palette() if palette is None else palette
    row1 = H if row1 is None else row1
    v = View(W, H, L)
    rgba = np.zeros((H, W), COLOR) if want_rgba else None
    gbuf = np.zeros((H, W), PIXEL) if want_gbuf else None
    texel = np.zeros((H, W), np.int32) if want_texel else None
    ctr = Counters()
    if sprite_ids is not None:
        sprite_ids = np.ascontiguousarray(sprite_ids, np.int32)
    old = os.environ.get("OMP_NUM_THREADS")
    if threads is not None:
        _omp_set_threads(threads)
    dbg_t = np.zeros((H, W, 4), np.float32) if dbg_light is not None else None
    dbg_f = np.zeros((H, W), np.float32) if dbg_light is not None else None
    rc = lib().orc_render_frame_atlas(C.byref(v), _p(boxes), _p(sprite_ids), len(boxes), C.byref(at),
                                      _p(palette), _p(lights), len(lights), _p(rgba), _p(gbuf),
                                      _p(texel), row0, row1, C.byref(ctr),
                                      -1 if dbg_light is None else int(dbg_light), _p(dbg_t), _p(dbg_f))
    if threads is not None:
        _omp_set_threads(0 if old is None else int(old))
    if rc != 0:
        raise MemoryError("orc_render_frame")
    return {"rgba": rgba, "gbuf": gbuf, "texel": texel, "counters": ctr.as_dict(),
            "ops": lib().orc_algorithmic_ops(C.byref(ctr)), "t": dbg_t, "factor": dbg_f}


def _omp_set_threads(n: int) -> None:
    try:
        gomp = C.CDLL("libgomp.so.1")
        if n <= 0:
            n = os.cpu_count() or 1
        gomp.omp_set_num_threads(int(n))
    except OSError:
        pass


def draw_overlay(W, H, L, gbuf, lights, frame, cx=0, cy=0) -> None:
    """alternative.cpp:762-772 into `frame` in place."""
    v = View(W, H, L)
    lib().orc_draw_overlay(C.byref(v), _p(gbuf), _p(lights), cx, cy, _p(frame))


def fnv1a64(buf: np.ndarray) -> int:
    b = np.ascontiguousarray(buf)
    return int(lib().orc_fnv1a64(_p(b), b.nbytes))
