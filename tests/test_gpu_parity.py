"""GPU parity: the CUDA path, called through the C ABI, against the oracle (same seeded
inputs) and against the reference's golden hashes.  Bar: bit-exact — grid contents, hit
entity ids, texel indices, world y/z, G-buffer bytes and final RGBA8 (SURVEY.md §8d
"Tolerance": integer outputs and RGBA exact; the fp32 intermediates — light direction, Lambert
term, acc + ambient — are exported by par_debug_intermediates and held to <= 1 ULP, measured 0)."""
import math

import numpy as np
import pytest

from conftest import sha256

pytestmark = pytest.mark.gpu


def _u32(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _render_both(par, O, W, H, L, boxes, lights, atlas=None, palette=None, sprite_ids=None,
                 check_grid=True):
    with par.Renderer(W, H, L) as r:
        r.set_atlas(atlas, palette)
        r.set_scene(boxes, sprite_ids)
        rgba, gbuf, stats = r.render(lights, want_gbuf=True)
        gbuf2, texel = r.gbuffer()
        grid = r.grid() if check_grid else None
    ref = O.render(W, H, L, np.ascontiguousarray(boxes, O.AABB), np.ascontiguousarray(lights, O.LIGHT),
                   atlas=None if atlas is None else np.ascontiguousarray(atlas, O.SPRITE),
                   palette=None if palette is None else np.ascontiguousarray(palette, O.COLOR),
                   sprite_ids=sprite_ids)
    assert gbuf.tobytes() == gbuf2.tobytes()
    if check_grid:
        count, bin_box, bin_ent = O.grid_build(W, H, L, np.ascontiguousarray(boxes, O.AABB))
        assert np.array_equal(grid[0], count), "bin counts (alternative.cpp:262-264)"
        slot = np.arange(8)[None, :]
        live = slot < count[:, None]
        assert np.array_equal(np.where(live, grid[1], -1), np.where(live, bin_ent.reshape(-1, 8), -1)), \
            "entity-index map in slot order (quirk Q2)"
    return (rgba, gbuf, texel, stats), ref


def _assert_frame_equal(got, ref):
    rgba, gbuf, texel, _ = got
    for name in ("entity", "y", "z"):
        assert np.array_equal(gbuf[name], ref["gbuf"][name]), f"G-buffer {name}"
    assert np.array_equal(texel, ref["texel"]), "texel index"
    assert gbuf.tobytes() == ref["gbuf"].tobytes(), "raw Pixel[] bytes"
    bad = np.argwhere(_u32(rgba) != _u32(ref["rgba"]))
    assert len(bad) == 0, f"{len(bad)} RGBA pixels differ, first at (row, col) {bad[:5].tolist()}"


def _random_atlas(rng, n_sprites, n_palette, axis_aligned=False):
    from par_b200 import SPRITE, COLOR
    atlas = np.zeros(n_sprites, SPRITE)
    atlas["color"] = rng.integers(0, n_palette, (n_sprites, 800))
    atlas["depth"] = rng.integers(0, 20, (n_sprites, 800))
    nrm = rng.standard_normal((n_sprites, 800, 3)).astype(np.float32)
    if axis_aligned:
        nrm = np.eye(3, dtype=np.float32)[rng.integers(0, 3, (n_sprites, 800))] * \
            rng.choice([-1.0, 1.0], (n_sprites, 800, 1)).astype(np.float32)
    atlas["normal"] = nrm
    pal = np.zeros(n_palette, COLOR)
    for ch in "rgba":
        pal[ch] = rng.integers(0, 256, n_palette)
    return atlas, pal


def _random_scene(rng, W, H, L, n, n_lights, cubes=False, n_sprites=1):
    from par_b200 import AABB, LIGHT
    a = np.zeros(n, AABB)
    a["px"] = rng.integers(-30, W + 30, n)
    a["py"] = rng.integers(-30, 160, n)
    a["pz"] = rng.integers(-60, L + 60, n)
    if cubes:
        a["ex"] = a["ey"] = a["ez"] = 20
    else:
        a["ex"] = rng.integers(1, 21, n)
        a["ey"] = rng.integers(0, 21, n)
        a["ez"] = rng.integers(0, 21, n)
    l = np.zeros(n_lights, LIGHT)
    l["x"] = rng.integers(-100, W + 200, n_lights)
    l["y"] = rng.integers(-50, 400, n_lights)
    l["z"] = rng.integers(-100, L + 100, n_lights)
    l["radius"] = 10
    ids = rng.integers(0, n_sprites, n).astype(np.int32) if n_sprites > 1 else None
    return a, l, ids


# ------------------------------------------------------------------ reference scenes

def test_c1_default_scene(par, oracle, golden):
    """Config 1: the reference's default scene at its built-in 480x320."""
    got, ref = _render_both(par, oracle, 480, 320, 320, par.scene_default(), par.light_default())
    _assert_frame_equal(got, ref)
    g = golden["tier1_480x320x320_frame0"]
    assert sha256(got[0]) == g["frame0_pre_overlay_sha256"]
    assert sha256(got[1]) == g["gbuf0_sha256"]
    # with the host-side overlay the frame equals the UNMODIFIED reference's
    frame = got[0].copy()
    par.draw_overlay(480, 320, got[1], par.light_default(), frame)
    assert sha256(frame) == golden["tier0_480x320x320_frame0"]["frame0_sha256"]
    assert got[3]["n_survivors"] == 968 and got[3]["n_inserts"] == 1095


@pytest.mark.parametrize("size,key", [((1920, 1080, 1080), "tier1_1920x1080x1080_frame0"),
                                      ((3840, 2160, 2160), "tier1_3840x2160x2160_frame0")])
def test_default_scene_large_views_vs_reference_hashes(par, golden, size, key):
    """Configs 2 and 4 (frame 0): full-size frames against hashes of the real reference's
    output — no oracle involved."""
    W, H, L = size
    g = golden[key]
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(par.scene_default())
        rgba, gbuf, _ = r.render(par.light_default(), want_gbuf=True)
    assert sha256(gbuf) == g["gbuf0_sha256"]
    assert sha256(rgba) == g["frame0_pre_overlay_sha256"]
    frame = rgba.copy()
    par.draw_overlay(W, H, gbuf, par.light_default(), frame)
    assert sha256(frame) == g["frame0_sha256"]


@pytest.mark.parametrize("k", range(18))
def test_random_scenes_vs_real_reference_hashes(par, oracle, golden_scenes, k):
    """Scenes the REAL reference rendered through its scene hook (random / ragged / lattice-snapped
    boxes, the light free or on a box face, views with length != height, scripted frames): the
    G-buffer bytes, the shaded frame and the final frames from the C ABI hash to the reference's."""
    from conftest import decode_scene, sha256
    if k >= len(golden_scenes["scenes"]):
        pytest.skip("fewer scenes in the golden file")
    e = golden_scenes["scenes"][k]
    W, H, L = e["view"]
    boxes, lights = decode_scene(e, par.AABB, par.LIGHT)
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        for f in range(e["frames"]):
            if e["script"]:
                for key in oracle.script_keys(e["script"], f):
                    par.apply_key(key, boxes, lights)
            r.set_scene(boxes)
            rgba, gbuf, _ = r.render(lights, want_gbuf=True)
            if f == 0:
                assert sha256(gbuf) == e["gbuf0_sha256"]
                assert sha256(rgba) == e["frame0_pre_overlay_sha256"]
            par.draw_overlay(W, H, gbuf, lights, rgba)
            assert "%016x" % oracle.fnv1a64(rgba) == e["fnv1a64"][f], f"frame {f}"


@pytest.mark.parametrize("view", ["480x320x320", "480x320x640", "640x480x200", "200x40x40", "40x1000x120"])
def test_default_scene_other_views_vs_real_reference_hashes(par, oracle, golden_scenes, view):
    """The built-in scene through views with length != height (light moved inside the grid),
    3 frames of key script D, against the real reference's hashes."""
    from conftest import sha256
    e = golden_scenes["default_scene"].get(view)
    if e is None:
        pytest.skip("no golden for this view")
    W, H, L = e["view"]
    boxes, lights = par.scene_default(), par.light_default()
    lights[0]["x"], lights[0]["y"], lights[0]["z"], lights[0]["radius"] = e["light"]
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        for f in range(e["frames"]):
            for key in oracle.script_keys(e["script"], f):
                par.apply_key(key, boxes, lights)
            r.set_scene(boxes)
            rgba, gbuf, _ = r.render(lights, want_gbuf=True)
            if f == 0:
                assert sha256(gbuf) == e["gbuf0_sha256"]
                assert sha256(rgba) == e["frame0_pre_overlay_sha256"]
            par.draw_overlay(W, H, gbuf, lights, rgba)
            assert "%016x" % oracle.fnv1a64(rgba) == e["fnv1a64"][f], f"frame {f}"


def test_script_d_frames_vs_reference_hashes(par, oracle, golden):
    """Config 4: per-frame scene upload + render under key script D (player and light move),
    FNV-1a-64 of every sampled frame (overlay applied) against the real reference's."""
    want = golden["tier1_1920x1080x1080_scriptD_240"]["fnv1a64"]
    boxes, lights = par.scene_default(), par.light_default()
    W, H, L = 1920, 1080, 1080
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        for f in range(240):
            for k in oracle.script_keys("D", f):
                par.apply_key(k, boxes, lights)
            if f % 8 == 0 or f == 239:
                r.set_scene(boxes)
                rgba, gbuf, _ = r.render(lights, want_gbuf=True)
                par.draw_overlay(W, H, gbuf, lights, rgba)
                assert "%016x" % oracle.fnv1a64(rgba) == want[f], f"frame {f}"


def test_script_c_frames_480(par, oracle, golden):
    want = golden["tier0_480x320x320_scriptC_240"]["fnv1a64"]
    boxes, lights = par.scene_default(), par.light_default()
    with par.Renderer(480, 320, 320) as r:
        r.set_atlas()
        for f in range(240):
            for k in oracle.script_keys("C", f):
                par.apply_key(k, boxes, lights)
            r.set_scene(boxes)
            rgba, gbuf, _ = r.render(lights, want_gbuf=True)
            par.draw_overlay(480, 320, gbuf, lights, rgba)
            assert "%016x" % oracle.fnv1a64(rgba) == want[f], f"frame {f}"


# ------------------------------------------------------------------ synthetic scenes vs oracle

@pytest.mark.parametrize("seed", range(6))
def test_random_scenes_multi_sprite_multi_light(par, oracle, seed):
    """Random boxes (ragged extents, partly outside the view), random multi-sprite atlas with
    NON-axis-aligned float normals (exposes any FMA contraction or reordering), lights inside
    and outside the grid."""
    rng = np.random.default_rng(1000 + seed)
    W, H, L = [(480, 320, 320), (640, 480, 480), (320, 640, 640)][seed % 3]
    n_sprites, n_pal = 1 + seed % 5 * 3, 2 + seed
    atlas, pal = _random_atlas(rng, n_sprites, n_pal)
    boxes, lights, ids = _random_scene(rng, W, H, L, 1500 + 700 * seed, 1 + 3 * seed,
                                       n_sprites=n_sprites)
    got, ref = _render_both(par, oracle, W, H, L, boxes, lights, atlas, pal, ids)
    _assert_frame_equal(got, ref)


@pytest.mark.parametrize("view", [(480, 320, 640), (640, 480, 200), (200, 40, 40), (40, 1000, 120)])
def test_view_length_differs_from_height(par, oracle, view):
    """The reference has view_length == view_height (alternative.cpp:118-119); the ABI takes them
    independently, and tiny / very tall views must work too."""
    W, H, L = view
    rng = np.random.default_rng(W * 7 + H * 3 + L)
    boxes, lights, _ = _random_scene(rng, W, H, L, 800, 4)
    got, ref = _render_both(par, oracle, W, H, L, boxes, lights)
    _assert_frame_equal(got, ref)


def test_dense_overflowing_bins(par, oracle):
    """Quirk Q2: many entities per bin so that rings wrap (n = 8, 9, 16, 17 ... inserts)."""
    rng = np.random.default_rng(7)
    W, H, L = 480, 320, 320
    boxes, lights, _ = _random_scene(rng, W, H, L, 9000, 4, cubes=True)
    count = oracle.grid_build(W, H, L, np.ascontiguousarray(boxes, oracle.AABB))[0]
    got, ref = _render_both(par, oracle, W, H, L, boxes, lights)
    _assert_frame_equal(got, ref)
    assert got[3]["n_inserts"] > 8 * np.count_nonzero(count)  # bins did wrap


def test_c3_recipe_reduced(par, oracle):
    """The C3 recipe at a view the oracle renders in seconds."""
    W, H, L = 1280, 720, 720
    boxes, lights = par.scene_synthetic(W, H, L, n=3000, n_lights=16)
    got, ref = _render_both(par, oracle, W, H, L, boxes, lights)
    _assert_frame_equal(got, ref)


def test_c3_full_size_row_sample(par, oracle):
    """Config 3 at full size (3840x2160, 10k sprites, 16 lights): the whole frame on the GPU,
    the oracle on three row bands (the grid is always built whole)."""
    W, H, L = 3840, 2160, 2160
    boxes, lights = par.scene_synthetic(W, H, L)
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(boxes)
        rgba, gbuf, stats = r.render(lights, want_gbuf=True)
    assert stats["n_survivors"] == 9750 and stats["n_inserts"] == 40832
    for row0 in (0, 1013, 2120):
        ref = oracle.render(W, H, L, boxes.view(oracle.AABB), lights.view(oracle.LIGHT),
                            row0=row0, row1=row0 + 40)
        assert gbuf[row0:row0 + 40].tobytes() == ref["gbuf"][row0:row0 + 40].tobytes()
        assert np.array_equal(_u32(rgba[row0:row0 + 40]), _u32(ref["rgba"][row0:row0 + 40]))


# ------------------------------------------------------------------ quirk micro-scenes

def _cube(x, y, z, e=20):
    return (x, y, z, e, e, e, (0, 0))


def test_quirk_nan_slab_and_zero_direction(par, oracle):
    """Q13: lights whose x/y/z equals pixel coordinates and box faces (0 * inf = NaN through
    std::min/std::max), and a light exactly ON a visible surface point (0/0 direction)."""
    from par_b200 import AABB, LIGHT
    boxes = np.array([_cube(100, 0, 100), _cube(120, 0, 100), _cube(100, 20, 120),
                      _cube(200, 0, 60), _cube(200, 40, 60), _cube(240, 0, 140),
                      _cube(60, 0, 200), _cube(60, 20, 200), _cube(300, 0, 100)], AABB)
    lights = np.zeros(6, LIGHT)
    lights["x"] = [100, 120, 210, 260, 60, 300]
    lights["y"] = [20, 40, 20, 20, 60, 20]     # top faces are at y = 20 / 40 / 60
    lights["z"] = [100, 120, 70, 140, 200, 110]
    got, ref = _render_both(par, oracle, 480, 320, 320, boxes, lights)
    _assert_frame_equal(got, ref)


def test_quirk_depth_ties_and_early_out(par, oracle):
    """Q8 (ties keep the first in bin_z/slot order) and Q9 (two adjacent hit bins end the
    march; an empty bin resets the run)."""
    from par_b200 import AABB
    rows = []
    for k in range(6):           # identical depth keys: same y - z
        rows.append(_cube(40, 10 + 5 * k, 10 + 5 * k))
    for k in range(5):           # stacks along z in adjacent bins, all covering the same pixels
        rows.append(_cube(200, 0, 40 * k + 10))
        rows.append(_cube(200, 30, 40 * k + 10))
    rows += [_cube(320, 0, 20), _cube(320, 0, 100), _cube(320, 50, 180), _cube(320, 90, 260)]
    boxes = np.array(rows, AABB)
    got, ref = _render_both(par, oracle, 480, 320, 320, boxes, par.light_default())
    _assert_frame_equal(got, ref)


def test_quirk_light_outside_grid_and_far_away(par, oracle):
    """Q18: light bins outside the grid on every side, including walks longer than one CTA's
    worth of steps (more than 320 bins away)."""
    from par_b200 import LIGHT
    rng = np.random.default_rng(3)
    boxes, _, _ = _random_scene(rng, 480, 320, 320, 1200, 1, cubes=True)
    lights = np.zeros(8, LIGHT)
    lights["x"] = [480, -500, 20000, 240, 240, -32000, 30000, 100]
    lights["y"] = [160, 100, 300, 5000, -4000, 32000, -32000, 50]
    lights["z"] = [80, -300, 150, 100, 20000, 100, 32000, -16000]
    got, ref = _render_both(par, oracle, 480, 320, 320, boxes, lights)
    _assert_frame_equal(got, ref)


@pytest.mark.parametrize("seed", [11, 12])
def test_dense_scene_many_lights_split_walks(par, oracle, seed):
    """Rounds of the render kernel under pressure: a dense scene (every bin along a walk occupied, several
    entities per bin) and 12 lights — some far outside the grid — so that walks overflow the occupied-bin list
    and the de-duplication set, are cut into step ranges, re-walked with measured densities and mixed with
    whole walks of the following lights.  Every (pixel, light) term must still be there exactly once."""
    from par_b200 import AABB, LIGHT
    rng = np.random.default_rng(seed)
    W, H, L = 1200, 440, 440
    n = 9000
    boxes = np.zeros(n, AABB)
    boxes["px"] = rng.integers(0, W - 20, n)
    boxes["py"] = rng.integers(0, 260, n)
    boxes["pz"] = rng.integers(0, L - 20, n)
    boxes["ex"] = boxes["ey"] = boxes["ez"] = 20
    lights = np.zeros(12, LIGHT)
    lights["x"] = rng.integers(-2500, 4000, 12)
    lights["y"] = rng.integers(20, 700, 12)
    lights["z"] = rng.integers(-1500, 2500, 12)
    lights["x"][:4] = rng.integers(0, W, 4)  # some inside the view
    lights["z"][:4] = rng.integers(0, L, 4)
    got, ref = _render_both(par, oracle, W, H, L, boxes, lights, check_grid=False)
    _assert_frame_equal(got, ref)


def test_empty_and_degenerate_scenes(par, oracle):
    from par_b200 import AABB, LIGHT
    got, ref = _render_both(par, oracle, 480, 320, 320, np.zeros(0, AABB), par.light_default())
    _assert_frame_equal(got, ref)
    assert sha256(got[0]) == sha256(np.full((320, 480), 31 | 31 << 8 | 31 << 16, np.uint32))  # Q10
    # zero lights: ambient only
    got, ref = _render_both(par, oracle, 480, 320, 320, par.scene_default()[:2000],
                            np.zeros(0, LIGHT))
    _assert_frame_equal(got, ref)
    # zero-extent boxes are legal and invisible
    flat = np.array([(100, 0, 100, 20, 0, 0, (0, 0)), (150, 0, 100, 0, 20, 20, (0, 0))], AABB)
    got, ref = _render_both(par, oracle, 480, 320, 320, flat, par.light_default())
    _assert_frame_equal(got, ref)


def test_maximum_lights(par, oracle):
    rng = np.random.default_rng(11)
    boxes, lights, _ = _random_scene(rng, 480, 320, 320, 2500, 64, cubes=True)
    got, ref = _render_both(par, oracle, 480, 320, 320, boxes, lights)
    _assert_frame_equal(got, ref)


# ------------------------------------------------------------------ bands, errors, properties

def test_row_bands_reassemble(par):
    """Multi-GPU building block: bands (not multiples of 40) written in place give the
    single-context frame, byte for byte."""
    W, H, L = 640, 480, 480
    boxes, lights = par.scene_synthetic(W, H, L, n=1500, n_lights=5)
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(boxes)
        full, gfull, _ = r.render(lights, want_gbuf=True)
    out = np.zeros((H, W), par.COLOR)
    gout = np.zeros((H, W), par.PIXEL)
    for a, b in [(0, 97), (97, 100), (100, 333), (333, 480)]:
        with par.Renderer(W, H, L, row_begin=a, row_end=b) as r:
            r.set_atlas()
            r.set_scene(boxes)
            band = np.zeros((H, W), par.COLOR)
            gband = np.zeros((H, W), par.PIXEL)
            r.render(lights, out=band)
            g2, _ = r.gbuffer()
            assert not _u32(band[:a]).any() and not _u32(band[b:]).any()
            out[a:b] = band[a:b]
            gout[a:b] = g2[a:b]
    assert np.array_equal(_u32(out), _u32(full))
    assert gout.tobytes() == gfull.tobytes()


@pytest.mark.parametrize("n", [2, 3, 8])
def test_stripes_reassemble(par, n):
    """Interleaved 40-row stripes (the multi-GPU partition): raster output of every stripe set,
    and the stripe-major staging frame + un-stripe copy, both give the one-context frame."""
    import torch
    W, H, L = 640, 480, 480  # 12 tile rows: not a multiple of 8 -> padded staging
    boxes, lights = par.scene_synthetic(W, H, L, n=1500, n_lights=5)
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(boxes)
        full, _ = r.render(lights)
    out = np.zeros((H, W), par.COLOR)
    staging = None
    rays = 0
    for i in range(n):
        with par.Renderer(W, H, L, stripe_count=n, stripe_index=i) as r:
            r.set_atlas()
            r.set_scene(boxes)
            part = np.zeros((H, W), par.COLOR)
            _, st = r.render(lights, out=part)
            rays += st["rays"]
            own = np.zeros(H, bool)
            for t in range(i, H // 40, n):
                own[t * 40:t * 40 + 40] = True
            assert not _u32(part[~own]).any()
            out[own] = part[own]
            if staging is None:
                staging = torch.zeros(r.staging_bytes(), dtype=torch.uint8, device="cuda")
            r.render_device_striped(lights, staging.data_ptr())
            r.sync()
            last = r
            if i == n - 1:
                raster = torch.zeros(H * W * 4, dtype=torch.uint8, device="cuda")
                r.unstripe_device(staging.data_ptr(), raster.data_ptr())
                r.sync()
    assert rays == W * H * 6
    assert np.array_equal(_u32(out), _u32(full))
    assert np.array_equal(raster.cpu().numpy(), full.view(np.uint8).reshape(-1))
    from par_b200.bands import unstripe
    assert np.array_equal(unstripe(staging, W, H, n).cpu().numpy(), full.view(np.uint8).reshape(-1))


@pytest.mark.parametrize("n,split,band", [(2, 2, None), (4, 2, None), (8, 4, None), (3, 2, None), (8, 2, (55, 301))])
def test_split_stripes_reassemble(par, n, split, band):
    """stripe_split: a tile row is cut into `split` stripes of equal width and stripe v = tile row * split +
    segment goes to rank v % n (equal stripe counts per rank when height / 40 is not a multiple of the rank
    count: 9 tile rows here).  Every form of output a striped context has — the blocking par_render into a
    host frame, par_read_stripes and the pipelined par_submit_frame into ONE shared host frame, and the fused
    peer stores of par_render_device_peers (all contexts on this one device) — reassembles the one-context
    frame; the stripe-major staging calls refuse such a context."""
    W, H, L = 640, 360, 360
    boxes, lights = par.scene_synthetic(W, H, L, n=1500, n_lights=5)
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(boxes)
        full, _ = r.render(lights)
    a, b = band or (0, H)
    tps = W // 40 // split
    own = []  # per rank: the pixels it owns
    for i in range(n):
        m = np.zeros((H, W), bool)
        for v in range(i, (H // 40) * split, n):
            t, sgm = v // split, (v % split + v // math.lcm(n, split)) % split  # segments rotate: par_device.cuh
            m[t * 40:t * 40 + 40, sgm * tps * 40:(sgm + 1) * tps * 40] = True
        m[:a] = False
        m[b:] = False
        own.append(m)
    assert sum(int(m.sum()) for m in own) == (b - a) * W
    if band is None and (H // 40 * split) % n == 0:
        assert len({int(m.sum()) for m in own}) == 1  # the point of the split: equal shares
    kw = dict(row_begin=a, row_end=b) if band else {}
    rens = [par.Renderer(W, H, L, stripe_count=n, stripe_index=i, stripe_split=split, **kw) for i in range(n)]
    try:
        out = np.zeros((H, W), par.COLOR)
        shared = par.pinned_empty((H, W), par.COLOR)
        piped = par.pinned_empty((H, W), par.COLOR)
        h_boxes = par.pinned_empty(len(boxes), par.AABB)
        h_boxes[:] = boxes
        _u32(shared)[:] = 0xDEADBEEF
        _u32(piped)[:] = 0xDEADBEEF
        rays = 0
        for i, r in enumerate(rens):
            r.set_atlas()
            r.set_scene(boxes)
            part = np.zeros((H, W), par.COLOR)
            _, st = r.render(lights, out=part)
            rays += st["rays"]
            assert not _u32(part)[~own[i]].any()
            out[own[i]] = part[own[i]]
            r.read_stripes(shared)
            r.sync()
            with pytest.raises(par.ParError):
                r.render_device_striped(lights, 1)
            # cursor: only on a pixel the context renders
            ys, xs = np.nonzero(own[i])
            r.set_cursor(int(xs[0]), int(ys[0]))
            other = np.nonzero(~own[i])
            with pytest.raises(par.ParError):
                r.set_cursor(int(other[1][0]), int(other[0][0]))
        assert rays == (b - a) * W * 6
        assert np.array_equal(_u32(out[a:b]), _u32(full[a:b]))
        assert np.array_equal(_u32(shared[a:b]), _u32(full[a:b]))
        assert (_u32(shared[:a]) == 0xDEADBEEF).all() and (_u32(shared[b:]) == 0xDEADBEEF).all()
        for r in rens:
            r.submit_frame(h_boxes, lights, piped)
        for r in rens:
            r.wait_frame()
        assert np.array_equal(_u32(piped[a:b]), _u32(full[a:b]))
        assert (_u32(piped[:a]) == 0xDEADBEEF).all() and (_u32(piped[b:]) == 0xDEADBEEF).all()
        if band is None:
            # fused frame exchange: every context stores its tiles into every other context's frame too
            for i, r in enumerate(rens):
                for j, q in enumerate(rens):
                    if i != j:
                        r.peer_set(j, q.device_frame())
            for r in rens:
                r.render_device_peers(lights)
            for r in rens:
                r.sync()
            for i, r in enumerate(rens):
                got = r.read_frame()
                r.sync()
                assert np.array_equal(_u32(got), _u32(full)), f"context {i}"
    finally:
        for r in rens:
            r.close()


@pytest.mark.parametrize("n,band", [(1, None), (2, None), (3, None), (8, None), (1, (97, 333)), (3, (50, 430))])
def test_read_stripes_into_one_host_frame(par, n, band):
    """par_read_stripes: every context copies only the rows it owns into ONE shared host frame
    (the parallel PCIe readback of the multi-GPU host path); together they give the full frame.
    Rows nobody owns (outside a band) stay untouched."""
    W, H, L = 640, 480, 480
    boxes, lights = par.scene_synthetic(W, H, L, n=1500, n_lights=5)
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(boxes)
        full, _ = r.render(lights)
    a, b = band or (0, H)
    host = par.pinned_empty((H, W), par.COLOR)
    _u32(host)[:] = 0xDEADBEEF
    for i in range(n):
        kw = dict(stripe_count=n, stripe_index=i) if n > 1 else {}
        with par.Renderer(W, H, L, row_begin=a, row_end=b, **kw) as r:
            r.set_atlas()
            r.set_scene(boxes)
            r.render_device(lights)
            r.read_stripes(host)
            r.sync()
    assert np.array_equal(_u32(host[a:b]), _u32(full[a:b]))
    assert (_u32(host[:a]) == 0xDEADBEEF).all() and (_u32(host[b:]) == 0xDEADBEEF).all()


def test_register_host_memory(par):
    """par_register_host: page-lock caller-owned memory (e.g. a shared-memory frame)."""
    W, H, L = 320, 200, 200
    boxes, lights = par.scene_synthetic(W, H, L, n=300, n_lights=2)
    buf = np.zeros((H, W), par.COLOR)
    assert par.lib().par_register_host(buf.ctypes.data, buf.nbytes) == 0
    try:
        with par.Renderer(W, H, L) as r:
            r.set_atlas()
            r.set_scene(boxes)
            want, _ = r.render(lights)
            r.read_stripes(buf)
            r.sync()
        assert np.array_equal(_u32(buf), _u32(want))
    finally:
        par.lib().par_unregister_host(buf.ctypes.data)


@pytest.mark.parametrize("stripes", [1, 2])
def test_pipelined_frames_moving_scene(par, oracle, stripes):
    """par_submit_frame / par_wait_frame: 24 frames of key script D (player and light move every
    frame), two in flight, each with its own pinned AABB array and host frame.  Every frame equals
    the synchronous par_render of the same scene; with 2 striped contexts feeding the same host
    frames the union is the full frame."""
    W, H, L = 640, 480, 480
    boxes, lights = par.scene_default(), par.light_default()
    scenes, light_seq = [], []
    for f in range(24):
        for k in oracle.script_keys("D", f):
            par.apply_key(k, boxes, lights)
        scenes.append(boxes.copy())
        light_seq.append(lights.copy())
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        want = []
        for f in range(24):
            r.set_scene(scenes[f])
            want.append(r.render(light_seq[f])[0].copy())
    assert any(not np.array_equal(_u32(want[0]), _u32(w)) for w in want[1:])  # the scene does move
    kw = [dict(stripe_count=stripes, stripe_index=i) if stripes > 1 else {} for i in range(stripes)]
    rens = [par.Renderer(W, H, L, **k) for k in kw]
    try:
        h_boxes = [par.pinned_empty(len(boxes), par.AABB) for _ in range(2)]
        h_out = [par.pinned_empty((H, W), par.COLOR) for _ in range(2)]
        for r in rens:
            r.set_atlas()
        for f in range(25):
            if f < 24:
                h_boxes[f & 1][:] = scenes[f]
                for r in rens:
                    r.submit_frame(h_boxes[f & 1], light_seq[f], h_out[f & 1])
            if f >= 1:
                for r in rens:
                    st = r.wait_frame()
                    assert st["ms_total"] > 0 and st["rays"] == W * H * 2 // stripes
                assert np.array_equal(_u32(h_out[(f - 1) & 1]), _u32(want[f - 1])), f"frame {f - 1}"
        # synchronous calls work again once nothing is in flight
        rens[0].set_scene(scenes[3])
        part, _ = rens[0].render(light_seq[3])
        own = np.zeros(H, bool)
        for t in range(0, H // 40, stripes):
            own[t * 40:t * 40 + 40] = True
        assert np.array_equal(_u32(part[own]), _u32(want[3][own]))
    finally:
        for r in rens:
            r.close()


def test_cursor_probe_and_overlay(par, oracle):
    """par_set_cursor / par_cursor_pixel: the record under the cursor arrives with every frame
    (blocking and pipelined) and equals the G-buffer's; the overlay drawn from it equals the one
    drawn from the whole G-buffer."""
    W, H, L = 640, 480, 480
    boxes, lights = par.scene_default(), par.light_default()
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        with pytest.raises(par.ParError) as e:
            r.set_cursor(W, 0)
        assert e.value.code == -1
        r.set_cursor(300, 200)
        with pytest.raises(par.ParError) as e:
            r.cursor_pixel()
        assert e.value.code == -6  # nothing rendered yet
        outs = [par.pinned_empty((H, W), par.COLOR) for _ in range(2)]
        hb = [par.pinned_empty(len(boxes), par.AABB) for _ in range(2)]
        for f in range(6):
            for k in oracle.script_keys("D", f):
                par.apply_key(k, boxes, lights)
            for cx, cy in ((300, 200), (0, 0), (639, 479)):
                r.set_cursor(cx, cy)
                r.set_scene(boxes)
                rgba, gbuf, _ = r.render(lights, want_gbuf=True)
                under = r.cursor_pixel()
                assert under.tobytes() == gbuf[cy, cx].tobytes()
                a, b = rgba.copy(), rgba.copy()
                par.draw_overlay(W, H, gbuf, lights, a, cx, cy)
                par.draw_overlay_at(W, H, under, lights, b, cx)
                assert np.array_equal(_u32(a), _u32(b))
                r.render_device(lights)  # asynchronous call: cursor_pixel waits for it
                assert r.cursor_pixel().tobytes() == gbuf[cy, cx].tobytes()
            # pipelined, two in flight with different cursors: each wait_frame brings ITS frame's record
            r.set_scene(boxes)
            _, gbuf, _ = r.render(lights, want_gbuf=True)
            hb[0][:] = boxes
            r.set_cursor(300, 200)
            r.submit_frame(hb[0], lights, outs[0])
            r.set_cursor(5, 470)
            r.submit_frame(hb[0], lights, outs[1])
            r.wait_frame()
            assert r.cursor_pixel().tobytes() == gbuf[200, 300].tobytes(), f"pipelined frame {f}a"
            r.wait_frame()
            assert r.cursor_pixel().tobytes() == gbuf[470, 5].tobytes(), f"pipelined frame {f}b"
        r.set_cursor(-1, -1)
        with pytest.raises(par.ParError):
            r.cursor_pixel()
    with par.Renderer(W, H, L, stripe_count=2, stripe_index=1) as r:
        r.set_cursor(10, 45)  # tile row 1: owned
        with pytest.raises(par.ParError):
            r.set_cursor(10, 5)  # tile row 0: rank 0's


def test_pipelined_frames_call_order_and_errors(par):
    from par_b200 import AABB
    W, H, L = 480, 320, 320
    boxes, lights = par.scene_default(), par.light_default()
    out = [par.pinned_empty((H, W), par.COLOR) for _ in range(3)]
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        with pytest.raises(par.ParError) as e:
            r.wait_frame()
        assert e.value.code == -6  # nothing in flight
        bad = np.array([(10, 0, 10, 30, 20, 20, (0, 0))], AABB)
        r.submit_frame(boxes, lights, out[0])
        r.submit_frame(bad, lights, out[1])
        with pytest.raises(par.ParError) as e:
            r.submit_frame(boxes, lights, out[2])
        assert e.value.code == -6  # two in flight
        with pytest.raises(par.ParError) as e:
            r.render(lights)
        assert e.value.code == -6  # synchronous call while frames are in flight
        r.wait_frame()  # frame 0 is fine
        with pytest.raises(par.ParError) as e:
            r.wait_frame()  # frame 1 had the bad scene
        assert e.value.code == -5
        r.set_scene(boxes)
        want, _ = r.render(lights)
        assert np.array_equal(_u32(out[0]), _u32(want))
        r.submit_frame(boxes, lights, out[2])  # still usable
        r.wait_frame()
        assert np.array_equal(_u32(out[2]), _u32(want))


def test_determinism_and_rebuild_idempotence(par):
    W, H, L = 1920, 1080, 1080
    boxes, lights = par.scene_synthetic(W, H, L, n=10000, n_lights=8)
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(boxes)
        a, _ = r.render(lights)
        grid_a = r.grid()
        r.rebuild_grid()
        r.rebuild_grid()
        b, _ = r.render(lights)
        grid_b = r.grid()
    assert np.array_equal(_u32(a), _u32(b))
    assert np.array_equal(grid_a[0], grid_b[0]) and np.array_equal(grid_a[1], grid_b[1])


def test_error_paths(par):
    from par_b200 import AABB
    with par.Renderer(480, 320, 320) as r:
        with pytest.raises(par.ParError) as e:
            r.set_scene(par.scene_default())
        assert e.value.code == -6  # atlas first
        r.set_atlas()
        with pytest.raises(par.ParError) as e:
            r.render(par.light_default())
        assert e.value.code == -6  # scene first
        bad = np.array([(10, 0, 10, 30, 20, 20, (0, 0))], AABB)  # extent.x 30 > sprite width
        r.set_scene(bad)
        with pytest.raises(par.ParError) as e:
            r.render(par.light_default())
        assert e.value.code == -5
        r.set_scene(par.scene_default()[:100])
        r.render(par.light_default())  # context stays usable
        with pytest.raises(par.ParError):
            r.set_scene(par.scene_default()[:10], sprite_ids=np.full(10, 3, np.int32))
            r.render(par.light_default())


# ------------------------------------------------------------------ round 2: full-size configs, A/B, intermediates

def _workload(par, name):
    if name in ("c1", "c2", "c4"):
        return par.scene_default(), par.light_default()
    W, H, L = {"c3": (3840, 2160, 2160), "c5": (7680, 4320, 4320), "c5b": (7680, 4320, 4320)}[name]
    return par.scene_synthetic(W, H, L, n=40000 if name == "c5b" else 10000, n_lights=16)


@pytest.fixture(scope="session")
def workload_ops():
    import json
    import os
    from conftest import ROOT
    with open(os.path.join(ROOT, "tests", "golden", "workload_ops.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", ["c2", "c3", "c5", "c5b"])
def test_full_size_configs_vs_oracle_hashes(par, oracle, workload_ops, name):
    """Every BASELINE config at FULL size: the whole frame, the raw Pixel[] G-buffer and the texel-index
    plane hash (FNV-1a-64) to what the oracle rendered for the same recipe
    (tests/golden/make_workload_ops.py; c5 / c5b take minutes on the CPU, hence committed hashes)."""
    g = workload_ops[name]
    W, H, L = g["view"]
    boxes, lights = _workload(par, name)
    assert len(boxes) == g["n_entities"] and len(lights) == g["n_lights"]
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(boxes)
        rgba, gbuf, _ = r.render(lights, want_gbuf=True)
        assert "%016x" % oracle.fnv1a64(rgba) == g["frame_fnv1a64"], "frame"
        assert "%016x" % oracle.fnv1a64(gbuf) == g["gbuf_fnv1a64"], "G-buffer bytes"
        del gbuf
        _, texel = r.gbuffer()
        assert "%016x" % oracle.fnv1a64(texel) == g["texel_fnv1a64"], "texel indices"
        # the production frame (no G-buffer written, longest-first tile order from the previous frame's
        # costs, graph replay) is the same frame
        for _ in range(3):
            r.render_resident(lights)
        got = r.read_frame()
        r.sync()
        assert np.array_equal(_u32(got), _u32(rgba))


@pytest.mark.parametrize("name", ["c3", "c2"])
def test_shaft_cull_off_gives_identical_bytes(par, monkeypatch, name):
    """The shaft cull only drops boxes no ray of the group can hit: switching it off
    (PAR_DEBUG_FLAGS=1) must not change a byte of a full-size frame."""
    W, H, L = 3840, 2160, 2160
    boxes, lights = _workload(par, name)
    frames = []
    for flags in ("0", "1"):
        monkeypatch.setenv("PAR_DEBUG_FLAGS", flags)
        with par.Renderer(W, H, L) as r:
            r.set_atlas()
            r.set_scene(boxes)
            frames.append(r.render(lights)[0])
    assert np.array_equal(_u32(frames[0]), _u32(frames[1]))


def _ulp_diff(a, b):
    """Distance in units in the last place between float32 arrays (NaN == NaN, +0 == -0)."""
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7fffffff), ia)
    ib = np.where(ib < 0, -(ib & 0x7fffffff), ib)
    d = np.abs(ia - ib)
    both_nan = np.isnan(a) & np.isnan(b)
    return np.where(both_nan, 0, d)


@pytest.mark.parametrize("seed", range(3))
def test_fp32_intermediates_within_one_ulp(par, oracle, seed):
    """SURVEY.md §8(d) tolerance: the fp32 intermediates — the L1-normalised light direction t, the
    Lambert term and acc + ambient — within 1 ULP of the oracle's for every hit pixel and light
    (stated allowance 1 ULP; the design target, IEEE arithmetic without contraction, is 0 and the
    measured maximum is asserted to be 0 too), with RGBA exact."""
    rng = np.random.default_rng(4200 + seed)
    W, H, L = [(480, 320, 320), (640, 480, 480), (320, 640, 640)][seed]
    atlas, pal = _random_atlas(rng, 3, 5)
    boxes, lights, ids = _random_scene(rng, W, H, L, 2500, 5, n_sprites=3)
    tol_ulp = 1
    with par.Renderer(W, H, L) as r:
        r.set_atlas(atlas, pal)
        r.set_scene(boxes, ids)
        rgba, gbuf, _ = r.render(lights, want_gbuf=True)
        _, texel = r.gbuffer()
        hit = texel >= 0
        assert hit.any() and (~hit).any()
        worst = 0
        for l in range(len(lights)):
            ref = oracle.render(W, H, L, np.ascontiguousarray(boxes, oracle.AABB), np.ascontiguousarray(lights, oracle.LIGHT),
                                atlas=np.ascontiguousarray(atlas, oracle.SPRITE), palette=np.ascontiguousarray(pal, oracle.COLOR),
                                sprite_ids=ids, dbg_light=l)
            t, f = r.intermediates(lights, l)
            if l == 0:
                assert np.array_equal(_u32(rgba), _u32(ref["rgba"]))
            dt = _ulp_diff(t[hit], ref["t"][hit])
            df = _ulp_diff(f[hit], ref["factor"][hit])
            worst = max(worst, int(dt.max()), int(df.max()))
            assert dt.max() <= tol_ulp and df.max() <= tol_ulp, f"light {l}: t {dt.max()} ulp, factor {df.max()} ulp"
            assert not t[~hit].any() and not f[~hit].any()  # never evaluated for miss pixels (quirk Q19)
        assert worst == 0, f"fp32 intermediates differ by up to {worst} ulp (allowed {tol_ulp}, expected 0)"


# ------------------------------------------------------------------ round 2: sprites of their own size (Q7 lifted)

def _sized_sprites(rng, dims, n_palette):
    out = []
    for (w, h) in dims:
        out.append((rng.integers(0, n_palette, (h, w)), rng.integers(0, 24, (h, w)),
                    rng.standard_normal((h, w, 3)).astype(np.float32)))
    return out


def _sized_scene(rng, W, H, L, n, dims, n_lights):
    from par_b200 import AABB, LIGHT
    a = np.zeros(n, AABB)
    ids = rng.integers(0, len(dims), n).astype(np.int32)
    wd = np.array([d[0] for d in dims])[ids]
    hd = np.array([d[1] for d in dims])[ids]
    a["px"] = rng.integers(-30, W + 30, n)
    a["py"] = rng.integers(-30, 160, n)
    a["pz"] = rng.integers(-60, L + 60, n)
    a["ex"] = rng.integers(1, wd + 1)
    a["ey"] = rng.integers(0, hd + 1)
    a["ez"] = hd - a["ey"] - rng.integers(0, np.maximum(hd - a["ey"], 0) + 1)
    l = np.zeros(n_lights, LIGHT)
    l["x"] = rng.integers(-100, W + 200, n_lights)
    l["y"] = rng.integers(-50, 400, n_lights)
    l["z"] = rng.integers(-100, L + 100, n_lights)
    return a, l, ids


@pytest.mark.parametrize("dims", [[(16, 32)], [(32, 48)], [(16, 32), (32, 48), (20, 40), (7, 90), (64, 5)]])
def test_sized_sprites(par, oracle, dims):
    """Sprites of their own width x height (par_set_atlas_sized; the reference hard-codes 20 x 40,
    alternative.cpp:328-332): texel = row * width + column, boxes up to the sprite's size."""
    rng = np.random.default_rng(sum(w * 131 + h for w, h in dims))
    W, H, L = 640, 480, 480
    sprites = _sized_sprites(rng, dims, 6)
    pal = _random_atlas(rng, 1, 6)[1]
    boxes, lights, ids = _sized_scene(rng, W, H, L, 1800, dims, 4)
    sized = oracle.ragged_atlas(sprites)
    ref = oracle.render(W, H, L, np.ascontiguousarray(boxes, oracle.AABB), np.ascontiguousarray(lights, oracle.LIGHT),
                        palette=np.ascontiguousarray(pal, oracle.COLOR), sprite_ids=ids, sized_atlas=sized)
    with par.Renderer(W, H, L) as r:
        r.set_atlas_sized(*sized, palette=pal)
        r.set_scene(boxes, ids)
        rgba, gbuf, stats = r.render(lights, want_gbuf=True)
        _, texel = r.gbuffer()
    assert (ref["texel"] >= 0).mean() > 0.2
    _assert_frame_equal((rgba, gbuf, texel, stats), ref)


def test_sized_sprite_padding_invariance(par, oracle):
    """The 20x40 sprite embedded in a 32x48 one (extra columns / rows are never indexed by boxes of
    extent <= 20 / 40) renders the default scene to the REAL reference's frame: pins the width
    generalisation of the texel index to the reference."""
    from conftest import sha256
    base = par.tile_floor()[0]
    color = np.full((48, 32), 3, np.int32)
    depth = np.full((48, 32), 7, np.int32)
    normal = np.ones((48, 32, 3), np.float32)
    color[:40, :20] = base["color"].reshape(40, 20)
    depth[:40, :20] = base["depth"].reshape(40, 20)
    normal[:40, :20] = base["normal"].reshape(40, 20, 3)
    W, H, L = 480, 320, 320
    with par.Renderer(W, H, L) as r:
        r.set_atlas_sized([32], [48], color.reshape(-1), depth.reshape(-1), normal.reshape(-1, 3))
        r.set_scene(par.scene_default())
        rgba, gbuf, _ = r.render(par.light_default(), want_gbuf=True)
    import json, os
    from conftest import ROOT
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_hashes.json")))["tier1_480x320x320_frame0"]
    assert sha256(rgba) == g["frame0_pre_overlay_sha256"]
    assert sha256(gbuf) == g["gbuf0_sha256"]


def test_sized_atlas_validation(par):
    from par_b200 import AABB
    with par.Renderer(480, 320, 320) as r:
        with pytest.raises(par.ParError) as e:
            r.set_atlas_sized([0], [40], np.zeros(0), np.zeros(0), np.zeros((0, 3)))
        assert e.value.code == -1
        with pytest.raises(par.ParError) as e:
            r.set_atlas_sized([4], [4], np.zeros(16), np.full(16, 5000), np.zeros((16, 3)))
        assert e.value.code == -1  # depth out of range
        r.set_atlas_sized([16], [32], np.zeros(512), np.zeros(512), np.zeros((512, 3)))
        r.set_scene(np.array([(10, 0, 10, 17, 10, 10, (0, 0))], AABB))  # extent.x 17 > width 16
        with pytest.raises(par.ParError) as e:
            r.render(par.light_default())
        assert e.value.code == -5 and "entity 0" in str(e.value)
        r.set_scene(np.array([(10, 0, 10, 16, 20, 13, (0, 0))], AABB))  # ey + ez = 33 > height 32
        with pytest.raises(par.ParError):
            r.render(par.light_default())
        r.set_scene(np.array([(10, 0, 10, 16, 20, 12, (0, 0))], AABB))
        r.render(par.light_default())


def test_culled_entities_are_never_validated(par, oracle):
    """An entity the reference culls (alternative.cpp:212-219) never has its sprite indexed, so neither
    its extents nor its sprite id can make the scene bad — only inserted boxes are validated."""
    from par_b200 import AABB
    W, H, L = 480, 320, 320
    boxes = par.scene_default()[:3000].copy()
    junk = np.array([(5000, 0, 10, 300, 500, 500, (0, 0)),      # off-screen to the right, absurd extents
                     (10, 0, 9000, 25, 30, 30, (0, 0))], AABB)   # far beyond the view length
    scene = np.concatenate([boxes, junk])
    ids = np.zeros(len(scene), np.int32)
    ids[-2:] = 77  # sprite ids outside the atlas
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(scene, ids)
        rgba, _ = r.render(par.light_default())
    ref = oracle.render(W, H, L, np.ascontiguousarray(boxes, oracle.AABB), oracle.light_default())
    assert np.array_equal(_u32(rgba), _u32(ref["rgba"]))


# ------------------------------------------------------------------ round 2: incremental update, pitch, resident frames

def _grid_and_frame(r, lights):
    rgba, gbuf, st = r.render(lights, want_gbuf=True)
    count, ids = r.grid()
    return rgba, gbuf, count, ids, st


def test_incremental_update_equals_full_rebuild(par, oracle):
    """par_update_entities patches the resident scene: after every update the grid (counts + slot
    order), the G-buffer, the frame and the loader's counters equal those of a fresh par_set_scene of
    the same boxes — for the reference's own motion (entity 0 under key script D), for entities
    leaving and re-entering the view, for several entities at once, for overflowing bins, for a
    sprite change and for more than PAR_MAX_UPDATE entities (device-side full re-bin)."""
    from par_b200 import AABB
    W, H, L = 640, 480, 480
    rng = np.random.default_rng(99)
    atlas, pal = _random_atlas(rng, 2, 4)
    boxes, lights = par.scene_default().copy(), par.light_default()
    dense, _, _ = _random_scene(rng, W, H, L, 6000, 1, cubes=True)  # wraps rings (quirk Q2)
    boxes[1000:7000] = dense
    ids = np.zeros(len(boxes), np.int32)
    with par.Renderer(W, H, L) as inc, par.Renderer(W, H, L) as full:
        for r in (inc, full):
            r.set_atlas(atlas, pal)
        inc.set_scene(boxes, ids)

        def check(tag):
            full.set_scene(boxes, ids)
            a, b = _grid_and_frame(inc, lights), _grid_and_frame(full, lights)
            assert np.array_equal(a[2], b[2]), f"{tag}: counts"
            live = np.arange(8)[None, :] < a[2][:, None]
            assert np.array_equal(np.where(live, a[3], -1), np.where(live, b[3], -1)), f"{tag}: slots"
            assert a[1].tobytes() == b[1].tobytes(), f"{tag}: G-buffer"
            assert np.array_equal(_u32(a[0]), _u32(b[0])), f"{tag}: frame"
            assert (a[4]["n_survivors"], a[4]["n_inserts"]) == (b[4]["n_survivors"], b[4]["n_inserts"]), f"{tag}: counters"

        check("initial")
        for f in range(1, 40):  # the reference's motion: entity 0 walks, light 0 drifts
            for k in oracle.script_keys("D", f):
                par.apply_key(k, boxes, lights)
            inc.update_entities(0, boxes[0:1])
            if f % 6 == 0:
                check(f"script D frame {f}")
        for step, pos in enumerate([(-500, 0, 100), (100, 20, 100), (100, 20, 9000), (300, 40, 200), (300, 40, 200)]):
            boxes[0]["px"], boxes[0]["py"], boxes[0]["pz"] = pos  # out of view, back in, out along z, back in, unchanged
            inc.update_entities(0, boxes[0:1])
            check(f"teleport {step}")
        for step in range(4):  # several entities, inside the dense part (ring wrap) and at the end of the scene
            first = [1200, 3000, len(boxes) - 5, 6990][step]
            n = [8, 3, 5, 8][step]
            boxes["px"][first:first + n] += rng.integers(-60, 60, n).astype(np.int16)
            boxes["pz"][first:first + n] += rng.integers(-60, 60, n).astype(np.int16)
            inc.update_entities(first, boxes[first:first + n])
            check(f"multi {step}")
        ids[2000:2004] = 1  # sprite change with the boxes unchanged
        inc.update_entities(2000, boxes[2000:2004], ids[2000:2004])
        check("sprites")
        boxes["py"][4000:4100] += 15  # a big update: uploaded and re-binned on the device
        inc.update_entities(4000, boxes[4000:4100])
        check("big")
        with pytest.raises(par.ParError) as e:
            inc.update_entities(len(boxes) - 1, boxes[0:2])
        assert e.value.code == -1
        bad = boxes[0:1].copy()
        bad["ex"] = 30
        inc.update_entities(0, bad)
        with pytest.raises(par.ParError) as e:
            inc.render(lights)
        assert e.value.code == -5


def test_pipelined_updates_moving_scene(par, oracle):
    """par_submit_update: 30 frames of key script D with 16 bytes of scene traffic per frame; every
    frame equals the synchronous par_set_scene + par_render of the same scene."""
    W, H, L = 640, 480, 480
    boxes, lights = par.scene_default(), par.light_default()
    scenes, light_seq = [], []
    for f in range(30):
        for k in oracle.script_keys("D", f):
            par.apply_key(k, boxes, lights)
        scenes.append(boxes[0:1].copy())
        light_seq.append(lights.copy())
    want = []
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        full = par.scene_default()
        for f in range(30):
            full[0] = scenes[f][0]
            r.set_scene(full)
            want.append(r.render(light_seq[f])[0].copy())
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(par.scene_default())
        out = [par.pinned_empty((H, W), par.COLOR) for _ in range(2)]
        r.set_cursor(0, 0)
        for f in range(31):
            if f < 30:
                r.submit_update(0, scenes[f], light_seq[f], out[f & 1])
            if f >= 1:
                st = r.wait_frame()
                assert st["n_survivors"] > 0 and st["rays"] == W * H * 2
                assert np.array_equal(_u32(out[(f - 1) & 1]), _u32(want[f - 1])), f"frame {f - 1}"
        r.submit_update(0, scenes[0][:0], light_seq[3], out[0])  # nothing moves, only the light
        st = r.wait_frame()
        assert st["n_survivors"] > 0
        full[0] = scenes[29][0]
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(full)
        assert np.array_equal(_u32(out[0]), _u32(r.render(light_seq[3])[0]))


@pytest.mark.parametrize("stripes", [1, 3])
def test_pitched_host_frames(par, stripes):
    """The blit contract of alternative.cpp:774-788: rows of the host frame `pitch` bytes apart
    (par_set_output_pitch for par_render / par_submit_frame / par_read_stripes, par_read_frame_pitched);
    bytes between the rows stay untouched."""
    W, H, L = 640, 480, 480
    boxes, lights = par.scene_synthetic(W, H, L, n=1500, n_lights=3)
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(boxes)
        want, _ = r.render(lights)
    pitch = W * 4 + 192
    got = np.zeros((H, pitch), np.uint8)
    pads = []
    for i in range(stripes):
        kw = dict(stripe_count=stripes, stripe_index=i) if stripes > 1 else {}
        with par.Renderer(W, H, L, **kw) as r:
            r.set_atlas()
            r.set_scene(boxes)
            r.set_output_pitch(pitch)
            buf = par.pinned_empty((H, pitch), np.uint8)
            buf[:] = 0xAB
            r.render(lights, out=buf)                    # blocking call, pitched
            sub = par.pinned_empty((H, pitch), np.uint8)
            sub[:] = 0xAB
            hb = par.pinned_empty(len(boxes), par.AABB)
            hb[:] = boxes
            lib = par.lib()
            assert lib.par_submit_frame(r._h, hb.ctypes.data, None, len(hb), np.ascontiguousarray(lights).ctypes.data,
                                        len(lights), sub.ctypes.data) == 0
            r.wait_frame()
            assert np.array_equal(buf, sub)
            own = np.zeros(H, bool)
            for t in range(i, H // 40, stripes):
                own[t * 40:t * 40 + 40] = True
            got[own] = buf[own]
            assert (buf[~own] == 0xAB).all() and (buf[:, W * 4:] == 0xAB).all()
            if stripes == 1:
                whole = par.pinned_empty((H, pitch), np.uint8)
                whole[:] = 0xCD
                r.read_frame_pitched(whole, pitch)
                r.sync()
                assert np.array_equal(whole[:, :W * 4], want.view(np.uint8).reshape(H, W * 4))
                assert (whole[:, W * 4:] == 0xCD).all()
                with pytest.raises(par.ParError):
                    r.set_output_pitch(W * 4 - 4)
    assert np.array_equal(got[:, :W * 4], want.view(np.uint8).reshape(H, W * 4))


@pytest.mark.parametrize("tile_order", [-1, 1])
def test_resident_frames_graph_replay(par, tile_order):
    """par_render_resident: loader + render kernel replayed as CUDA graphs (one per grid generation),
    with and without the longest-first tile order; every frame equals par_render, also after the
    lights, the scene or the cursor change in between."""
    W, H, L = 1280, 720, 720
    boxes, lights = par.scene_synthetic(W, H, L, n=3000, n_lights=6)
    with par.Renderer(W, H, L, tile_order=tile_order) as r:
        r.set_atlas()
        r.set_scene(boxes)
        want, _ = r.render(lights)
        for k in range(7):
            r.render_resident(lights)
            got = r.read_frame()
            r.sync()
            assert np.array_equal(_u32(got), _u32(want)), f"replay {k}"
        lights2 = lights.copy()
        lights2["x"] += 40
        want2, _ = r.render(lights2)
        r.set_cursor(100, 100)
        for k in range(4):
            r.render_resident(lights2)
            got = r.read_frame()
            r.sync()
            assert np.array_equal(_u32(got), _u32(want2)), f"lights changed, replay {k}"
        moved = boxes.copy()
        moved["px"][:50] += 33
        r.update_entities(0, moved[:4])
        r.set_scene(moved)
        want3, gbuf3, _ = r.render(lights2, want_gbuf=True)
        for k in range(4):
            r.render_resident(lights2)
            got = r.read_frame()
            r.sync()
            assert np.array_equal(_u32(got), _u32(want3)), f"scene changed, replay {k}"
        assert r.cursor_pixel().tobytes() == gbuf3[100, 100].tobytes()
        assert r.stats()["n_survivors"] > 0


@pytest.mark.parametrize("flags", ["0", "128"])
def test_resident_frames_interleaved_with_scene_calls(par, oracle, monkeypatch, flags):
    """The overlapped resident frame renders from one grid generation while the other is rebuilt beside it
    (PAR_DEBUG_FLAGS=128: loader in front of the kernel instead).  Whatever comes in between — incremental
    updates, a new scene, a rebuild, blocking renders, pipelined frames, the parity checkpoints — every
    frame and every grid read back equals the oracle's for the scene resident at that moment."""
    monkeypatch.setenv("PAR_DEBUG_FLAGS", flags)
    W, H, L = 960, 680, 680
    boxes, lights = par.scene_synthetic(W, H, L, n=4000, n_lights=3)
    boxes = boxes.copy()
    rng = np.random.default_rng(7)

    def check(r, tag, resident_frames):
        ref = oracle.render(W, H, L, np.ascontiguousarray(boxes, oracle.AABB), lights.view(oracle.LIGHT))
        for k in range(resident_frames):
            r.render_resident(lights)
        got = r.read_frame()
        r.sync()
        assert np.array_equal(_u32(got), _u32(ref["rgba"])), f"{tag}: resident frame"
        count, ids = r.grid()
        ocount, _, oent = oracle.grid_build(W, H, L, np.ascontiguousarray(boxes, oracle.AABB))
        assert np.array_equal(count, ocount), f"{tag}: grid counts"
        oent = oent.reshape(-1, 8)
        live = np.arange(8)[None, :] < ocount[:, None]
        assert np.array_equal(np.where(live, ids, -1), np.where(live, oent, -1)), f"{tag}: grid entity map"
        gbuf, _ = r.gbuffer()
        assert gbuf.tobytes() == ref["gbuf"].tobytes(), f"{tag}: G-buffer"
        return ref

    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(boxes)
        check(r, "fresh scene", 1)
        check(r, "replays", 5)
        for step in range(3):  # incremental updates between resident frames (odd and even numbers of frames)
            boxes["px"][:3] += rng.integers(-30, 30, 3).astype(np.int16)
            boxes["pz"][:3] += rng.integers(-30, 30, 3).astype(np.int16)
            r.update_entities(0, boxes[:3])
            check(r, f"update {step}", 1 + step)
        boxes[100:2000] = boxes[2000:3900]  # a different scene of the same size
        r.set_scene(boxes)
        check(r, "new scene", 2)
        r.rebuild_grid()
        ref = check(r, "after rebuild", 3)
        rgba, _ = r.render(lights)  # blocking render from the grid the resident frames left behind
        assert np.array_equal(_u32(rgba), _u32(ref["rgba"]))
        check(r, "after blocking render", 1)
        h_boxes = par.pinned_empty(len(boxes), par.AABB)
        boxes["py"][:200] += 15
        h_boxes[:] = boxes
        out = par.pinned_empty((H, W), par.COLOR)
        r.submit_frame(h_boxes, lights, out)  # pipelined frame with a full upload
        r.wait_frame()
        ref = check(r, "after a pipelined frame", 2)
        assert np.array_equal(_u32(out), _u32(ref["rgba"]))
        boxes["px"][0] += 25
        r.submit_update(0, boxes[:1], lights, out)  # pipelined frame with an incremental update
        r.wait_frame()
        ref = check(r, "after a pipelined update", 3)
        assert np.array_equal(_u32(out), _u32(ref["rgba"]))
        small = boxes[:1500].copy()  # a smaller scene: the survivor lists of both generations shrink
        boxes = small
        r.set_scene(boxes)
        check(r, "smaller scene", 4)


def test_two_contexts_share_a_device(par, oracle):
    """Contexts with different views and atlases on one device do not disturb each other."""
    rng = np.random.default_rng(5)
    a_boxes, a_lights, _ = _random_scene(rng, 480, 320, 320, 900, 2)
    b_boxes, b_lights = par.scene_synthetic(3840, 2160, 2160, n=4000, n_lights=3)
    atlas, pal = _random_atlas(rng, 6, 4)
    with par.Renderer(3840, 2160, 2160) as big, par.Renderer(480, 320, 320) as small:
        big.set_atlas()
        small.set_atlas(atlas, pal)
        big.set_scene(b_boxes)
        ids = rng.integers(0, 6, len(a_boxes)).astype(np.int32)
        small.set_scene(a_boxes, ids)
        first, _ = big.render(b_lights)
        got, _ = small.render(a_lights)
        again, _ = big.render(b_lights)
    ref = oracle.render(480, 320, 320, np.ascontiguousarray(a_boxes, oracle.AABB), np.ascontiguousarray(a_lights, oracle.LIGHT),
                        atlas=np.ascontiguousarray(atlas, oracle.SPRITE), palette=np.ascontiguousarray(pal, oracle.COLOR),
                        sprite_ids=ids)
    assert np.array_equal(_u32(got), _u32(ref["rgba"]))
    assert np.array_equal(_u32(first), _u32(again))


def test_many_groups_and_long_columns(par, oracle):
    """Tiles whose pixels spread over many start bins (columns of floating cubes at all depths, more
    groups than one pass handles is not reachable at this view, but > 30 are) and bin columns with
    more entries than one staging chunk holds."""
    from par_b200 import AABB
    W, H, L = 200, 2000, 4000
    rows = []
    rng = np.random.default_rng(17)
    for k in range(900):  # cubes stacked along z with rising y: every tile sees dozens of depths
        rows.append((int(rng.integers(0, 180)), int(rng.integers(0, 1500)), int(rng.integers(0, 3900)), 20, 20, 20, (0, 0)))
    for k in range(700):  # one bin column crowded along z (7 kept per bin x 100 bins > 256 entries)
        rows.append((40 + int(rng.integers(0, 20)), 20 * int(rng.integers(0, 3)), 40 * (k % 100) + int(rng.integers(0, 20)), 20, 20, 20, (0, 0)))
    boxes = np.array(rows, AABB)
    lights = np.zeros(3, par.LIGHT)
    lights["x"], lights["y"], lights["z"] = [100, 10, 190], [800, 300, 1500], [500, 2500, 100]
    got, ref = _render_both(par, oracle, W, H, L, boxes, lights)
    _assert_frame_equal(got, ref)


@pytest.mark.timeout(180)
@pytest.mark.parametrize("root,split", [(-1, 1), (0, 1), (-1, 2), (0, 2), (1, 2)])
def test_resident_flag_exchange_two_contexts_one_device(par, root, split):
    """The multi-GPU step of bench.py (par_exchange_setup + par_render_resident: the render kernels store
    their stripes into the consumers' frames, arrival / credit flags in the frame footers order the frames,
    the whole step one replayed CUDA graph) with both ranks' contexts on ONE device, so that a 1-GPU box
    exercises it too: whole tile rows and stripe_split = 2 over an odd number of tile rows (17), all-gather
    and gather-to-root.  Replayed frames, frames after a light change (re-captured graphs) and after an entity
    update must leave the one-context frame on every consumer."""
    W, H, L = 1280, 680, 680
    boxes, lights = par.scene_synthetic(W, H, L, n=3000, n_lights=6)
    kw = dict(stripe_split=split) if split > 1 else {}
    with par.Renderer(W, H, L) as ref, \
            par.Renderer(W, H, L, stripe_count=2, stripe_index=0, **kw) as a, \
            par.Renderer(W, H, L, stripe_count=2, stripe_index=1, **kw) as b:
        ref.set_atlas()
        a.peer_set(1, b.device_frame())
        b.peer_set(0, a.device_frame())
        for r in (a, b):
            r.set_atlas()
            r.set_scene(boxes)
            r.exchange_setup(root)

        def check(tag):
            ref.set_scene(boxes)
            want, _ = ref.render(lights)
            for r in (a, b):
                r.sync()
            for idx, r in enumerate((a, b)):
                if root in (-1, idx):
                    got = r.read_frame()
                    r.sync()
                    assert np.array_equal(_u32(want), _u32(got)), f"{tag}: rank {idx}"

        for _ in range(7):  # first frames run plainly / are captured, the later ones are graph replays
            a.render_resident(lights)
            b.render_resident(lights)
        check("replayed frames")
        lights = lights.copy()
        lights["x"] += 35
        lights["z"] += 20
        for _ in range(3):
            b.render_resident(lights)  # the enqueue order between the ranks must not matter
            a.render_resident(lights)
        check("after moving the lights")
        boxes = boxes.copy()
        boxes["px"][:4] += 60
        for r in (a, b):
            r.update_entities(0, boxes[:4])
        for _ in range(2):
            a.render_resident(lights)
            b.render_resident(lights)
        check("after an entity update")


def _dense_scene(seed, W=1200, H=440, L=440, n=9000, n_lights=12):
    from par_b200 import AABB, LIGHT
    rng = np.random.default_rng(seed)
    boxes = np.zeros(n, AABB)
    boxes["px"] = rng.integers(0, W - 20, n)
    boxes["py"] = rng.integers(0, 260, n)
    boxes["pz"] = rng.integers(0, L - 20, n)
    boxes["ex"] = boxes["ey"] = boxes["ez"] = 20
    lights = np.zeros(n_lights, LIGHT)
    lights["x"] = rng.integers(-2500, 4000, n_lights)
    lights["y"] = rng.integers(20, 700, n_lights)
    lights["z"] = rng.integers(-1500, 2500, n_lights)
    k = min(4, n_lights)
    lights["x"][:k] = rng.integers(0, W, k)  # some inside the view
    lights["z"][:k] = rng.integers(0, L, k)
    return boxes, lights


@pytest.mark.parametrize("flags", ["256", "512"])
@pytest.mark.parametrize("n_lights", [1, 12])
def test_both_kernel_configurations_vs_oracle(par, oracle, monkeypatch, flags, n_lights):
    """The render kernel is built twice (tile.cu / tile_one_light.cu: 5 CTAs per SM with large shared lists,
    6 CTAs per SM with the smallest ones) and the library picks per frame (one light and many CTA waves ->
    the second).  PAR_DEBUG_FLAGS 256 / 512 pin the first / second: a dense scene whose rounds overflow the
    lists (re-walked, cut into step ranges) must give the oracle's frame on either, production frames
    (no G-buffer checkpoint: that is the debug instantiation) and graph-replayed resident frames alike."""
    W, H, L = 1200, 440, 440
    boxes, lights = _dense_scene(21, W, H, L, n_lights=n_lights)
    ref = oracle.render(W, H, L, boxes.view(oracle.AABB), lights.view(oracle.LIGHT), want_gbuf=False, want_texel=False)
    monkeypatch.setenv("PAR_DEBUG_FLAGS", flags)
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(boxes)
        rgba, _ = r.render(lights)
        assert np.array_equal(_u32(rgba), _u32(ref["rgba"]))
        for _ in range(3):
            r.render_resident(lights)
        got = r.read_frame()
        r.sync()
        assert np.array_equal(_u32(got), _u32(ref["rgba"]))


@pytest.mark.parametrize("name", ["c2", "c3", "c5b"])
def test_kernel_configurations_agree_at_full_size(par, oracle, workload_ops, monkeypatch, name):
    """Full-size BASELINE frames on the pinned configurations (PAR_DEBUG_FLAGS 256 / 512) and on the automatic
    choice: the committed oracle frame hash every time."""
    g = workload_ops[name]
    W, H, L = g["view"]
    boxes, lights = _workload(par, name)
    for flags in ("256", "512", "0"):
        monkeypatch.setenv("PAR_DEBUG_FLAGS", flags)
        with par.Renderer(W, H, L) as r:
            r.set_atlas()
            r.set_scene(boxes)
            rgba, _ = r.render(lights)
            assert "%016x" % oracle.fnv1a64(rgba) == g["frame_fnv1a64"], f"PAR_DEBUG_FLAGS={flags}"
