"""Host-side pieces of the product (libpar_b200.so, no GPU needed) against the oracle's
independent restatement: sprite table, palette, default scene, synthetic recipe, key map,
debug overlay.  These are the INPUTS of the render path."""
import numpy as np

from conftest import sha256


def test_tile_floor_and_palette(par, oracle):
    assert par.tile_floor().tobytes() == oracle.tile_floor().tobytes()
    assert par.default_palette().tobytes() == oracle.default_palette().tobytes()


def test_default_scene(par, oracle):
    a, b = par.scene_default(), oracle.scene_default()
    assert len(a) == 162308  # SURVEY.md §2, probe-confirmed on the reference
    assert a.tobytes() == b.tobytes()
    assert par.light_default().tobytes() == oracle.light_default().tobytes()
    assert tuple(par.light_default()[0])[:3] == (480, 160, 80)  # alternative.cpp:626


def test_synthetic_recipe_check_values(par, oracle):
    """SURVEY.md §8(d) check values for the splitmix64 recipe."""
    a, l = par.scene_synthetic(3840, 2160, 2160)
    assert (a[0]["px"], a[0]["py"], a[0]["pz"]) == (3579, 128, 1376)
    assert (a[1]["px"], a[1]["py"], a[1]["pz"]) == (3367, 28, 781)
    assert (l[0]["x"], l[0]["y"], l[0]["z"]) == (1044, 295, 163)
    a5, l5 = par.scene_synthetic(7680, 4320, 4320)
    assert (a5[0]["px"], a5[0]["py"], a5[0]["pz"]) == (7419, 128, 456)
    assert (l5[0]["x"], l5[0]["y"], l5[0]["z"]) == (1044, 295, 3163)
    oa, ol = oracle.scene_synthetic(3840, 2160, 2160)
    assert a.tobytes() == oa.tobytes() and l.tobytes() == ol.tobytes()


def test_key_map(par, oracle):
    for key in "LRUDpPakjuhoX":
        a, l = par.scene_default()[:1].copy(), par.light_default()
        b, m = a.copy(), l.copy()
        par.apply_key(key, a, l)
        oracle.apply_key(key, b, m)
        assert a.tobytes() == b.tobytes() and l.tobytes() == m.tobytes(), key


def test_overlay(par, oracle, golden):
    """alternative.cpp:762-772: product overlay on an oracle frame reproduces the
    unmodified reference's frame hash."""
    O = oracle
    boxes, lights = O.scene_default(), O.light_default()
    r = O.render(480, 320, 320, boxes, lights)
    frame = r["rgba"].copy()
    par.draw_overlay(480, 320, r["gbuf"], lights, frame)
    assert sha256(frame) == golden["tier0_480x320x320_frame0"]["frame0_sha256"]
    # off-screen light and a different cursor: same pixels as the oracle's overlay
    lights2 = lights.copy()
    lights2[0]["x"], lights2[0]["y"], lights2[0]["z"] = 700, -50, 30
    f1, f2 = r["rgba"].copy(), r["rgba"].copy()
    par.draw_overlay(480, 320, r["gbuf"], lights2, f1, 100, 200)
    O.draw_overlay(480, 320, 320, r["gbuf"], lights2, f2, 100, 200)
    assert np.array_equal(f1.view(np.uint32), f2.view(np.uint32))
    # the single-record form (cursor probe, mouse_pixel of alternative.cpp:380-382) draws the same line
    f3 = r["rgba"].copy()
    par.draw_overlay_at(480, 320, r["gbuf"][200, 100], lights2, f3, 100)
    assert np.array_equal(f1.view(np.uint32), f3.view(np.uint32))
    f4 = r["rgba"].copy()
    par.draw_overlay_at(480, 320, r["gbuf"][0, 0], lights, f4, 0)
    assert sha256(f4) == golden["tier0_480x320x320_frame0"]["frame0_sha256"]
