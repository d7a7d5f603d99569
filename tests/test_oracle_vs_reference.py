"""Pin the oracle against the REAL reference.

The reference ships no tests or golden vectors, so the goldens are hashes of what the
reference executable itself renders (tests/golden/reference_hashes.json, written by
tests/golden/make_reference_goldens.py from the headless builds of oracle/build_ref.sh:
tier 0 = unmodified source, tier 1 = view size parameterised with sed).
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, decode_scene, sha256

REF = os.path.join(ROOT, "oracle", "_ref")


def _frame(O, W, H, L, boxes, lights):
    r = O.render(W, H, L, boxes, lights)
    post = r["rgba"].copy()
    O.draw_overlay(W, H, L, r["gbuf"], lights, post)  # alternative.cpp:762-772
    return r, post


@pytest.mark.parametrize("tier", ["tier0_480x320x320_frame0", "tier1_480x320x320_frame0"])
def test_c1_frame_and_gbuffer(oracle, golden, tier):
    O = oracle
    g = golden[tier]
    r, post = _frame(O, 480, 320, 320, O.scene_default(), O.light_default())
    assert sha256(post) == g["frame0_sha256"]
    assert "%016x" % O.fnv1a64(post) == g["fnv1a64"][0]
    if "gbuf0_sha256" in g:
        assert sha256(r["rgba"]) == g["frame0_pre_overlay_sha256"]
        assert sha256(r["gbuf"]) == g["gbuf0_sha256"]  # raw Pixel[] bytes, sprites.hpp:53-58
        for name in ("entity", "y", "z"):
            assert sha256(r["gbuf"][name].astype("<i4")) == g[f"gbuf0_{name}_sha256"]


def test_c1_counters_match_survey(oracle):
    """SURVEY.md §8(d) raw C1 counter values, measured on the instrumented reference."""
    O = oracle
    c = O.render(480, 320, 320, O.scene_default(), O.light_default())["counters"]
    assert c == {"pixels": 153600, "primary_bins": 1228800, "primary_slot_tests": 1329600,
                 "primary_passed": 446800, "primary_accepts": 199200,
                 "shaded_px_lights": 153600, "lit_px_lights": 130682, "shadow_probes": 6040904,
                 "shadow_slot_entries": 1576841, "slab_tests": 1568629, "pixels_hit": 150400}


def test_script_c_240_frames(oracle, golden):
    """240-frame key script C on the UNMODIFIED reference: every frame's FNV-1a-64."""
    O = oracle
    want = golden["tier0_480x320x320_scriptC_240"]["fnv1a64"]
    boxes, lights = O.scene_default(), O.light_default()
    for f in range(240):
        for k in O.script_keys("C", f):
            O.apply_key(k, boxes, lights)
        _, post = _frame(O, 480, 320, 320, boxes, lights)
        assert "%016x" % O.fnv1a64(post) == want[f], f"frame {f}"


def test_1080p_frame0(oracle, golden):
    O = oracle
    g = golden["tier1_1920x1080x1080_frame0"]
    r, post = _frame(O, 1920, 1080, 1080, O.scene_default(), O.light_default())
    assert sha256(post) == g["frame0_sha256"]
    assert sha256(r["rgba"]) == g["frame0_pre_overlay_sha256"]
    assert sha256(r["gbuf"]) == g["gbuf0_sha256"]


def test_4k_frame0(oracle, golden):
    O = oracle
    g = golden["tier1_3840x2160x2160_frame0"]
    r, post = _frame(O, 3840, 2160, 2160, O.scene_default(), O.light_default())
    assert sha256(post) == g["frame0_sha256"]
    assert sha256(r["rgba"]) == g["frame0_pre_overlay_sha256"]
    assert sha256(r["gbuf"]) == g["gbuf0_sha256"]
    # SURVEY.md §8(d) raw C2 counters
    c = r["counters"]
    assert (c["primary_bins"], c["primary_slot_tests"], c["primary_passed"],
            c["primary_accepts"]) == (447897600, 50488000, 16854400, 8389600)
    assert (c["lit_px_lights"], c["shadow_probes"], c["shadow_slot_entries"], c["slab_tests"],
            c["pixels_hit"]) == (8205245, 2524893691, 789402551, 789394339, 8291200)


def test_script_d_1080p_sampled(oracle, golden):
    """C4's script D (player + light move): a sample of frames by default, all 240 with
    PAR_SLOW=1 (the light leaves x = 480, so Q18's out-of-grid light bin is exercised
    differently every frame)."""
    O = oracle
    want = golden["tier1_1920x1080x1080_scriptD_240"]["fnv1a64"]
    check = range(240) if os.environ.get("PAR_SLOW") else {0, 1, 30, 60, 120, 180, 239}
    boxes, lights = O.scene_default(), O.light_default()
    for f in range(240):
        for k in O.script_keys("D", f):
            O.apply_key(k, boxes, lights)
        if f in check:
            _, post = _frame(O, 1920, 1080, 1080, boxes, lights)
            assert "%016x" % O.fnv1a64(post) == want[f], f"frame {f}"


def test_sprite_table_equals_reference(oracle):
    """make_tile_floor() bytes straight from the reference header (needs oracle/_ref)."""
    tool = os.path.join(REF, "ref_sprite_dump")
    if not os.path.exists(tool):
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    raw = subprocess.run([tool], check=True, capture_output=True).stdout
    assert len(raw) == 16016
    assert oracle.tile_floor().tobytes() == raw[:16000]
    assert oracle.default_palette().tobytes() == raw[16000:]


def test_live_tier0_if_present(oracle, tmp_path):
    """Run the unmodified reference here and compare 3 scripted frames byte for byte."""
    exe = os.path.join(REF, "ref_tier0")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    O = oracle
    env = dict(os.environ, PAR_REF_FRAMES="3", PAR_REF_SCRIPT="D",
               PAR_REF_DUMP=str(tmp_path / "f.rgba"), PAR_REF_DUMP_FRAMES="0,1,2")
    subprocess.run([exe], env=env, check=True, stdout=subprocess.DEVNULL)
    frames = np.fromfile(tmp_path / "f.rgba", dtype=O.COLOR).reshape(3, 320, 480)
    boxes, lights = O.scene_default(), O.light_default()
    for f in range(3):
        for k in O.script_keys("D", f):
            O.apply_key(k, boxes, lights)
        _, post = _frame(O, 480, 320, 320, boxes, lights)
        assert np.array_equal(post.view(np.uint32), frames[f].view(np.uint32)), f"frame {f}"


# ---- arbitrary scenes rendered by the REAL reference through its scene hook -------------------

def _check_against_reference_entry(O, entry, boxes, lights):
    W, H, L = entry["view"]
    for f in range(entry["frames"]):
        if entry["script"]:
            for k in O.script_keys(entry["script"], f):
                O.apply_key(k, boxes, lights)
        r, post = _frame(O, W, H, L, boxes, lights)
        if f == 0:
            assert sha256(r["gbuf"]) == entry["gbuf0_sha256"], "G-buffer"
            assert sha256(r["rgba"]) == entry["frame0_pre_overlay_sha256"], "frame before the overlay"
        assert "%016x" % O.fnv1a64(post) == entry["fnv1a64"][f], f"frame {f}"


@pytest.mark.parametrize("k", range(18))
def test_random_scenes_vs_real_reference(oracle, golden_scenes, k):
    """Random / ragged / lattice-snapped scenes, the light free or on a box face, at views with
    length != height too: the oracle's G-buffer, shaded frame and scripted final frames equal the
    real reference's.  (One light: the reference shades with lights[0] only.)"""
    if k >= len(golden_scenes["scenes"]):
        pytest.skip("fewer scenes in the golden file (the real reference crashed on some)")
    entry = golden_scenes["scenes"][k]
    boxes, lights = decode_scene(entry, oracle.AABB, oracle.LIGHT)
    assert len(boxes) == entry["n_boxes"] and len(lights) == entry["n_lights"]
    _check_against_reference_entry(oracle, entry, boxes, lights)


@pytest.mark.parametrize("view", ["480x320x320", "480x320x640", "640x480x200", "200x40x40", "40x1000x120"])
def test_default_scene_other_views_light_inside(oracle, golden_scenes, view):
    """The built-in scene (scene constants 480/320/320) through views whose length differs from
    their height, the light moved inside each view's grid, 3 frames of key script D."""
    entry = golden_scenes["default_scene"].get(view)
    if entry is None:
        pytest.skip("the real reference crashed on this configuration when the goldens were made")
    lights = oracle.light_default()
    lights[0]["x"], lights[0]["y"], lights[0]["z"], lights[0]["radius"] = entry["light"]
    _check_against_reference_entry(oracle, entry, oracle.scene_default(), lights)
