"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): the single-process par_multi_*
mode (row bands + in-place ncclAllGather) must reproduce the one-GPU frame byte for byte; and
the headless C++ driver must print the real reference's per-frame hashes."""
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("n", [2, 4, 8])
def test_multi_equals_single(par, n):
    if _n_gpus() < n:
        pytest.skip(f"needs {n} GPUs")
    W, H, L = 1280, 720, 720
    boxes, lights = par.scene_synthetic(W, H, L, n=3000, n_lights=6)
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(boxes)
        one, _ = r.render(lights)
    with par.MultiRenderer(W, H, L, list(range(n))) as m:
        m.set_atlas()
        m.set_scene(boxes)
        many, st = m.render(lights)
        again, _ = m.render(lights)
    assert np.array_equal(one.view(np.uint32), many.view(np.uint32))
    assert np.array_equal(one.view(np.uint32), again.view(np.uint32))
    assert st["rays"] == W * H * 7


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_multi_device_resident_frame_complete_everywhere(par, monkeypatch, exchange):
    """par_multi_render(out = NULL): the frame is completed in HBM on EVERY device — by the
    peer-memory stores fused into the shade kernel, or by the NCCL all-gather fallback."""
    n = min(_n_gpus(), 4)
    if n < 2:
        pytest.skip("needs 2 GPUs")
    if exchange == "nccl":
        monkeypatch.setenv("PAR_MULTI_EXCHANGE", "nccl")
    W, H, L = 1280, 720, 720
    boxes, lights = par.scene_synthetic(W, H, L, n=3000, n_lights=6)
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(boxes)
        one, _ = r.render(lights)
    with par.MultiRenderer(W, H, L, list(range(n))) as m:
        m.set_atlas()
        m.set_scene(boxes)
        for _ in range(2):  # twice: the second frame overwrites a complete one
            m.render_resident(lights)
        for i in range(n):
            assert np.array_equal(one.view(np.uint32), m.device_frame_copy(i).view(np.uint32)), f"device {i}"
        host, _ = m.render(lights)  # host consumer afterwards: parallel stripe readback
    assert np.array_equal(one.view(np.uint32), host.view(np.uint32))


def test_fused_peer_exchange_two_contexts(par):
    """par_render_device_peers: two striped contexts on two GPUs write into each other's frames;
    after both finish, BOTH frames equal the one-GPU frame."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    W, H, L = 1280, 720, 720
    boxes, lights = par.scene_synthetic(W, H, L, n=3000, n_lights=6)
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(boxes)
        one, _ = r.render(lights)
    with par.Renderer(W, H, L, device=0, stripe_count=2, stripe_index=0) as a, \
            par.Renderer(W, H, L, device=1, stripe_count=2, stripe_index=1) as b:
        a.peer_set(1, b.device_frame())
        b.peer_set(0, a.device_frame())
        for r in (a, b):
            r.set_atlas()
            r.set_scene(boxes)
        for r in (a, b):
            r.render_device_peers(lights)
        a.sync()
        b.sync()
        fa, fb = a.read_frame(), b.read_frame()
        a.sync()
        b.sync()
    assert np.array_equal(one.view(np.uint32), fa.view(np.uint32))
    assert np.array_equal(one.view(np.uint32), fb.view(np.uint32))


def test_fused_gather_to_root(par):
    """Only rank 0's frame is imported (by rank 1): the frame completes on rank 0 alone."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    W, H, L = 1280, 720, 720
    boxes, lights = par.scene_synthetic(W, H, L, n=3000, n_lights=6)
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(boxes)
        one, _ = r.render(lights)
    with par.Renderer(W, H, L, device=0, stripe_count=2, stripe_index=0) as a, \
            par.Renderer(W, H, L, device=1, stripe_count=2, stripe_index=1) as b:
        b.peer_set(0, a.device_frame())
        for r in (a, b):
            r.set_atlas()
            r.set_scene(boxes)
            r.render_device_peers(lights)
        a.sync()
        b.sync()
        fa, fb = a.read_frame(), b.read_frame()
        a.sync()
        b.sync()
    assert np.array_equal(one.view(np.uint32), fa.view(np.uint32))
    own_b = np.zeros(H, bool)
    for t in range(1, H // 40, 2):
        own_b[t * 40:t * 40 + 40] = True
    assert np.array_equal(one.view(np.uint32)[own_b], fb.view(np.uint32)[own_b])
    assert not fb.view(np.uint32)[~own_b].any()


def test_multi_single_device_is_plain(par):
    W, H, L = 480, 320, 320
    with par.Renderer(W, H, L) as r:
        r.set_atlas()
        r.set_scene(par.scene_default())
        one, _ = r.render(par.light_default())
    with par.MultiRenderer(W, H, L, [0]) as m:
        m.set_atlas()
        m.set_scene(par.scene_default())
        many, _ = m.render(par.light_default())
    assert np.array_equal(one.view(np.uint32), many.view(np.uint32))


@pytest.mark.parametrize("mode", [[], ["--sync"]])
def test_headless_driver_prints_reference_hashes(par, golden, mode):
    """C++ host (Entities::insert + FrameRenderer + overlay) vs the UNMODIFIED reference's
    per-frame FNV-1a-64 hashes under key script C — pipelined frames with the cursor probe
    (default) and the blocking render_frame with the whole G-buffer (--sync)."""
    exe = os.path.join(PKG, "build", "par_headless")
    if not os.path.exists(exe):
        pytest.skip("par_headless not built")
    out = subprocess.run([exe, "--frames", "40", "--script", "C"] + mode, check=True, capture_output=True, text=True).stdout
    got = [ln.split()[1] for ln in out.splitlines()]
    assert got == golden["tier0_480x320x320_scriptC_240"]["fnv1a64"][:40]


def test_headless_driver_full_c4_sequence(par, golden):
    """Config 4 end to end in C++: all 240 frames of key script D at 1920x1080 (per-frame scene
    upload + render + overlay); the printed hash file must have the sha256 of the one the real
    reference produced."""
    import hashlib
    exe = os.path.join(PKG, "build", "par_headless")
    if not os.path.exists(exe):
        pytest.skip("par_headless not built")
    res = subprocess.run([exe, "--view", "1920", "1080", "1080", "--frames", "240", "--script", "D"],
                         check=True, capture_output=True)
    print(res.stderr.decode())  # frames/s of the sequence (shown with pytest -s / on failure)
    assert hashlib.sha256(res.stdout).hexdigest() == golden["tier1_1920x1080x1080_scriptD_240"]["hash_file_sha256"]


@pytest.mark.parametrize("extra", [["--full-upload"], ["--pitch", "2048"], ["--pitch", "2048", "--full-upload"]])
def test_headless_driver_upload_modes_and_pitched_frames(par, golden, extra):
    """The same 40 frames of key script C when the whole scene is re-sent every frame (the reference's
    shape, alternative.cpp:689-693) instead of the 16-byte update, and when the host frames have a row
    pitch wider than W * 4 (locked-texture contract, alternative.cpp:774-783): same hashes."""
    exe = os.path.join(PKG, "build", "par_headless")
    if not os.path.exists(exe):
        pytest.skip("par_headless not built")
    out = subprocess.run([exe, "--frames", "40", "--script", "C"] + extra, check=True, capture_output=True, text=True).stdout
    got = [ln.split()[1] for ln in out.splitlines()]
    assert got == golden["tier0_480x320x320_scriptC_240"]["fnv1a64"][:40]


def test_headless_driver_writes_the_frame_sequence(par, golden, tmp_path):
    """--ppm-seq: every finished frame (overlay included) lands as DIR/frame_NNN.ppm; the RGB bytes of each
    file are the frame whose FNV-1a-64 the driver printed (frame sink, alternative.cpp:774-788)."""
    exe = os.path.join(PKG, "build", "par_headless")
    if not os.path.exists(exe):
        pytest.skip("par_headless not built")
    W, H, n = 480, 320, 12
    out = subprocess.run([exe, "--frames", str(n), "--script", "C", "--ppm-seq", str(tmp_path)],
                         check=True, capture_output=True, text=True).stdout
    hashes = [ln.split()[1] for ln in out.splitlines()]
    assert hashes == golden["tier0_480x320x320_scriptC_240"]["fnv1a64"][:n]
    files = sorted(os.listdir(tmp_path))
    assert files == [f"frame_{f:03d}.ppm" for f in range(n)]
    header = f"P6\n{W} {H}\n255\n".encode()
    seen = set()
    for f in files:
        raw = open(os.path.join(tmp_path, f), "rb").read()
        assert raw.startswith(header) and len(raw) == len(header) + W * H * 3
        seen.add(raw)
        rgb = np.frombuffer(raw[len(header):], np.uint8).reshape(H, W, 3)
        assert (rgb[0, 0] == (255, 0, 0)).all()  # the red overlay line starts under the cursor (0, 0)
    assert len(seen) == n  # the player moves every frame


def test_headless_driver_png_and_gif_sinks(par, golden, tmp_path):
    """--png-seq / --gif (include/par/frame_sink.hpp, the writers are tested on the CPU by tests/test_frame_sink.py):
    the PNG of every frame decodes to the bytes of its PPM, the animated GIF holds all frames (exact when a frame
    has <= 256 colours, else within half a level of the 6x7x6 palette) — what gif.gif is to the reference."""
    Image = pytest.importorskip("PIL.Image")
    exe = os.path.join(PKG, "build", "par_headless")
    if not os.path.exists(exe):
        pytest.skip("par_headless not built")
    W, H, n = 480, 320, 8
    ppm_dir, png_dir, gif_path = tmp_path / "ppm", tmp_path / "png", tmp_path / "seq.gif"
    ppm_dir.mkdir()
    png_dir.mkdir()
    out = subprocess.run([exe, "--frames", str(n), "--script", "C", "--ppm-seq", str(ppm_dir), "--png-seq", str(png_dir),
                          "--gif", str(gif_path)], check=True, capture_output=True, text=True).stdout
    assert [ln.split()[1] for ln in out.splitlines()] == golden["tier0_480x320x320_scriptC_240"]["fnv1a64"][:n]
    gif = Image.open(gif_path)
    assert gif.n_frames == n and gif.size == (W, H)
    for f in range(n):
        ppm = np.asarray(Image.open(ppm_dir / f"frame_{f:03d}.ppm").convert("RGB"))
        png = np.asarray(Image.open(png_dir / f"frame_{f:03d}.png").convert("RGB"))
        assert np.array_equal(png, ppm)
        gif.seek(f)
        got = np.asarray(gif.convert("RGB")).astype(int)
        if len(np.unique(ppm.reshape(-1, 3), axis=0)) <= 256:
            assert np.array_equal(got, ppm)
        else:
            assert np.abs(got - ppm.astype(int)).max() <= 26


@pytest.mark.parametrize("root,split", [(-1, 1), (0, 1), (1, 1), (-1, 2), (0, 4)])
def test_resident_frames_with_flag_exchange(par, root, split):
    """par_render_resident + par_exchange_setup on two GPUs: the render kernels store their stripes into the
    consumers' frames, arrival / credit flags in the frame footers order the frames (no collective, no host
    round trip).  Frames replayed from the captured graphs, frames after a light change (graphs re-captured)
    and after an entity update must all leave the 1-GPU frame on every consumer."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    W, H, L = 1280, 720, 720
    boxes, lights = par.scene_synthetic(W, H, L, n=3000, n_lights=6)
    with par.Renderer(W, H, L) as ref, \
            par.Renderer(W, H, L, device=0, stripe_count=2, stripe_index=0, stripe_split=split) as a, \
            par.Renderer(W, H, L, device=1, stripe_count=2, stripe_index=1, stripe_split=split) as b:
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
                    assert np.array_equal(want.view(np.uint32), got.view(np.uint32)), f"{tag}: rank {idx}"

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
