"""Host logic of the multi-GPU mode on CPU: world_size-2 (and 3, unequal bands) gloo process
groups; every rank fills its row band from the oracle and the in-place gather must give the
full oracle frame on every rank."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _worker(rank, world, port, W, H, L, out_dir):
    for p in (ROOT, PKG):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from par_b200.bands import band_rows, gather_bands
    boxes, lights = O.scene_synthetic(W, H, L, n=600, n_lights=3)
    r0, r1 = band_rows(H, world, rank)
    band = O.render(W, H, L, boxes, lights, row0=r0, row1=r1, want_gbuf=False, want_texel=False)["rgba"]
    frame = torch.from_numpy(band.view(np.uint8).reshape(-1).copy())
    assert not frame[: r0 * W * 4].any() and not frame[r1 * W * 4:].any()
    gather_bands(frame, W, H, world, rank)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), frame.numpy())
    dist.destroy_process_group()


def _stripe_worker(rank, world, port, W, H, L, out_dir):
    for p in (ROOT, PKG):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from par_b200.bands import gather_stripes, owned_rows, stripes_per_rank, unstripe
    boxes, lights = O.scene_synthetic(W, H, L, n=600, n_lights=3)
    T, stripe = stripes_per_rank(H, world), 40 * W * 4
    staging = torch.zeros(world * T * stripe, dtype=torch.uint8)
    for k, (r0, r1) in enumerate(owned_rows(H, world, rank)):  # what par_render_device_striped writes
        rows = O.render(W, H, L, boxes, lights, row0=r0, row1=r1, want_gbuf=False, want_texel=False)["rgba"][r0:r1]
        staging[(rank * T + k) * stripe:(rank * T + k + 1) * stripe] = torch.from_numpy(rows.view(np.uint8).reshape(-1).copy())
    gather_stripes(staging, world, rank)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), unstripe(staging, W, H, world).numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_stripe_gather(oracle, tmp_path, world):
    """The default multi-GPU partition: interleaved stripes, stripe-major staging, in-place
    all-gather, un-stripe (8 tile rows over 3 ranks: padded staging)."""
    W, H, L = 480, 320, 320
    port = 29700 + os.getpid() % 2000 + world
    mp.spawn(_stripe_worker, args=(world, port, W, H, L, str(tmp_path)), nprocs=world, join=True)
    boxes, lights = oracle.scene_synthetic(W, H, L, n=600, n_lights=3)
    want = oracle.render(W, H, L, boxes, lights, want_gbuf=False, want_texel=False)["rgba"].view(np.uint8).reshape(-1)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"rank{r}.npy"), want), f"rank {r}"


@pytest.mark.parametrize("world,H", [(2, 320), (3, 320)])
def test_band_gather(oracle, tmp_path, world, H):
    W, L = 480, 320
    port = 29500 + os.getpid() % 2000 + world
    mp.spawn(_worker, args=(world, port, W, H, L, str(tmp_path)), nprocs=world, join=True)
    boxes, lights = oracle.scene_synthetic(W, H, L, n=600, n_lights=3)
    full = oracle.render(W, H, L, boxes, lights, want_gbuf=False, want_texel=False)["rgba"]
    want = full.view(np.uint8).reshape(-1)
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npy")
        assert np.array_equal(got, want), f"rank {r}"


def test_band_rows():
    from par_b200.bands import band_rows
    assert [band_rows(2160, 8, r) for r in (0, 7)] == [(0, 270), (1890, 2160)]
    assert [band_rows(320, 3, r) for r in range(3)] == [(0, 106), (106, 212), (212, 320)]
    assert band_rows(4320, 1, 0) == (0, 4320)
    with pytest.raises(ValueError):
        band_rows(320, 2, 2)


@pytest.mark.parametrize("W,H", [(3840, 2160), (7680, 4320), (1920, 1080), (640, 360), (480, 320)])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_split_stripes_partition(W, H, world):
    """par_config.stripe_split as bench.py picks it: the ranks' stripes tile the frame exactly once and,
    whenever a split that evens the counts exists, every rank owns the same number of pixels."""
    sys.path.insert(0, PKG)
    from par_b200.bands import owned_rects, owned_rows, stripe_split_for
    s = stripe_split_for(W, H, world)
    assert s in (1, 2, 4, 8) and (W // 40) % s == 0
    cover = np.zeros((H, W), np.int32)
    px = []
    for r in range(world):
        rects = owned_rects(W, H, world, r, s)
        for r0, r1, c0, c1 in rects:
            cover[r0:r1, c0:c1] += 1
        px.append(sum((r1 - r0) * (c1 - c0) for r0, r1, c0, c1 in rects))
        if s == 1:
            assert [(a, b) for a, b, _, _ in rects] == owned_rows(H, world, r)
    assert (cover == 1).all()
    if (H // 40 * s) % world == 0:
        assert len(set(px)) == 1
    if (W, H, world) == (7680, 4320, 8):
        assert s == 2 and px[0] == W * H // 8
    if s > 1:  # the column segments rotate: no rank renders only one side of the image
        cols = (W // 40 // s) * 40
        for r in range(world):
            per_seg = [sum(1 for _, _, c0, _ in owned_rects(W, H, world, r, s) if c0 == k * cols) for k in range(s)]
            assert max(per_seg) - min(per_seg) <= 1, (r, per_seg)


@pytest.fixture(scope="module")
def stripe_harness(tmp_path_factory):
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path_factory.mktemp("stripes") / "stripe_partition"
    subprocess.run([nvcc, "-std=c++17", "-x", "cu", "-I", os.path.join(PKG, "csrc"), "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "stripe_partition.cpp"), "-o", str(exe)], check=True)
    return str(exe)


@pytest.mark.parametrize("W,H,world,split", [(7680, 4320, 8, 2), (3840, 2160, 8, 4), (3840, 2160, 4, 2), (640, 360, 3, 2),
                                             (1280, 680, 2, 2), (1280, 720, 6, 4), (3840, 2160, 8, 8), (7680, 4320, 5, 1)])
def test_device_stripe_arithmetic_equals_the_host_partition(stripe_harness, W, H, world, split):
    """The stripe -> (tile row, column segment) arithmetic of the render kernel / C ABI (par_device.cuh, compiled
    for the host) deals exactly the rectangles par_b200.bands.owned_rects states, rank by rank, and
    stripe_of_segment (the cursor probe's inverse) undoes it."""
    import subprocess
    sys.path.insert(0, PKG)
    from par_b200.bands import owned_rects
    res = subprocess.run([stripe_harness, str(W), str(H), str(world), str(split)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-300:] + res.stderr
    got = {}
    for ln in res.stdout.splitlines():
        r, *rect = map(int, ln.split())
        got.setdefault(r, []).append(tuple(rect))
    for r in range(world):
        assert got.get(r, []) == owned_rects(W, H, world, r, split), f"rank {r}"
