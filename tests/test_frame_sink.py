"""CPU test of the frame sink of the headless host (include/par/frame_sink.hpp; SURVEY.md §8(f) row 3: the step
after the path, alternative.cpp:774-788): PPM, PNG and animated-GIF writers against an independent decoder
(Pillow) on synthetic frames — packed and pitched rows, few and many colours, sizes that put the end of the LZW
stream on every code-width boundary."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

Image = pytest.importorskip("PIL.Image")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = tmp_path_factory.mktemp("sink") / "frame_sink_main"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "frame_sink_main.cpp"), "-o", str(exe)], check=True)
    return str(exe)


def _frames(W, H, n, colours):
    y, x = np.mgrid[0:H, 0:W].astype(np.int64)
    out = []
    for f in range(n):
        i = (7 * x + 13 * y + 31 * f + (x * y) % 11) % colours
        out.append(np.stack([37 * i % 256, 91 * i % 256, 53 * i % 256], -1).astype(np.uint8))
    return out


@pytest.mark.parametrize("W,H,n,pad,colours", [(48, 32, 3, 0, 7), (50, 31, 2, 24, 200), (200, 120, 4, 0, 256),
                                               (64, 40, 2, 8, 1000), (640, 360, 2, 0, 97), (1, 1, 1, 0, 2), (3, 1, 2, 4, 2)])
def test_writers_against_pillow(harness, tmp_path, W, H, n, pad, colours):
    res = subprocess.run([harness, str(tmp_path), str(W), str(H), str(n), str(W * 4 + pad), str(colours)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    exact = [ln.split()[1] == "1" for ln in res.stdout.splitlines()]
    want = _frames(W, H, n, colours)
    gif = Image.open(os.path.join(tmp_path, "seq.gif"))
    assert gif.n_frames == n and gif.size == (W, H)
    assert gif.info.get("loop") == 0 and gif.info.get("duration") == 40
    strict = _strict_gif_frames(os.path.join(tmp_path, "seq.gif"))
    assert len(strict) == n
    for f in range(n):
        ppm = np.asarray(Image.open(os.path.join(tmp_path, f"frame_{f:03d}.ppm")).convert("RGB"))
        assert np.array_equal(ppm, want[f])
        png = Image.open(os.path.join(tmp_path, f"frame_{f:03d}.png"))
        png.verify()  # chunk CRCs
        png = np.asarray(Image.open(os.path.join(tmp_path, f"frame_{f:03d}.png")).convert("RGB"))
        assert np.array_equal(png, want[f])
        gif.seek(f)
        got = np.asarray(gif.convert("RGB"))
        n_colours = len(np.unique(want[f].reshape(-1, 3), axis=0))
        assert exact[f] == (n_colours <= 256)
        if exact[f]:
            assert np.array_equal(got, want[f])
            assert np.array_equal(strict[f], want[f])
        else:  # 6x7x6 uniform palette: every channel within half a level
            assert np.abs(got.astype(int) - want[f].astype(int)).max() <= 26


def _strict_gif_frames(path):
    """Minimal GIF89a reader with the textbook LZW decoder (codes widen when the NEXT free table slot reaches
    2^width, one entry behind the encoder); unlike Pillow it insists on reading the end code where it must be."""
    b = open(path, "rb").read()
    assert b[:6] == b"GIF89a"
    W, H = int.from_bytes(b[6:8], "little"), int.from_bytes(b[8:10], "little")
    pos = 13 + (3 << ((b[10] & 7) + 1) if b[10] & 0x80 else 0)
    frames = []
    while b[pos] != 0x3B:
        if b[pos] == 0x21:  # extension: skip its sub-blocks
            pos += 2
            while b[pos]:
                pos += 1 + b[pos]
            pos += 1
            continue
        assert b[pos] == 0x2C
        w, h, flags = int.from_bytes(b[pos + 5:pos + 7], "little"), int.from_bytes(b[pos + 7:pos + 9], "little"), b[pos + 9]
        assert (w, h) == (W, H) and flags & 0x80 and not flags & 0x40
        pos += 10
        n_tab = 2 << (flags & 7)
        table_rgb = np.frombuffer(b[pos:pos + 3 * n_tab], np.uint8).reshape(-1, 3)
        pos += 3 * n_tab
        min_bits = b[pos]
        pos += 1
        data = bytearray()
        while b[pos]:
            data += b[pos + 1:pos + 1 + b[pos]]
            pos += 1 + b[pos]
        pos += 1
        bits = int.from_bytes(bytes(data), "little")
        n_bits, at = 8 * len(data), 0
        clear, eoi = 1 << min_bits, (1 << min_bits) + 1
        width, nxt, prev = min_bits + 1, eoi + 1, None
        table = {i: bytes([i]) for i in range(clear)}
        out = bytearray()
        while True:
            assert at + width <= n_bits, "ran out of data before the end code"
            code = (bits >> at) & ((1 << width) - 1)
            at += width
            if code == clear:
                width, nxt, prev = min_bits + 1, eoi + 1, None
                table = {i: bytes([i]) for i in range(clear)}
                continue
            if code == eoi:
                break
            if prev is None:
                entry = table[code]
            else:
                assert code in table or code == nxt, f"code {code} beyond the table ({nxt})"
                entry = table[code] if code in table else table[prev] + table[prev][:1]
                if nxt < 4096:
                    table[nxt] = table[prev] + entry[:1]
                    nxt += 1
                    if nxt == (1 << width) and width < 12:
                        width += 1
            out += entry
            prev = code
        assert len(out) == W * H, "the end code is not where the pixels end"
        assert n_bits - at < 8, "data after the end code"
        frames.append(table_rgb[np.frombuffer(bytes(out), np.uint8)].reshape(H, W, 3))
    return frames


@pytest.mark.parametrize("colours", [2, 5])
def test_gif_stream_ends_on_every_code_width_boundary(harness, tmp_path, colours):
    """One-row images of every width up to 159: the last data code of the LZW stream lands on table sizes
    around 8, 16, 32, 64 — where the decoder widens its codes just before the end code (an encoder that writes
    the end code one bit short fails the strict reader at 2 colours x 31..33 pixels, for one)."""
    for W in range(1, 160):
        res = subprocess.run([harness, str(tmp_path), str(W), "1", "1", str(W * 4), str(colours)], capture_output=True, text=True)
        assert res.returncode == 0
        want = _frames(W, 1, 1, colours)[0]
        got = np.asarray(Image.open(os.path.join(tmp_path, "seq.gif")).convert("RGB"))
        assert np.array_equal(got, want), W
        assert np.array_equal(_strict_gif_frames(os.path.join(tmp_path, "seq.gif"))[0], want), W
