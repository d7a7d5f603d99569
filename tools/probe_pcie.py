"""Developer probe: H2D time of the 2.6 MB scene while a 33 MB D2H runs on another stream; and
kernel time of a resident frame while a D2H runs."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pixel-art-raytracer_b200"))
import numpy as np
import torch
import par_b200 as par

dev = torch.device("cuda", 0)
n_s, n_f = 2596928, 33177600
hs = torch.from_numpy(par.pinned_empty(n_s, np.uint8)); hs.fill_(1)
hf = torch.from_numpy(par.pinned_empty(n_f, np.uint8))
ds = torch.empty(n_s, dtype=torch.uint8, device=dev)
df = torch.zeros(n_f, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def h2d_time(with_d2h):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if with_d2h:
        with torch.cuda.stream(s2):
            for _ in range(4):
                hf.copy_(df, non_blocking=True)
    with torch.cuda.stream(s1):
        torch.cuda._sleep(200000)  # ~0.1 ms: let the D2H get going
        a.record(s1)
        for _ in range(5):
            ds.copy_(hs, non_blocking=True)
        b.record(s1)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / 5


print(f"H2D 2.6 MB alone: {h2d_time(False):.4f} ms; beside a running 33 MB D2H: {h2d_time(True):.4f} ms")

W, H, L = 3840, 2160, 2160
ren = par.Renderer(W, H, L)
ren.set_stream(s1.cuda_stream)
ren.set_atlas()
ren.set_scene(par.scene_default())
lights = par.light_default()
for with_d2h in (False, True):
    ren.render_device(lights)
    torch.cuda.synchronize()
    if with_d2h:
        with torch.cuda.stream(s2):
            for _ in range(6):
                hf.copy_(df, non_blocking=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s1):
        torch.cuda._sleep(200000)
        a.record(s1)
        for _ in range(5):
            ren.rebuild_grid()
            ren.render_device(lights)
        b.record(s1)
    torch.cuda.synchronize()
    print(f"loader+primary+shade {'beside a running D2H' if with_d2h else 'alone'}: {a.elapsed_time(b) / 5:.4f} ms", ren.stats())
