"""Developer probe: PCIe copy times that bound the e2e (host scene -> host frame) path."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pixel-art-raytracer_b200"))
import numpy as np
import torch
import par_b200 as par

dev = torch.device("cuda", 0)
W, H, L = 3840, 2160, 2160
d_frame = torch.zeros(W * H * 4, dtype=torch.uint8, device=dev)
h_frame = torch.zeros(W * H * 4, dtype=torch.uint8).pin_memory()
h_scene = torch.zeros(2596928, dtype=torch.uint8).pin_memory()
d_scene = torch.zeros(2596928, dtype=torch.uint8, device=dev)


def ev_time(fn, n=20):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


t = ev_time(lambda: h_frame.copy_(d_frame, non_blocking=True))
print(f"D2H 33.2 MB pinned: {t:.4f} ms = {33.1776 / t:.1f} GB/s")
t = ev_time(lambda: d_scene.copy_(h_scene, non_blocking=True))
print(f"H2D 2.6 MB pinned: {t:.4f} ms = {2.596928 / t:.1f} GB/s")

boxes, lights = par.scene_default(), par.light_default()
ren = par.Renderer(W, H, L)
ren.set_atlas()
hb = par.pinned_empty(len(boxes), par.AABB)
hb[:] = boxes
outs = [par.pinned_empty((H, W), par.COLOR) for _ in range(2)]
for steps in (20, 100):
    for i in range(4):
        ren.submit_frame(hb, lights, outs[i & 1])
        if i:
            ren.wait_frame()
    ren.wait_frame()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps + 1):
        if i < steps:
            ren.submit_frame(hb, lights, outs[i & 1])
        if i:
            st = ren.wait_frame()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3 / steps
    print(f"pipelined, no flush, {steps} steps: {dt:.4f} ms/frame (last frame submit->done {st['ms_total']:.3f} ms, of which kernels done->frame on host {st['ms_readback']:.3f})")
ren.set_scene(hb)
ren.render(lights, out=outs[0])
t0 = time.perf_counter()
for i in range(20):
    ren.set_scene(hb)
    ren.render(lights, out=outs[0])
print(f"sync set_scene+render: {(time.perf_counter() - t0) * 1e3 / 20:.4f} ms/frame")
