import os, sys, time
sys.path.insert(0, "/root/repo/pixel-art-raytracer_b200")
import numpy as np, torch
import par_b200 as par
W,H,L=3840,2160,2160
boxes, lights = par.scene_default(), par.light_default()
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
for stripes in (1, 8):
    A = par.Renderer(W,H,L, stripe_count=stripes, stripe_index=0); B = par.Renderer(W,H,L, stripe_count=stripes, stripe_index=0)
    A.set_stream(sA.cuda_stream); B.set_stream(sB.cuda_stream)
    for r in (A,B): r.set_atlas(); r.set_scene(boxes); r.render_device(lights); r.sync()
    def timed(fn, n=200):
        fn(); torch.cuda.synchronize()
        a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(sA)
        for _ in range(n): fn()
        b.record(sA); torch.cuda.synchronize()
        return a.elapsed_time(b)/n
    ev = torch.cuda.Event(); ev2 = torch.cuda.Event()
    def serial():
        A.rebuild_grid(); A.render_device(lights)
    def overlapped():
        ev.record(sA); sB.wait_event(ev)       # fork
        B.rebuild_grid()                        # loader of the 'next frame' on the side stream
        A.render_device(lights)                 # render of this frame
        ev2.record(sB); sA.wait_event(ev2)      # join
    def render_only():
        A.render_device(lights)
    print(f"stripes {stripes}: serial loader+render {timed(serial):.4f} ms, render only {timed(render_only):.4f} ms, loader beside render {timed(overlapped):.4f} ms")
    A.close(); B.close()
