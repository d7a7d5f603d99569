"""Developer probe: where the host->host (e2e) frame time goes at C2.  D2H of the 33 MB frame (linear and pitched),
and the pipelined par_submit_frame / par_wait_frame loop with the per-frame stats the library reports."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pixel-art-raytracer_b200"))
import numpy as np
import torch
import par_b200 as par

W, H, L = 3840, 2160, 2160
boxes, lights = par.scene_default(), par.light_default()
h_boxes = par.pinned_empty(len(boxes), par.AABB); h_boxes[:] = boxes
out = [par.pinned_empty((H, W), par.COLOR) for _ in range(2)]
ren = par.Renderer(W, H, L)
s1 = torch.cuda.Stream()
ren.set_stream(s1.cuda_stream)
ren.set_atlas(); ren.set_scene(h_boxes); ren.render_device(lights); ren.sync()

def d2h(pitch, reps=10):
    buf = par.pinned_empty(H * pitch, np.uint8)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ren.read_frame_pitched(buf, pitch); ren.sync()
    a.record(s1)
    for _ in range(reps):
        ren.read_frame_pitched(buf, pitch)
    b.record(s1); ren.sync()
    return a.elapsed_time(b) / reps

t_lin = d2h(W * 4); t_2d = d2h(W * 4 + 256)
print(f"D2H 33.18 MB back to back: packed rows {t_lin:.4f} ms ({33.1776 / t_lin:.1f} GB/s), pitched rows {t_2d:.4f} ms ({33.1776 / t_2d:.1f} GB/s)")

for flush_on in (False, True):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    def run(n):
        st = []
        for i in range(n + 1):
            if i < n:
                if flush_on:
                    with torch.cuda.stream(s1):
                        flush.fill_(1)
                ren.submit_frame(h_boxes, lights, out[i & 1])
            if i:
                st.append(ren.wait_frame())
        return st
    run(10)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); st = run(100); dt = (time.perf_counter() - t0) / 100 * 1e3
    rb = np.mean([s["ms_readback"] for s in st[10:]]); tot = np.mean([s["ms_total"] for s in st[10:]])
    print(f"pipelined, L2 flush {'on' if flush_on else 'off'}: {dt:.4f} ms/frame wall; per frame submit->complete {tot:.4f} ms, kernels done->complete {rb:.4f} ms")
# CPU cost of a submit alone (no GPU wait): time the call
ts = []
for i in range(40):
    t0 = time.perf_counter(); ren.submit_frame(h_boxes, lights, out[i & 1]); ts.append(time.perf_counter() - t0); ren.wait_frame()
print(f"par_submit_frame call: {np.median(ts) * 1e6:.1f} us median on the host (graph re-capture + update + launch + copy enqueue)")
ts = []
player = boxes[0:1].copy()
for i in range(40):
    player["px"] += 1
    t0 = time.perf_counter(); ren.submit_update(0, player, lights, out[i & 1]); ts.append(time.perf_counter() - t0); ren.wait_frame()
print(f"par_submit_update call: {np.median(ts) * 1e6:.1f} us median on the host")
