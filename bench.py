#!/usr/bin/env python
"""bench.py — Mrays/s of the render path (BASELINE.json metric) on 1..8 B200.

    python bench.py --gpus 1 --steps K --warmup W              # our arm, one GPU
    torchrun --nproc-per-node N ... bench.py --gpus N ...      # row bands + NCCL frame gather
    python bench.py --impl reference ...                       # the reference's own CPU loop

A step = one pass of the hot path (alternative.cpp:689-760: grid build, primary rays,
shading + shadow rays, RGBA8 frame) over one frame of the workload.  Default workload `c2`
= BASELINE.json configs[1]: the reference's default scene at 3840x2160, one light.
Rays are reference-equivalent rays: W*H*(1 + n_lights) per frame.

`value`  : whole-job Mrays/s with the scene already resident in HBM (device scene loader +
           both kernels + (N>1) the NCCL all-gather of the row bands), CUDA-event timed, L2
           flushed between steps, max over ranks.
`e2e`    : the same through the public C ABI with HOST buffers: par_set_scene (H2D of the
           AABBs from pinned memory) + render + D2H of the finished frame, every step.
`roofline`: the shade kernel (dominant) against the FP32/INT ALU issue roofline
           (SMs x 128 lanes x max SM clock) in reference-equivalent algorithmic lane-ops
           (SURVEY.md §8d); HBM figures are given beside it because the path is not HBM bound.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "pixel-art-raytracer_b200"))

WORKLOADS = {
    # name: (W, H, L, description)
    "c1": (480, 320, 320, "default scene (162308 sprites, 1 light) at the reference's built-in 480x320"),
    "c4": (1920, 1080, 1080, "default scene at 1920x1080, frame 0 of the 240-frame sequence"),
    "c2": (3840, 2160, 2160, "default scene (162308 sprites, 1 light) at 3840x2160"),
    "c3": (3840, 2160, 2160, "synthetic 10k sprites + 16 lights at 3840x2160 (seed 0xB200)"),
    "c5": (7680, 4320, 4320, "synthetic 10k sprites + 16 lights at 7680x4320 (seed 0xB200)"),
    "c5b": (7680, 4320, 4320, "synthetic 40k sprites + 16 lights at 7680x4320 (seed 0xB200)"),
}


def load_json(path, default=None):
    try:
        with open(path) as f:
            return json.load(f)
    except (OSError, ValueError):
        return default


# ------------------------------------------------------------------ clocks

class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip().split(", ")))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows]
        sm = sorted(int(r[0]) for r in rows if r[0].isdigit())
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        mx = [int(r[1]) for r in rows if r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(rows), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ reference arm / cpu baseline

def run_reference_binary(W, H, L, frames, budget_s):
    """Time the REAL reference (oracle/_ref tier-1 build, 1 thread: it has no threading) on
    the default scene at this view.  Returns (ms per frame list, note) or None."""
    exe = os.path.join(ROOT, "oracle", "_ref", f"ref_tier1_{W}x{H}x{L}")
    if not os.path.exists(exe):
        return None
    with tempfile.TemporaryDirectory() as td:
        env = dict(os.environ, PAR_REF_FRAMES=str(frames), PAR_REF_TIMES=f"{td}/t.txt")
        t0 = time.time()
        try:
            subprocess.run([exe], env=env, check=True, stdout=subprocess.DEVNULL, timeout=budget_s)
        except subprocess.TimeoutExpired:
            pass
        except (OSError, subprocess.CalledProcessError):
            return None
        try:
            ms = [int(ln.split()[1]) / 1e6 for ln in open(f"{td}/t.txt")]
        except OSError:
            ms = []
        return (ms, time.time() - t0) if ms else None


def run_oracle_port(name, W, H, L, budget_s):
    """Fallback CPU arm: the oracle port on all host cores, on a bounded row sample."""
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    boxes, lights = (O.scene_default(), O.light_default()) if name in ("c1", "c2", "c4") else \
        O.scene_synthetic(W, H, L, n=40000 if name == "c5b" else 10000)
    rows = max(40, min(H, (H // 40 // 8) * 40))
    row0 = (H // 2 // 40) * 40
    t0 = time.time()
    O.render(W, H, L, boxes, lights, row0=row0, row1=row0 + rows, want_gbuf=False, want_texel=False)
    dt = time.time() - t0
    return dt, rows, len(lights), os.cpu_count()


def reference_arm(args):
    W, H, L, desc = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    steps, warmup = args.steps, args.warmup
    line = {"impl": "reference", "metric": "Mrays/s", "unit": "Mrays/s", "n_gpus": args.gpus,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": f"{args.workload}: {desc}"}}
    res = None
    if args.workload in ("c1", "c2", "c4"):
        # ~17 s per 4K frame on one core: clamp the frame count to the time budget
        per = {"c1": 0.06, "c4": 3.0, "c2": 20.0}[args.workload]
        frames = max(2, min(steps + warmup, int(args.cpu_budget / per)))
        warm = 1 if (warmup and frames > 1) else 0  # one untimed frame is enough for a CPU loop
        res = run_reference_binary(W, H, L, frames, args.cpu_budget * 2 + 60)
    if res:
        ms, wall = res
        ms_t = ms[warm:] if len(ms) > warm else ms
        mean = sum(ms_t) / len(ms_t)
        value = W * H * 2 / mean / 1e3
        line.update({"value": value, "steps": len(ms_t), "warmup": warm, "ms_per_step": mean,
                     "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": 1, "kind": "reference",
                                      "sample": f"{len(ms_t)} whole frames of the unmodified reference loop "
                                                f"(alternative.cpp:689-772, oracle/_ref tier-1 build, -O3, single "
                                                f"thread: the reference has no threading); {steps} steps requested"}})
    else:
        dt, rows, nl, cores = run_oracle_port(args.workload, W, H, L, args.cpu_budget)
        value = rows * W * (1 + nl) / dt / 1e6
        line.update({"value": value, "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3,
                     "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port",
                                      "sample": f"oracle port (OpenMP, {cores} threads) on rows of a {rows}-row band "
                                                f"of the frame incl. the whole-frame grid build"}})
    line["e2e"] = {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ our arm

def make_workload(par, name):
    W, H, L, desc = WORKLOADS[name]
    if name in ("c1", "c2", "c4"):
        boxes, lights = par.scene_default(), par.light_default()
    else:
        boxes, lights = par.scene_synthetic(W, H, L, n=40000 if name == "c5b" else 10000, n_lights=16)
    return W, H, L, desc, boxes, lights


def ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import par_b200 as par

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the render path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    trace_on = bool(os.environ.get("PAR_BENCH_TRACE"))

    def trace(msg):
        if trace_on:
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    W, H, L, desc, boxes, lights = make_workload(par, args.workload)
    n_lights = len(lights)
    from par_b200.bands import gather_stripes, owned_rows
    # interleaved 40-row stripes: tile row t belongs to rank t % N (balances the walk cost)
    my_rows = sum(b - a for a, b in owned_rows(H, world, rank))
    rays_frame = W * H * (1 + n_lights)

    stream = torch.cuda.Stream(device=dev)
    ren = par.Renderer(W, H, L, device=local, stripe_count=world, stripe_index=rank)
    ren.set_stream(stream.cuda_stream)
    ren.set_atlas()
    frame = torch.zeros(H * W * 4, dtype=torch.uint8, device=dev)  # the full raster frame in HBM
    staging = torch.zeros(ren.staging_bytes(), dtype=torch.uint8, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    h_boxes = par.pinned_empty(len(boxes), par.AABB)
    h_boxes[:] = boxes
    h_frame = par.pinned_empty((H, W), par.COLOR) if rank == 0 else None
    t_hframe = torch.from_numpy(h_frame.view(np.uint8).reshape(-1)) if rank == 0 else None

    # Frame exchange at N > 1.  "root" (default) / "peer": the shade kernel stores its finished
    # stripes straight into rank 0's / every rank's raster frame through CUDA-IPC-mapped peer memory
    # (NVLink); a tiny all-reduce is the barrier.  "nccl": stripe-major staging + in-place
    # all-gather + un-stripe.
    exchange = "none" if world == 1 else args.exchange
    token = torch.zeros(1, dtype=torch.int32, device=dev)
    if exchange in ("peer", "root"):
        try:
            handles = [None] * world
            dist.all_gather_object(handles, ren.peer_export())
            for r in ([0] if exchange == "root" else range(world)):
                if r != rank:
                    ren.peer_import(r, handles[r])
            ok = torch.ones(1, dtype=torch.int32, device=dev)
        except Exception as e:  # no IPC / no peer access on this box: use the NCCL gather
            if rank == 0:
                print(f"bench.py: peer exchange unavailable ({e}); using nccl", file=sys.stderr)
            ok = torch.zeros(1, dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            exchange = "nccl"
    trace(f"exchange = {exchange}")

    # Host-side frames at N > 1: two frames (alternating) in shared memory, page-locked by every rank;
    # each rank DMAs the stripes it rendered straight into them over its own PCIe link.
    shared = shared2 = None
    if world > 1:
        name = [f"/dev/shm/par_bench_{os.getpid()}" if rank == 0 else None]
        dist.broadcast_object_list(name, src=0)
        try:
            if rank == 0:
                shared2 = par.shared_host_frame(name[0], H, W, create=True, frames=2)
            dist.barrier()
            if rank != 0:
                shared2 = par.shared_host_frame(name[0], H, W, create=False, frames=2)
            shared = shared2[0]
            ok = torch.ones(1, dtype=torch.int32, device=dev)
        except Exception as e:
            trace(f"shared host frame unavailable: {e}")
            ok = torch.zeros(1, dtype=torch.int32, device=dev)
            if rank == 0 and shared2 is None:
                dist.barrier()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if rank == 0 and os.path.exists(name[0]):
            os.unlink(name[0])                   # the mappings keep it alive
        if int(ok.item()) == 0:
            shared = shared2 = None
    trace(f"shared host frame: {shared is not None}")

    def render_and_gather():
        if world == 1:
            ren.render_device(lights, frame.data_ptr())
        elif exchange in ("peer", "root"):
            ren.render_device_peers(lights)      # my stripes -> my frame and, in place, the imported peer frames
            dist.all_reduce(token)               # barrier: every rank's kernel (and its remote stores) is done
        else:
            ren.render_device_striped(lights, staging.data_ptr())  # my stripes, contiguous in staging
            gather_stripes(staging, world, rank)                    # in-place NCCL all-gather over NVLink
            ren.unstripe_device(staging.data_ptr(), frame.data_ptr())  # staging -> raster frame

    def step_resident():
        ren.rebuild_grid()                       # device scene loader on the resident scene
        render_and_gather()

    def step_e2e():
        ren.set_scene(h_boxes)                   # H2D from pinned memory + scene loader
        if world == 1:
            ren.render(lights, out=h_frame)      # the drop-in call: render + D2H into a host frame
            return
        if shared is not None:
            ren.render_device(lights)            # my stripes into my own frame: no GPU-to-GPU exchange needed
            ren.read_stripes(shared)             # ... and from there into the shared host frame (my PCIe link)
            dist.all_reduce(token)               # the host frame is complete when every rank's DMA is
            return
        render_and_gather()
        if exchange in ("peer", "root"):
            if rank == 0:
                ren.read_frame(h_frame)          # D2H of the finished frame (it lives in the context's frame)
            dist.all_reduce(token)               # nobody starts overwriting frames before the reader is done
        elif rank == 0:
            t_hframe.copy_(frame, non_blocking=True)  # D2H of the gathered frame

    sig_stream = torch.cuda.Stream(device=dev)

    def timed_pipelined(steps, warmup):
        """e2e through par_submit_frame / par_wait_frame: two frames in flight, every step uploads the
        scene from pinned memory and reads its frame back into pinned memory (at N > 1: every rank
        its own stripes, into the shared host frames; after each completed frame a 4-byte all-reduce
        on a side stream tells every rank that the whole frame is on the host).  The timed region
        starts with an empty pipeline and ends when the last frame is complete on the host."""
        h_out = [h_frame, par.pinned_empty((H, W), par.COLOR)] if world == 1 else [shared2[0], shared2[1]]

        def run(n):
            for i in range(n + 1):
                if i < n:
                    flush.fill_(1)               # L2 flush between iterations (inside the timed region here)
                    ren.submit_frame(h_boxes, lights, h_out[i & 1])
                if i:
                    ren.wait_frame()
                    if world > 1:
                        with torch.cuda.stream(sig_stream):
                            dist.all_reduce(token)
            sig_stream.synchronize()

        with torch.cuda.stream(stream):
            run(warmup)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            t0 = time.perf_counter()
            run(steps)
            b.record(stream)
            barrier()
            t1 = time.perf_counter()
        if (steps - 1) & 1:                      # the frame check looks at h_out[0]
            h_out[0][:] = h_out[1]
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, t0, t1

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(step_fn, steps, warmup):
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                step_fn()
            barrier()
            evs = []
            t0 = time.perf_counter()
            for _ in range(steps):
                flush.fill_(1)                   # L2 flush between timed iterations (untimed)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                step_fn()
                b.record(stream)
                evs.append((a, b))
            barrier()
            t1 = time.perf_counter()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, t0, t1

    sampler = ClockSampler(local) if rank == 0 else None
    with torch.cuda.stream(stream):
        ren.set_scene(h_boxes)
        step_resident()
        trace("first step enqueued")
        torch.cuda.synchronize(dev)
        # untimed pre-warm: lets the SM clock ramp from idle and gives nvidia-smi time to sample.
        # The number of steps is the same on every rank (the steps contain collectives).
        t_a = time.perf_counter()
        for _ in range(5):
            step_resident()
        torch.cuda.synchronize(dev)
        n_pre = torch.tensor([int(args.prewarm_ms / 1e3 / max(time.perf_counter() - t_a, 1e-6) * 5) + 1],
                             dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(n_pre, op=dist.ReduceOp.MAX)
        for i in range(min(int(n_pre.item()), 100000)):
            step_resident()
            if i % 8 == 7:
                torch.cuda.synchronize(dev)
    torch.cuda.synchronize(dev)
    trace("pre-warm done")

    t_load0 = time.perf_counter()
    ms, t0, t1 = timed(step_resident, args.steps, args.warmup)
    st = ren.stats()                            # per-kernel CUDA-event times of the last step
    trace("resident timing done")
    ms_e2e, _, t1 = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    ms_e2e_sync = ms_e2e
    pipelined = world == 1 or shared2 is not None
    if pipelined:
        ms_e2e, _, t1 = timed_pipelined(args.steps, max(2, args.warmup // 2))
    trace("e2e timing done")
    clocks = sampler.stop(t_load0 - args.prewarm_ms / 1e3, t1) if sampler else None  # pre-warm + timed regions: under load

    # per-kernel times over a few more steps, for the roofline of the dominant kernel
    shade_ms, prim_ms, build_ms = [], [], []
    with torch.cuda.stream(stream):
        for _ in range(min(args.steps, 10)):
            flush.fill_(1)
            step_resident()
            s = ren.stats()
            shade_ms.append(s["ms_shade"])
            prim_ms.append(s["ms_primary"])
            build_ms.append(s["ms_grid_build"])
    # correctness spot check of what was timed (untimed): the frame the last e2e step left in host
    # memory, and the device frame of the last resident step, equal a fresh 1-context render
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    frame_check = None
    if rank == 0:
        with par.Renderer(W, H, L, device=local) as chk:
            chk.set_atlas()
            chk.set_scene(boxes)
            want, _ = chk.render(lights)
        got_host = shared if shared is not None else h_frame
        if world == 1:
            got_dev = frame.cpu().numpy().view(par.COLOR).reshape(H, W)
        elif exchange in ("peer", "root"):
            got_dev = ren.read_frame()
            ren.sync()
        else:
            got_dev = frame.cpu().numpy().view(par.COLOR).reshape(H, W)
        same_host = bool(np.array_equal(np.asarray(got_host).view(np.uint32), want.view(np.uint32)))
        same_dev = bool(np.array_equal(got_dev.view(np.uint32), want.view(np.uint32)))
        frame_check = {"host_frame_equals_1ctx_render": same_host, "device_frame_equals_1ctx_render": same_dev}
        if not (same_host and same_dev):
            raise SystemExit(f"bench.py: timed frames differ from a 1-context render: {frame_check}")
    if world > 1:
        # orderly teardown: nobody frees a frame another rank still has mapped
        torch.cuda.synchronize(dev)
        dist.barrier()
        ren.close()
        dist.barrier()
        trace("contexts closed")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # correctness spot check of what was timed: the gathered frame equals a 1-context render
    value = rays_frame * args.steps / ms / 1e3
    e2e = rays_frame * args.steps / ms_e2e / 1e3
    peaks = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"), {}) or {}
    ops_tab = load_json(os.path.join(ROOT, "tests", "golden", "workload_ops.json"), {}) or {}
    prop = torch.cuda.get_device_properties(dev)
    sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    peak_ops = prop.multi_processor_count * 128 * sm_max * 1e6 / 1e12  # T lane-ops/s
    ops = (ops_tab.get(args.workload) or {}).get("algorithmic_ops")
    shade = sum(shade_ms) / len(shade_ms)
    frac_rows = my_rows / H
    # shade kernel's share of the algorithmic ops: everything but the primary-ray counters
    roofline = {"bound": "fp32_alu", "kernel": "k_shade", "unit": "Tlane-op/s", "peak": round(peak_ops, 2),
                "peak_source": f"{prop.multi_processor_count} SMs x 128 lanes x {sm_max:.0f} MHz (1 op/lane/clk; no FMA "
                               "on this path) — nominal issue ceiling, of nominal",
                "traffic": None}
    if ops:
        c = ops_tab[args.workload]["counters"]
        shade_ops = (59.0 * c["shaded_px_lights"] + 17.0 * c["lit_px_lights"] + 19.0 * c["shadow_probes"] +
                     5.0 * c["shadow_slot_entries"] + 32.0 * c["slab_tests"]) * frac_rows
        ach = shade_ops / (shade * 1e-3) / 1e12
        roofline.update({"achieved": round(ach, 2), "frac": round(ach / peak_ops, 3),
                         "algorithmic_ops_per_launch": shade_ops, "ms_per_launch": round(shade, 4),
                         "note": "reference-equivalent algorithmic lane-ops (SURVEY.md §8d weights x oracle counters) "
                                 "/ CUDA-event time of k_shade; >1 means the kernel legally skips work the reference "
                                 "does (shared grid walks, de-duplicated probes, Q19); see profiles/ for ncu pipe "
                                 "utilisation"})
    traffic = (load_json(os.path.join(ROOT, "profiles", "roofline_traffic.json"), {}) or {}).get(args.workload, {}).get("k_shade")
    if traffic and world == 1:
        roofline["traffic"] = traffic["dram_bytes_read"] + traffic["dram_bytes_write"]
        roofline["traffic_source"] = traffic["source"] + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)"
        roofline["ncu_issue_active_pct"] = traffic.get("smsp_issue_active_pct")
        if traffic.get("warp_instructions"):
            # what the SMs actually issued: ncu's warp-instruction count of one launch x 32 lanes / live time
            ex = traffic["warp_instructions"] * 32 / (shade * 1e-3) / 1e12
            roofline["executed"] = {"warp_instructions_per_launch": traffic["warp_instructions"],
                                    "achieved": round(ex, 2), "frac": round(ex / peak_ops, 3), "unit": "Tlane-op/s",
                                    "note": "issue-slot utilisation of k_shade: executed warp instructions (ncu "
                                            "smsp__inst_executed.sum) x 32 / CUDA-event time / the same peak"}
    hbm_bytes = 16.0 * W * my_rows + 4.0 * W * my_rows
    roofline["hbm"] = {"algorithmic_bytes_per_launch": hbm_bytes, "achieved_gbs": round(hbm_bytes / (shade * 1e-3) / 1e9, 1),
                       "peak_gbs": peaks.get("hbm_gbs"), "note": "G-buffer read + RGBA8 write; not the bound"}

    line = {
        "metric": "Mrays/s", "value": round(value, 1), "unit": "Mrays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "view": [W, H, L], "n_entities": int(len(boxes)),
                   "n_lights": int(n_lights), "rays_per_frame": rays_frame,
                   "parallelism": ("1 GPU" if world == 1 else
                                   f"interleaved 40-row stripes x{world}, frame exchange fused into the shade kernel "
                                   "(peer-memory stores over NVLink + all-reduce barrier), frame complete on "
                                   + ("rank 0 (gather-to-root)" if exchange == "root" else "every GPU (all-gather)")
                                   if exchange in ("peer", "root") else
                                   f"interleaved 40-row stripes x{world} + in-place NCCL all-gather of the RGBA8 frame"),
                   "l2": "flushed between timed steps (256 MB fill)", "frames_per_s": round(1e3 * args.steps / ms, 2)},
        "e2e": {"value": round(e2e, 1), "unit": "Mrays/s", "ms_per_step": round(ms_e2e / args.steps, 4),
                "frames_per_s": round(1e3 * args.steps / ms_e2e, 2),
                "h2d_bytes_per_step": int(h_boxes.nbytes) * world, "d2h_bytes_per_step": int(H * W * 4),
                "readback": ("each rank DMAs its own stripes into one shared pinned host frame (N PCIe links)"
                             if shared is not None else "rank 0 / the one context copies the whole frame"),
                "api": ("par_submit_frame / par_wait_frame, two frames in flight (readback of frame k beside upload "
                        "and kernels of frame k+1, the frame's GPU work as one CUDA graph launch); timed from an "
                        "empty pipeline to the last frame complete on the host, L2 flush inside the timed region"
                        + ("" if world == 1 else "; every rank ships its own stripes, a 4-byte all-reduce per frame "
                                                 "signals completion") if pipelined else
                        "par_set_scene + par_render_device + par_read_stripes per rank, barrier per step"),
                "sync_call_ms": round(ms_e2e_sync / args.steps, 4),
                "sync_call": ("par_set_scene + par_render (blocking drop-in call) per step" if world == 1 else
                              "par_set_scene + par_render_device + par_read_stripes per rank, barrier per step")},
        "gpu_launches": ((6 if world == 1 or exchange in ("peer", "root") else 7) * args.steps) * world,
        "kernels_ms": {"scene_loader": round(sum(build_ms) / len(build_ms), 4),
                       "k_primary": round(sum(prim_ms) / len(prim_ms), 4), "k_shade": round(shade, 4)},
        "roofline": roofline, "clocks": clocks, "frame_check": frame_check,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, W, H, L)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args, W, H, L):
    """Reported baseline (not the target): the real reference loop, 1 thread, on this box."""
    if args.workload in ("c1", "c2", "c4"):
        frames = {"c1": 50, "c4": 4, "c2": 2}[args.workload]
        res = run_reference_binary(W, H, L, frames, args.cpu_budget * 2 + 60)
        if res:
            ms = res[0][1:] or res[0]
            mean = sum(ms) / len(ms)
            out = {"value": round(W * H * 2 / mean / 1e3, 3), "unit": "Mrays/s", "cores": 1, "kind": "reference",
                   "ms_per_frame": round(mean, 1), "host_cores_available": os.cpu_count(),
                   "sample": f"{len(ms)} whole frame(s) of the unmodified reference loop (oracle/_ref tier-1 build, "
                             "-O3, single thread: the reference has no threading), after 1 warm-up frame"}
            try:  # the same arithmetic on all host cores (the oracle port, OpenMP over rows)
                dt, rows, nl, cores = run_oracle_port(args.workload, W, H, L, args.cpu_budget)
                out["port_all_cores"] = {"value": round(rows * W * (1 + nl) / dt / 1e6, 3), "unit": "Mrays/s",
                                         "cores": cores, "kind": "port",
                                         "sample": f"oracle port (OpenMP, {cores} threads), one {rows}-row band of the "
                                                   "frame incl. whole-frame grid build"}
            except Exception as e:  # informational only
                out["port_all_cores"] = {"unavailable": str(e)[:120]}
            return out
    dt, rows, nl, cores = run_oracle_port(args.workload, W, H, L, args.cpu_budget)
    return {"value": round(rows * W * (1 + nl) / dt / 1e6, 3), "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": f"oracle port (OpenMP, {cores} threads), one {rows}-row band of the frame incl. whole-frame grid build"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-budget", type=float, default=60.0, help="seconds of CPU work for the CPU legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="root", choices=["root", "peer", "nccl"],
                    help="N>1 frame exchange, fused into the shade kernel as peer-memory stores: 'root' gathers the "
                         "frame on rank 0 (default; SURVEY.md 8e gather-to-root), 'peer' completes it on every GPU; "
                         "'nccl' = stripe-major staging + NCCL all-gather + un-stripe")
    ap.add_argument("--prewarm-ms", type=float, default=400.0, help="untimed GPU warm-up before the W warm-up steps")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
