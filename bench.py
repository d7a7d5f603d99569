#!/usr/bin/env python
"""bench.py — Mrays/s of the render path (BASELINE.json metric) on 1..8 B200.

    python bench.py --gpus 1 --steps K --warmup W              # our arm, one GPU
    torchrun --nproc-per-node N ... bench.py --gpus N ...      # interleaved stripes, frame exchange over NVLink
    python bench.py --impl reference ...                       # the reference's own CPU loop

A step = one pass of the hot path (alternative.cpp:689-760: grid build, primary rays,
shading + shadow rays, RGBA8 frame) over one frame of the workload.  Default workload `c2`
= BASELINE.json configs[1]: the reference's default scene at 3840x2160, one light.
Rays are reference-equivalent rays: W*H*(1 + n_lights) per frame.

`value`    : whole-job Mrays/s with the scene already resident in HBM: par_render_resident = the render
             kernel and, on a side branch of the same CUDA graph, the device scene loader rebuilding the
             other grid generation for the next frame — every step runs one loader and one render kernel
             (+ at N>1 the frame exchange: peer-memory stores fused into the kernel, arrival/credit flags in
             the frame footers, no collective); CUDA-event timed on the launching stream, L2 flushed between
             steps, max over ranks.
`e2e`      : the same through the public C ABI with HOST buffers: par_submit_frame / par_wait_frame
             (H2D of the AABBs from pinned memory, D2H of the finished frame), every step.
`roofline` : the render kernel against the FP32/INT ALU issue roofline (SMs x 128 lanes x max SM
             clock).  `frac` is the HARDWARE fraction: warp instructions the kernel executes (ncu
             smsp__inst_executed.sum of a capture of THIS build, profiles/roofline_r02.json, refused when
             the source hash differs) x 32 / live CUDA-event time / peak.  `algorithmic_speedup` is the
             reference-equivalent figure (SURVEY.md §8d weights x oracle counters / time / peak): how
             much of the reference's work the kernel legally never does.
`scale_8k` : the 7680x4320 workloads of BASELINE.json configs[4] (c5b: 40k sprites, c5: 10k sprites,
             16 lights) timed the same way as `value`, at every N, in the same line; `c3_series` likewise for
             configs[2] (10k sprites + 16 lights at 3840x2160).
`c4_sequence`: configs[3] — 240 frames of key script D at 1920x1080 through the pipelined calls, with
             the per-frame hash file checked against the real reference's (N=1 only).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "pixel-art-raytracer_b200")
sys.path.insert(0, PKG)

WORKLOADS = {
    # name: (W, H, L, description)
    "c1": (480, 320, 320, "default scene (162308 sprites, 1 light) at the reference's built-in 480x320"),
    "c4": (1920, 1080, 1080, "default scene at 1920x1080, frame 0 of the 240-frame sequence"),
    "c2": (3840, 2160, 2160, "default scene (162308 sprites, 1 light) at 3840x2160"),
    "c3": (3840, 2160, 2160, "synthetic 10k sprites + 16 lights at 3840x2160 (seed 0xB200)"),
    "c5": (7680, 4320, 4320, "synthetic 10k sprites + 16 lights at 7680x4320 (seed 0xB200)"),
    "c5b": (7680, 4320, 4320, "synthetic 40k sprites + 16 lights at 7680x4320 (seed 0xB200)"),
}
ROOFLINE_PROFILE = os.path.join(ROOT, "profiles", "roofline_r02.json")


def load_json(path, default=None):
    try:
        with open(path) as f:
            return json.load(f)
    except (OSError, ValueError):
        return default


def kernel_source_sha() -> str:
    """Hash of everything the device code is built from; keys the ncu capture the roofline quotes."""
    h = hashlib.sha256()
    files = [os.path.join(PKG, "build_native.sh"), os.path.join(ROOT, "include", "par", "par.h")]
    csrc = os.path.join(PKG, "csrc")
    files += sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh")))
    for p in files:
        h.update(os.path.basename(p).encode() + b"\0")
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


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
        # ~8-17 s per 4K frame on one core: clamp the frame count to the time budget
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


def script_d_keys(f):
    """Keys delivered before frame f of key script D: script C's player key + light key 'o'."""
    keys, k = [], f - 1
    if k >= 0:
        for n, key in ((30, "R"), (20, "U"), (50, "L"), (30, "D"), (30, "P"), (40, "R"), (30, "p"), (9, "U")):
            if k < n:
                keys.append(key)
                break
            k -= n
        keys.append("o")
    return keys


class Job:
    """One workload on this rank's GPU: a striped context (tile row t -> rank t % N), the frame exchange
    of par_render_resident at N > 1, and the timing loop of the device-resident step."""

    def __init__(self, env, name, exchange, want_sha=None):
        import numpy as np
        import torch
        import torch.distributed as dist
        par = env["par"]
        self.env, self.name, self.par = env, name, par
        world, rank, local, dev = env["world"], env["rank"], env["local"], env["dev"]
        self.W, self.H, self.L, self.desc, self.boxes, self.lights = make_workload(par, name)
        self.rays_frame = self.W * self.H * (1 + len(self.lights))
        # equal stripe counts per rank (e.g. 108 tile rows over 8 GPUs: half tile rows); the NCCL fallback's
        # stripe-major staging needs whole tile rows.  PAR_BENCH_STRIPE_SPLIT overrides the choice.
        from par_b200.bands import stripe_split_for
        forced = os.environ.get("PAR_BENCH_STRIPE_SPLIT")
        self.split = 1 if exchange == "nccl" or world == 1 else \
            (max(1, int(forced)) if forced else stripe_split_for(self.W, self.H, world))
        self.h_boxes = par.pinned_empty(len(self.boxes), par.AABB)
        self.h_boxes[:] = self.boxes
        self.np = np
        self.split_fallback = None
        self._build(exchange)
        if self.split > 1 and not self._split_frames_ok(want_sha):
            # never observed; the partition is new this round and a driver run must not die of it
            self.split_fallback = f"stripe_split {self.split} gave a wrong or no frame at N = {world}; whole tile rows used"
            if rank == 0:
                print(f"bench.py: {self.split_fallback}", file=sys.stderr)
            self.close()
            self.split = 1
            self._build(exchange)

    def _build(self, exchange):
        import torch
        import torch.distributed as dist
        env, par = self.env, self.par
        world, rank, local, dev = env["world"], env["rank"], env["local"], env["dev"]

        def make_renderer():
            ren = par.Renderer(self.W, self.H, self.L, device=local, stripe_count=world, stripe_index=rank,
                               stripe_split=self.split)
            ren.set_stream(env["stream"].cuda_stream)
            ren.set_atlas()
            return ren

        self.ren = make_renderer()
        self.exchange = "none" if world == 1 else exchange
        self.staging = self.frame = None
        if world > 1 and self.exchange in ("peer", "root"):
            # the render kernel stores its finished stripes straight into rank 0's / every rank's raster
            # frame through CUDA-IPC-mapped peer memory (NVLink); completion travels as flags in the frame
            # footers (par_exchange_setup) — no collective, no host round trip
            try:
                handles = [None] * world
                dist.all_gather_object(handles, self.ren.peer_export())
                for r in range(world):
                    if r != rank:
                        self.ren.peer_import(r, handles[r])
                self.ren.exchange_setup(0 if self.exchange == "root" else -1)
                ok = torch.ones(1, dtype=torch.int32, device=dev)
            except Exception as e:  # no IPC / no peer access on this box: use the NCCL gather
                if rank == 0:
                    print(f"bench.py: peer exchange unavailable ({e}); using nccl", file=sys.stderr)
                ok = torch.zeros(1, dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                self.exchange = "nccl"
                if self.split > 1:  # the staging frame of the fallback is made of whole tile rows
                    self.ren.close()
                    self.split = 1
                    self.ren = make_renderer()
        if world > 1 and self.exchange == "nccl":
            self.staging = torch.zeros(self.ren.staging_bytes(), dtype=torch.uint8, device=dev)
            self.frame = torch.zeros(self.H * self.W * 4, dtype=torch.uint8, device=dev)
        with torch.cuda.stream(env["stream"]):
            self.ren.set_scene(self.h_boxes)
        self.ren.sync()

    def _split_frames_ok(self, want_sha):
        """Two resident frames with the split partition, compared with the committed oracle frame on every rank
        that holds a whole frame; False on any rank -> False on all (no host collective inside the try)."""
        import torch
        import torch.distributed as dist
        env = self.env
        try:
            with torch.cuda.stream(env["stream"]):
                for _ in range(2):
                    self.step_resident()
            torch.cuda.synchronize(env["dev"])
            good = True
            if want_sha and (env["rank"] == 0 or self.exchange in ("peer", "nccl")):
                good = self.device_frame_sha() == want_sha
            # the host-side form of the same partition (e2e: every rank DMAs its own stripes into one host frame)
            from par_b200.bands import owned_rects
            par = self.par
            with par.Renderer(self.W, self.H, self.L, device=env["local"]) as one:
                one.set_atlas()
                one.set_scene(self.boxes)
                full, _ = one.render(self.lights)
            part = par.pinned_empty((self.H, self.W), par.COLOR)
            self.ren.read_stripes(part)
            self.ren.sync()
            for r0, r1, c0, c1 in owned_rects(self.W, self.H, env["world"], env["rank"], self.split):
                good = good and self.np.array_equal(part[r0:r1, c0:c1], full[r0:r1, c0:c1])
        except Exception as e:
            print(f"bench.py rank {env['rank']}: split check failed: {e}", file=sys.stderr)
            good = False
        ok = torch.tensor([1 if good else 0], dtype=torch.int32, device=env["dev"])
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        return int(ok.item()) == 1

    def step_resident(self):
        if self.exchange == "nccl":  # stripe-major staging + in-place NCCL all-gather + un-stripe
            from par_b200.bands import gather_stripes
            self.ren.rebuild_grid()
            self.ren.render_device_striped(self.lights, self.staging.data_ptr())
            gather_stripes(self.staging, self.env["world"], self.env["rank"])
            self.ren.unstripe_device(self.staging.data_ptr(), self.frame.data_ptr())
        else:  # loader + render kernel (+ fused exchange and its flags) as one graph launch
            self.ren.render_resident(self.lights)

    def launches_per_step(self):
        n = self.ren.stats()["kernel_launches"]
        return n + (1 if self.exchange == "nccl" else 0)  # + the un-stripe kernel (NCCL's own kernels not counted)

    def timed(self, step_fn, steps, warmup):
        import torch
        import torch.distributed as dist
        env = self.env
        stream, dev, flush = env["stream"], env["dev"], env["flush"]
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                step_fn()
            env["barrier"]()
            evs = []
            t0 = time.perf_counter()
            for _ in range(steps):
                flush.fill_(1)                   # L2 flush between timed iterations (untimed)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                step_fn()
                b.record(stream)
                evs.append((a, b))
            env["barrier"]()
            t1 = time.perf_counter()
        per = [a.elapsed_time(b) for a, b in evs]
        ms = sum(per)
        if env["world"] > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, t0, t1, per

    def kernel_times(self, reps):
        """Per-kernel CUDA-event times of un-graphed steps (a graph replay carries no timing events)."""
        import torch
        build, render = [], []
        with torch.cuda.stream(self.env["stream"]):
            for _ in range(reps):
                self.env["flush"].fill_(1)
                self.ren.rebuild_grid()
                if self.exchange in ("peer", "root"):
                    self.ren.render_device_peers(self.lights)
                elif self.exchange == "nccl":
                    self.ren.render_device_striped(self.lights, self.staging.data_ptr())
                else:
                    self.ren.render_device(self.lights)
                s = self.ren.stats()
                build.append(s["ms_grid_build"])
                render.append(s["ms_render"])
        return sum(build) / len(build), sum(render) / len(render)

    def device_frame_sha(self):
        """sha256 of the frame the last resident step left in HBM on this rank."""
        if self.exchange == "nccl":
            got = self.frame.cpu().numpy()
        else:
            got = self.ren.read_frame()
            self.ren.sync()
        return hashlib.sha256(self.np.ascontiguousarray(got).tobytes()).hexdigest()

    def close(self):
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize(self.env["dev"])
        if self.env["world"] > 1:  # nobody frees a frame another rank still has mapped
            dist.barrier()
        self.ren.close()
        if self.env["world"] > 1:
            dist.barrier()


def series_other(env, args, ops_tab, name):
    """The device-resident step of another BASELINE.json workload (configs[2]: c3; configs[4]: c5b / c5) at this N."""
    import torch
    import torch.distributed as dist
    job = Job(env, name, args.exchange, (ops_tab.get(name) or {}).get("frame_sha256"))
    steps = max(3, min(args.steps, args.steps_8k))
    with torch.cuda.stream(env["stream"]):
        for _ in range(3):                       # first frames: graph capture + tile costs for the CTA order
            job.step_resident()
    torch.cuda.synchronize(env["dev"])
    ms, _, _, per = job.timed(job.step_resident, steps, 3)
    sha = job.device_frame_sha() if (env["rank"] == 0 or job.exchange in ("peer", "nccl")) else None
    build_ms, render_ms = job.kernel_times(3)
    mine = {"rank": env["rank"], "render_kernel_ms": round(render_ms, 4), "scene_loader_ms": round(build_ms, 4),
            "step_ms": round(sum(per) / len(per), 4), "frame_sha256": sha}
    ranks = [mine]
    if env["world"] > 1:
        ranks = [None] * env["world"]
        dist.all_gather_object(ranks, mine)
    want = (ops_tab.get(name) or {}).get("frame_sha256")
    out = None
    if env["rank"] == 0:
        k = [r["render_kernel_ms"] for r in ranks]
        out = {"workload": f"{name}: {job.desc}", "view": [job.W, job.H, job.L], "n_entities": int(len(job.boxes)),
               "n_lights": int(len(job.lights)), "steps": steps, "ms_per_step": round(ms / steps, 4),
               "value": round(job.rays_frame * steps / ms / 1e3, 1), "unit": "Mrays/s",
               "frames_per_s": round(1e3 * steps / ms, 2),
               "render_kernel_ms_per_rank": k, "render_kernel_ms_min": min(k), "render_kernel_ms_max": max(k),
               "scene_loader_ms": ranks[0]["scene_loader_ms"], "exchange": job.exchange, "stripe_split": job.split,
               "stripe_split_fallback": job.split_fallback,
               "frame_check": {"oracle_frame_sha256": want,
                               "device_frames_equal_oracle": [r["frame_sha256"] == want for r in ranks
                                                              if r["frame_sha256"] is not None]}}
        if want and not all(out["frame_check"]["device_frames_equal_oracle"]):
            raise SystemExit(f"bench.py: {name} frame differs from the committed oracle hash: {out['frame_check']}")
    job.close()
    return out


def c4_sequence(env, golden_sha):
    """BASELINE.json configs[3]: 240 frames of key script D at 1920x1080 — per frame a scene update, a
    render and the frame read back to the host, two frames in flight; then the debug overlay from the cursor
    probe (alternative.cpp:762-772), as the reference's frame carries it.  Two upload modes: the whole
    2.6 MB scene every frame (par_submit_frame; what alternative.cpp:689-693 does) and the 16-byte record of
    the one entity that moved (par_submit_update).  The hash file of each mode (FNV-1a-64 per frame) must have
    the sha256 of the REAL reference's."""
    import numpy as np
    import torch
    par = env["par"]
    W, H, L, frames = 1920, 1080, 1080, 240
    scene0, light0 = par.scene_default(), par.light_default()
    h_boxes = [par.pinned_empty(len(scene0), par.AABB) for _ in range(2)]
    out = [par.pinned_empty((H, W), par.COLOR) for _ in range(2)]
    res = {}
    with par.Renderer(W, H, L, device=env["local"]) as ren:
        ren.set_atlas()
        ren.set_cursor(0, 0)

        def run(incremental, want_hash):
            player = scene0[0:1].copy()
            lights = light0.copy()
            for b in h_boxes:
                b[:] = scene0
            ren.set_scene(h_boxes[0])
            ren.sync()
            light_of, lines = [None, None], []
            for f in range(frames + 1):
                if f < frames:
                    for key in script_d_keys(f):
                        par.apply_key(key, player, lights)
                    light_of[f & 1] = lights.copy()
                    if incremental:
                        ren.submit_update(0, player, lights, out[f & 1])
                    else:
                        h_boxes[f & 1][0] = player[0]
                        ren.submit_frame(h_boxes[f & 1], lights, out[f & 1])
                if f >= 1:
                    g = f - 1
                    ren.wait_frame()
                    par.draw_overlay_at(W, H, ren.cursor_pixel(), light_of[g & 1], out[g & 1])
                    if want_hash:
                        lines.append("%03d %016x\n" % (g, par.fnv1a64(out[g & 1])))
            return lines

        for mode, incremental in (("incremental_update", True), ("full_upload", False)):
            run(incremental, False)              # warm-up pass
            torch.cuda.synchronize(env["dev"])
            t0 = time.perf_counter()
            run(incremental, False)
            dt = time.perf_counter() - t0
            sha = hashlib.sha256("".join(run(incremental, True)).encode()).hexdigest()
            res[mode] = {"frames_per_s": round(frames / dt, 1), "ms_per_frame": round(dt / frames * 1e3, 4),
                         "h2d_bytes_per_frame": 16 + 8 if incremental else int(h_boxes[0].nbytes) + 8,
                         "d2h_bytes_per_frame": W * H * 4 + 28,
                         "hash_file_sha256_equals_reference": sha == golden_sha}
    ok = all(m["hash_file_sha256_equals_reference"] for m in res.values())
    # the same sequence driven by the compiled host (host/par_headless.cpp on par::FrameRenderer — the reference's
    # main loop is C++ too): informational, its own wall clock, its own hash file
    cpp = None
    exe = os.path.join(PKG, "build", "par_headless")
    if os.path.exists(exe):
        try:
            import re
            args = [exe, "--view", str(W), str(H), str(L), "--frames", str(frames), "--script", "D", "--device", str(env["local"])]
            subprocess.run(args + ["--no-hash"], check=True, capture_output=True, timeout=120)  # warm-up (module load, clocks)
            fast = subprocess.run(args + ["--no-hash"], check=True, capture_output=True, text=True, timeout=120)
            hashed = subprocess.run(args, check=True, capture_output=True, timeout=120)
            m = re.search(r"= ([0-9.]+) frames/s", fast.stderr)
            cpp = {"frames_per_s": float(m.group(1)) if m else None,
                   "hash_file_sha256_equals_reference": hashlib.sha256(hashed.stdout).hexdigest() == golden_sha,
                   "host": "par_headless (C++, par::FrameRenderer::submit_frame_moved / wait_frame, cursor probe, overlay), "
                           "a separate process, wall clock of its frame loop with no warm-up pass (the first frames carry the graph captures)"}
        except Exception as e:  # informational only
            cpp = {"unavailable": str(e)[:160]}
    line = {"workload": "c4: 240 frames of key script D, default scene at 1920x1080 (player and light move)",
            "frames": frames, "frames_per_s": res["incremental_update"]["frames_per_s"],
            "ms_per_frame": res["incremental_update"]["ms_per_frame"], "modes": res,
            "timing": "wall clock of the host loop (scene update, submit, wait, cursor probe, overlay), two frames in "
                      "flight; hashes taken in a separate untimed pass",
            "api": {"incremental_update": "par_submit_update / par_wait_frame", "full_upload": "par_submit_frame / par_wait_frame"},
            "reference_hash_file_sha256": golden_sha, "all_hash_files_equal_reference": ok, "cpp_host": cpp}
    if not ok:
        raise SystemExit(f"bench.py: C4 sequence hashes differ from the reference's: {line}")
    return line


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

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    stream = torch.cuda.Stream(device=dev)
    env = {"par": par, "world": world, "rank": rank, "local": local, "dev": dev, "stream": stream,
           "flush": torch.empty(256 << 20, dtype=torch.uint8, device=dev),  # > 126 MB L2
           "barrier": barrier}
    flush = env["flush"]
    ops_tab = load_json(os.path.join(ROOT, "tests", "golden", "workload_ops.json"), {}) or {}

    job = Job(env, args.workload, args.exchange, (ops_tab.get(args.workload) or {}).get("frame_sha256"))
    ren, W, H, L, lights, boxes, h_boxes = job.ren, job.W, job.H, job.L, job.lights, job.boxes, job.h_boxes
    n_lights, rays_frame, exchange = len(lights), job.rays_frame, job.exchange
    from par_b200.bands import owned_rects
    my_px = sum((r1 - r0) * (c1 - c0) for r0, r1, c0, c1 in owned_rects(W, H, world, rank, job.split))
    trace(f"exchange = {exchange}")
    h_frame = par.pinned_empty((H, W), par.COLOR) if rank == 0 else None
    token = torch.zeros(1, dtype=torch.int32, device=dev)

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

    def step_e2e_sync():
        """The blocking drop-in shape: par_set_scene + par_render (N = 1), or per rank set_scene +
        render + stripe readback into the shared host frame and a barrier (N > 1)."""
        ren.set_scene(h_boxes)                   # H2D from pinned memory + scene loader
        if world == 1:
            ren.render(lights, out=h_frame)      # render + D2H into a host frame
            return
        if shared is not None:
            ren.render_device(lights)            # my stripes into my own frame: no GPU-to-GPU exchange needed
            ren.read_stripes(shared)             # ... and from there into the shared host frame (my PCIe link)
        else:                                    # no shared host frame: exchange on the GPUs, rank 0 copies the frame
            job.step_resident()
            if rank == 0 and exchange == "nccl":
                torch.from_numpy(h_frame.view(np.uint8).reshape(-1)).copy_(job.frame, non_blocking=True)
            elif rank == 0:
                ren.read_frame(h_frame)
        dist.all_reduce(token)                   # the host frame is complete when every rank's DMA is

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

    sampler = ClockSampler(local) if rank == 0 else None
    with torch.cuda.stream(stream):
        job.step_resident()
        trace("first step enqueued")
        torch.cuda.synchronize(dev)
        # untimed pre-warm: lets the SM clock ramp from idle and gives nvidia-smi time to sample.
        # The number of steps is the same on every rank (the steps depend on each other).
        t_a = time.perf_counter()
        for _ in range(5):
            job.step_resident()
        torch.cuda.synchronize(dev)
        n_pre = torch.tensor([int(args.prewarm_ms / 1e3 / max(time.perf_counter() - t_a, 1e-6) * 5) + 1],
                             dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(n_pre, op=dist.ReduceOp.MAX)
        for i in range(min(int(n_pre.item()), 100000)):
            job.step_resident()
            if i % 8 == 7:
                torch.cuda.synchronize(dev)
    torch.cuda.synchronize(dev)
    trace("pre-warm done")

    t_load0 = time.perf_counter()
    ms, t0, t1, per_step = job.timed(job.step_resident, args.steps, args.warmup)
    launches_step = job.launches_per_step()
    trace("resident timing done")
    dev_sha = job.device_frame_sha() if (rank == 0 or exchange in ("peer", "nccl")) else None
    e2e_steps_sync = max(3, min(args.steps, 50))
    ms_e2e_sync, _, t1, _ = job.timed(step_e2e_sync, e2e_steps_sync, 2)
    pipelined = world == 1 or shared2 is not None
    if pipelined:
        ms_e2e, _, t1 = timed_pipelined(args.steps, max(2, args.warmup // 2))
        e2e_steps = args.steps
    else:
        ms_e2e, e2e_steps = ms_e2e_sync, e2e_steps_sync
    trace("e2e timing done")
    clocks = sampler.stop(t_load0 - args.prewarm_ms / 1e3, t1) if sampler else None  # pre-warm + timed regions: under load

    # per-kernel times (un-graphed steps with timing events), for the roofline of the dominant kernel
    build_ms, render_ms = job.kernel_times(max(3, min(args.steps, 30)))
    per_rank = [{"rank": rank, "render_kernel_ms": round(render_ms, 4), "scene_loader_ms": round(build_ms, 4),
                 "frame_sha256": dev_sha}]
    if world > 1:
        mine = per_rank[0]
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)

    # correctness of what was timed (untimed): the frame the last e2e step left in host memory and the device
    # frame of the last resident step equal the COMMITTED oracle frame (tests/golden/workload_ops.json) and a
    # fresh 1-context render
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    frame_check = None
    if rank == 0:
        want_sha = (ops_tab.get(args.workload) or {}).get("frame_sha256")
        with par.Renderer(W, H, L, device=local) as chk:
            chk.set_atlas()
            chk.set_scene(boxes)
            want, _ = chk.render(lights)
        one_sha = hashlib.sha256(want.tobytes()).hexdigest()
        got_host = shared if shared is not None else h_frame
        host_sha = hashlib.sha256(np.ascontiguousarray(got_host).tobytes()).hexdigest()
        dev_shas = [r["frame_sha256"] for r in per_rank if r["frame_sha256"] is not None]
        frame_check = {"oracle_frame_sha256": want_sha,
                       "host_frame_equals_oracle": host_sha == want_sha if want_sha else None,
                       "device_frames_equal_oracle": [s == want_sha for s in dev_shas] if want_sha else None,
                       "host_frame_equals_1ctx_render": host_sha == one_sha,
                       "device_frames_equal_1ctx_render": [s == one_sha for s in dev_shas]}
        bad = host_sha != one_sha or any(s != one_sha for s in dev_shas) or (want_sha and one_sha != want_sha)
        if bad:
            raise SystemExit(f"bench.py: timed frames differ from the oracle / a 1-context render: {frame_check}")
    job.close()
    trace("main workload done")

    # ---- the 8K scaling workloads (BASELINE.json configs[4]) in the same line, at every N ----
    scale_8k = None
    if not args.no_scale_8k and args.workload not in ("c5", "c5b"):
        scale_8k = {}
        for name in ("c5b", "c5"):
            r8 = series_other(env, args, ops_tab, name)
            if rank == 0:
                scale_8k[name] = r8
            trace(f"8K series {name} done")
    # ---- BASELINE.json configs[2]: dense synthetic scene, 16 lights with shadow rays, at 3840x2160 ----
    c3 = None
    if not args.no_scale_8k and args.workload != "c3":
        c3 = series_other(env, args, ops_tab, "c3")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the 240-frame sequence (BASELINE.json configs[3]) ----
    c4 = None
    if world == 1 and not args.no_c4:
        golden = load_json(os.path.join(ROOT, "tests", "golden", "reference_hashes.json"), {}) or {}
        c4 = c4_sequence(env, (golden.get("tier1_1920x1080x1080_scriptD_240") or {}).get("hash_file_sha256"))
        trace("c4 sequence done")

    value = rays_frame * args.steps / ms / 1e3
    e2e = rays_frame * e2e_steps / ms_e2e / 1e3
    peaks = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"), {}) or {}
    prop = torch.cuda.get_device_properties(dev)
    sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    peak_ops = prop.multi_processor_count * 128 * sm_max * 1e6 / 1e12  # T lane-ops/s
    frac_rows = my_px / (W * H)
    roofline = {"bound": "fp32_alu", "kernel": "k_tile (primary rays + shadow walks + shading + RGBA8 pack, one CTA per 40x40 tile)",
                "unit": "Tlane-op/s", "peak": round(peak_ops, 2),
                "peak_source": f"{prop.multi_processor_count} SMs x 128 lanes x {sm_max:.0f} MHz = one warp instruction per "
                               "SM sub-partition per clock (1 op/lane/clk; no FMA on this path, -fmad=false)",
                "ms_per_launch": round(render_ms, 4), "achieved": None, "frac": None, "traffic": None}
    prof = load_json(ROOFLINE_PROFILE, {}) or {}
    cap = (prof.get("kernels") or {}).get(args.workload)
    src_sha = kernel_source_sha()
    if cap and prof.get("source_sha") == src_sha:
        # (N > 1: rank 0 renders its share of the tile rows; the capture is of the whole frame on one GPU)
        ex = cap["warp_instructions"] * frac_rows * 32 / (render_ms * 1e-3) / 1e12
        roofline.update({"achieved": round(ex, 2), "frac": round(ex / peak_ops, 3),
                         "warp_instructions_per_launch": cap["warp_instructions"] * frac_rows,
                         "traffic": (cap["dram_bytes_read"] + cap["dram_bytes_write"]) if world == 1 else None,
                         "ncu": {k: cap.get(k) for k in ("kernel", "issue_active_pct", "warps_active_pct", "duration_us_under_ncu",
                                                         "registers_per_thread", "source")},
                         "note": "frac = executed warp instructions (ncu smsp__inst_executed.sum, capture of this very "
                                 "build: source hash matches) x 32 lanes / live CUDA-event kernel time / peak = the "
                                 "fraction of issue slots the kernel fills"
                                 + ("" if world == 1 else f"; at N = {world} the one-GPU capture's count is scaled by rank 0's "
                                    "share of the tile rows (an estimate: tiles differ in cost)")})
    else:
        roofline["note"] = ("no ncu capture of this build for this workload/N (profiles/roofline_r02.json source hash "
                            f"{prof.get('source_sha')} vs built {src_sha}): hardware fraction not quoted")
    ops = (ops_tab.get(args.workload) or {}).get("algorithmic_ops")
    if ops:
        alg = ops * frac_rows / (render_ms * 1e-3) / 1e12
        roofline["algorithmic_speedup"] = round(alg / peak_ops, 3)
        roofline["algorithmic"] = {"ops_per_launch": ops * frac_rows, "achieved": round(alg, 2),
                                   "note": "reference-equivalent lane-ops (SURVEY.md §8d weights x oracle counters: what the "
                                           "reference's loops execute for this frame) / kernel time; algorithmic_speedup = "
                                           "that / peak, > 1 because one grid walk serves a whole tile x z-group, probes "
                                           "are de-duplicated, the shaft cull drops boxes no ray of a group can hit (Q19)"}
    hbm_bytes = 4.0 * my_px + 16.0 * len(boxes)
    roofline["hbm"] = {"algorithmic_bytes_per_launch": hbm_bytes,
                       "achieved_gbs": round(hbm_bytes / (render_ms * 1e-3) / 1e9, 1), "peak_gbs": peaks.get("hbm_gbs"),
                       "note": "RGBA8 frame written + scene read (the G-buffer never leaves the SM); not the bound"}

    k_render = [r["render_kernel_ms"] for r in per_rank]
    step_mean = sum(per_step) / len(per_step)
    line = {
        "metric": "Mrays/s", "value": round(value, 1), "unit": "Mrays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {job.desc}", "view": [W, H, L], "n_entities": int(len(boxes)),
                   "n_lights": int(n_lights), "rays_per_frame": rays_frame, "stripe_split": job.split,
                   "stripe_split_fallback": job.split_fallback,
                   "parallelism": ("1 GPU" if world == 1 else
                                   f"interleaved 40-row stripes x{world}, frame exchange fused into the render kernel "
                                   "(peer-memory stores over NVLink, arrival/credit flags in the frame footers, no "
                                   "collective), frame complete on "
                                   + ("rank 0 (gather-to-root)" if exchange == "root" else "every GPU (all-gather)")
                                   if exchange in ("peer", "root") else
                                   f"interleaved 40-row stripes x{world} + in-place NCCL all-gather of the RGBA8 frame"),
                   "step": "par_render_resident: one CUDA graph launch = render kernel of this frame + (side branch) device "
                           "scene loader rebuilding the other grid generation from the resident scene for the next frame",
                   "l2": "flushed between timed steps (256 MB fill)", "frames_per_s": round(1e3 * args.steps / ms, 2)},
        "e2e": {"value": round(e2e, 1), "unit": "Mrays/s", "ms_per_step": round(ms_e2e / e2e_steps, 4),
                "frames_per_s": round(1e3 * e2e_steps / ms_e2e, 2), "steps": e2e_steps,
                "h2d_bytes_per_step": int(h_boxes.nbytes) * world, "d2h_bytes_per_step": int(H * W * 4),
                "readback": ("each rank DMAs its own stripes into one shared pinned host frame (N PCIe links)"
                             if shared is not None else "rank 0 / the one context copies the whole frame"),
                "api": ("par_submit_frame / par_wait_frame, two frames in flight (readback of frame k beside upload "
                        "and kernels of frame k+1, the frame's GPU work as one CUDA graph launch); timed from an "
                        "empty pipeline to the last frame complete on the host, L2 flush inside the timed region"
                        + ("" if world == 1 else "; every rank ships its own stripes, a 4-byte all-reduce per frame "
                                                 "signals completion") if pipelined else
                        "par_set_scene + par_render_device + par_read_stripes per rank, barrier per step"),
                "sync_call_ms": round(ms_e2e_sync / e2e_steps_sync, 4),
                "sync_call": ("par_set_scene + par_render (blocking drop-in call) per step" if world == 1 else
                              "par_set_scene + par_render_device + par_read_stripes per rank, barrier per step")},
        "gpu_launches": launches_step * args.steps * world,
        "gpu_launches_per_step_per_rank": launches_step,
        "kernels_ms": {"scene_loader": round(build_ms, 4), "k_tile": round(render_ms, 4),
                       "k_tile_per_rank": k_render, "k_tile_min": min(k_render), "k_tile_max": max(k_render),
                       "step_ms": round(step_mean, 4),
                       "step_minus_render_kernel_ms": round(step_mean - render_ms, 4),
                       "note": "scene_loader and k_tile are CUDA-event times of un-graphed launches; in the timed step the "
                               "loader rebuilds the other grid generation on a side branch of the graph, beside the render "
                               "kernel, so rank 0's step minus its render kernel = launch gaps + (N > 1) the exchange flags "
                               "and waiting for the slowest rank's arrival"},
        "roofline": roofline, "clocks": clocks, "frame_check": frame_check,
    }
    if scale_8k is not None:
        line["scale_8k"] = scale_8k
    if c3 is not None:
        line["c3_series"] = c3
    if c4 is not None:
        line["c4_sequence"] = c4
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
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-budget", type=float, default=60.0, help="seconds of CPU work for the CPU legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scale-8k", action="store_true", help="skip the 8K series (scale_8k) and the c3 series")
    ap.add_argument("--no-c4", action="store_true", help="skip the 240-frame sequence (c4_sequence)")
    ap.add_argument("--steps-8k", type=int, default=20, help="timed steps of each 8K series")
    ap.add_argument("--exchange", default="root", choices=["root", "peer", "nccl"],
                    help="N>1 frame exchange, fused into the render kernel as peer-memory stores + flags: 'root' gathers "
                         "the frame on rank 0 (default; SURVEY.md 8e gather-to-root), 'peer' completes it on every GPU; "
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
