#!/usr/bin/env python
"""Developer probe (torchrun, N GPUs): the device-resident step of bench.py's workloads with whole tile rows
(stripe_split 1) and with the split bench.py would pick, alternated on one box.  One JSON line per run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 \
        tools/split_ab.py c5b c5 c2
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pixel-art-raytracer_b200"))
import bench  # noqa: E402


def main():
    import datetime
    import torch
    import torch.distributed as dist
    import par_b200 as par
    from par_b200.bands import stripe_split_for
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))

    def barrier():
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)

    env = {"par": par, "world": world, "rank": rank, "local": local, "dev": dev, "stream": torch.cuda.Stream(device=dev),
           "flush": torch.empty(256 << 20, dtype=torch.uint8, device=dev), "barrier": barrier}
    ops_tab = bench.load_json(os.path.join(ROOT, "tests", "golden", "workload_ops.json"), {}) or {}

    class Args:
        exchange = os.environ.get("EXCHANGE", "root")
        steps = int(os.environ.get("STEPS", "20"))
        steps_8k = steps

    names = sys.argv[1:] or ["c5b", "c5", "c2"]
    for rep in range(int(os.environ.get("REPS", "2"))):
        for name in names:
            W, H = bench.WORKLOADS[name][0], bench.WORKLOADS[name][1]
            auto = stripe_split_for(W, H, world)
            for split in sorted({1, auto}):
                os.environ["PAR_BENCH_STRIPE_SPLIT"] = str(split)
                out = bench.series_other(env, Args, ops_tab, name)
                if rank == 0:
                    print(json.dumps({"n_gpus": world, "workload": name, "rep": rep, "stripe_split": out["stripe_split"],
                                      "fallback": out["stripe_split_fallback"], "ms_per_step": out["ms_per_step"],
                                      "k_tile_min": out["render_kernel_ms_min"], "k_tile_max": out["render_kernel_ms_max"],
                                      "frames_ok": out["frame_check"]["device_frames_equal_oracle"]}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
