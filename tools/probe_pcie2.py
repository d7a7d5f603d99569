"""Developer probe: H2D/D2H bandwidth vs size and CPU/NUMA placement of the pinned buffer."""
import os, subprocess, sys
import torch

print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
print("cpus allowed:", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8], "...")
try:
    for n in sorted(os.listdir("/sys/devices/system/node")):
        if n.startswith("node"):
            print(n, open(f"/sys/devices/system/node/{n}/cpulist").read().strip())
except Exception as e:
    print("no numa info", e)
dev = torch.device("cuda", 0)


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


def sweep(tag):
    for mb in (0.25, 1, 2.6, 8, 33.2, 128):
        n = int(mb * 1e6)
        h = torch.empty(n, dtype=torch.uint8).pin_memory()
        h.fill_(1)
        d = torch.empty(n, dtype=torch.uint8, device=dev)
        t1 = ev_time(lambda: d.copy_(h, non_blocking=True))
        t2 = ev_time(lambda: h.copy_(d, non_blocking=True))
        print(f"{tag} {mb:6.2f} MB  H2D {t1:.4f} ms {n / t1 / 1e6:6.1f} GB/s   D2H {t2:.4f} ms {n / t2 / 1e6:6.1f} GB/s")


sweep("default ")
