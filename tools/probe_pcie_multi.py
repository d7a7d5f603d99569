"""Developer probe: aggregate D2H rate of N GPUs copying concurrently into host memory — N private pinned buffers
vs ONE buffer (cudaMallocHost, and a page-locked /dev/shm mapping as bench.py's e2e uses at N > 1).  Says whether
the multi-GPU e2e frame is bound by the PCIe topology / host memory or by the library."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pixel-art-raytracer_b200"))
import numpy as np
import torch
import par_b200 as par

n = torch.cuda.device_count()
total = 33177600  # the C2 frame
share = total // n
src = [torch.zeros(share, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]
streams = [torch.cuda.Stream(device=i) for i in range(n)]


def run(dsts, reps=20):
    for i in range(n):
        torch.cuda.synchronize(i)
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(n):
            with torch.cuda.stream(streams[i]):
                dsts[i].copy_(src[i], non_blocking=True)
    for i in range(n):
        streams[i].synchronize()
    dt = (time.perf_counter() - t0) / reps
    return dt * 1e3, total / dt / 1e9


private = [torch.from_numpy(par.pinned_empty(share, np.uint8)) for _ in range(n)]
one = torch.from_numpy(par.pinned_empty(total, np.uint8))
one_parts = [one[i * share:(i + 1) * share] for i in range(n)]
path = f"/dev/shm/par_probe_{os.getpid()}"
shm = par.shared_host_frame(path, 2160, 3840, create=True)
os.unlink(path)
shm_t = torch.from_numpy(shm.view(np.uint8).reshape(-1))
shm_parts = [shm_t[i * share:(i + 1) * share] for i in range(n)]
for name, d in (("private pinned buffers", private), ("one cudaMallocHost buffer", one_parts), ("one page-locked /dev/shm mapping", shm_parts)):
    run(d, 3)
    ms, gbs = run(d)
    print(f"{n} GPUs x {share / 1e6:.1f} MB D2H concurrently into {name}: {ms:.4f} ms per frame = {gbs:.1f} GB/s aggregate")
if n > 1:  # one GPU alone, for reference
    n_save = n
    n = 1
    ms, gbs = run([private[0]])
    print(f"GPU 0 alone, {share / 1e6:.1f} MB: {ms:.4f} ms = {share / ms / 1e6:.1f} GB/s")
