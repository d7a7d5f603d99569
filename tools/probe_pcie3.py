"""Developer probe: does the backing of the pinned scene buffer (cudaMallocHost vs 2 MB-aligned
transparent-huge-page memory + cudaHostRegister) change the H2D time of a 2.6 MB / 33 MB upload?"""
import ctypes as C, mmap, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pixel-art-raytracer_b200"))
import numpy as np
import torch
import par_b200 as par

dev = torch.device("cuda", 0)
libc = C.CDLL("libc.so.6", use_errno=True)
print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())


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


def huge_registered(n):
    size = (n + (2 << 20) - 1) & ~((2 << 20) - 1)
    p = C.c_void_p()
    assert libc.posix_memalign(C.byref(p), 2 << 20, size) == 0
    rc = libc.madvise(p, C.c_size_t(size), 14)  # MADV_HUGEPAGE
    buf = (C.c_uint8 * size).from_address(p.value)
    arr = np.frombuffer(buf, dtype=np.uint8)
    arr[:] = 1
    assert par.lib().par_register_host(p, size) == 0
    return arr[:n], rc


for n in (2596928, 33177600):
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for rep in range(3):
        a = par.pinned_empty(n, np.uint8)
        a[:] = 1
        ta = torch.from_numpy(a)
        t1 = ev_time(lambda: d.copy_(ta, non_blocking=True))
        t1b = ev_time(lambda: ta.copy_(d, non_blocking=True))
        b, rc = huge_registered(n)
        tb = torch.from_numpy(b)
        t2 = ev_time(lambda: d.copy_(tb, non_blocking=True))
        t2b = ev_time(lambda: tb.copy_(d, non_blocking=True))
        print(f"{n / 1e6:6.2f} MB rep {rep}: cudaMallocHost H2D {t1:.4f} D2H {t1b:.4f} | THP+register (madvise rc {rc}, pinned {tb.is_pinned()}) H2D {t2:.4f} D2H {t2b:.4f} ms")
