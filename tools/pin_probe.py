"""How long does it take to get 0.71 GB of page-locked, device-visible host memory?
cudaHostAlloc (torch pin_memory) against cudaHostRegister of a huge-page advised anonymous mapping."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
torch.zeros(1, device="cuda")
from f2cnn_b200 import engine
N = 712_231_424 // 4
rt = torch.cuda.cudart()
for rep in range(3):
    t = time.perf_counter(); a = torch.empty(N, dtype=torch.float32, pin_memory=True); t1 = time.perf_counter() - t
    del a
    t = time.perf_counter()
    h = engine.host_empty((N,), np.float32)
    h[::1024] = 0   # touch every page (4 KiB stride)
    t2 = time.perf_counter() - t
    t = time.perf_counter()
    rc = rt.cudaHostRegister(h.ctypes.data, h.nbytes, 1 | 2)   # portable | mapped
    t3 = time.perf_counter() - t
    th = torch.from_numpy(h)
    d = torch.empty(1 << 20, dtype=torch.float32, device="cuda").normal_()
    th[:1 << 20].copy_(d, non_blocking=True); torch.cuda.synchronize()
    ok = bool(torch.equal(th[:1 << 20], d.cpu()))
    print("cudaHostAlloc %.0f ms | huge-page mapping + touch %.0f ms + cudaHostRegister %.0f ms (rc %s, is_pinned %s, copy ok %s)" % (
        t1 * 1e3, t2 * 1e3, t3 * 1e3, rc, th.is_pinned(), ok), flush=True)
    rt.cudaHostUnregister(h.ctypes.data)
    del th, h
    engine.release_host_pool()
