"""PCIe D2H rate of this box: one copy stream vs two, 312 MB chunks (the pipeline's sub-batch size)."""
import time, torch
n = 312 * 1024 * 1024 // 4
dev = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(2)]
host = torch.empty(24 * n, dtype=torch.float32, pin_memory=True)
def run(streams):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(24):
        with torch.cuda.stream(ss[i % streams]):
            host[i * n:(i + 1) * n].copy_(dev[i % 2], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    return 24 * n * 4 / dt / 1e9
for s in (1, 2, 3, 1, 2):
    print("streams %d: %.1f GB/s" % (s, run(s)))
# split each chunk in two halves on two streams
def run_split():
    ss = [torch.cuda.Stream() for _ in range(2)]
    torch.cuda.synchronize()
    t = time.perf_counter()
    h = n // 2
    for i in range(24):
        for k in range(2):
            with torch.cuda.stream(ss[k]):
                host[i * n + k * h:i * n + (k + 1) * h].copy_(dev[i % 2][k * h:(k + 1) * h], non_blocking=True)
    torch.cuda.synchronize()
    return 24 * n * 4 / (time.perf_counter() - t) / 1e9
print("split halves on 2 streams: %.1f GB/s" % run_split())
