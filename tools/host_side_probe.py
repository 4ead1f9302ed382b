"""What the host side of a GPU box can take: topology, /dev/shm, huge pages, the window-placement
access pattern (tools/host_probe.c) alone and next to a saturating pinned D2H stream."""
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sh(cmd):
    print("$ " + cmd, flush=True)
    r = subprocess.run(cmd, shell=True, capture_output=True, text=True)
    print((r.stdout + r.stderr).rstrip(), flush=True)


def main():
    sh("lscpu | head -30")
    sh("nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; taskset -p $$")
    sh("free -g; df -h /dev/shm /tmp")
    sh("cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/defrag")
    sh("numactl -H 2>/dev/null | head -20; nvidia-smi topo -m 2>/dev/null | head -30")
    exe = os.path.join(ROOT, "tools", "host_probe")
    sh("gcc -O2 -pthread -o %s %s" % (exe, os.path.join(ROOT, "tools", "host_probe.c")))
    sh("%s 3.0" % exe)

    import torch
    n = 256 * 1024 * 1024 // 4
    dev = torch.empty(n, dtype=torch.float32, device="cuda")
    host = torch.empty(8 * n, dtype=torch.float32, pin_memory=True)
    stop = threading.Event()
    moved = [0, 0.0]

    def d2h_loop():
        s = torch.cuda.Stream()
        t0 = time.perf_counter()
        while not stop.is_set():
            with torch.cuda.stream(s):
                for i in range(8):
                    host[i * n:(i + 1) * n].copy_(dev, non_blocking=True)
            s.synchronize()
            moved[0] += 8 * n * 4
        moved[1] = time.perf_counter() - t0

    # D2H alone
    th = threading.Thread(target=d2h_loop)
    th.start()
    time.sleep(2.0)
    stop.set()
    th.join()
    print("pinned D2H alone: %.1f GB/s" % (moved[0] / moved[1] / 1e9), flush=True)
    # D2H next to the placement pattern
    for threads in (8, 16, len(os.sched_getaffinity(0))):
        stop.clear()
        moved[0], moved[1] = 0, 0.0
        th = threading.Thread(target=d2h_loop)
        th.start()
        time.sleep(0.3)
        r = subprocess.run([exe, "3.0", str(threads)], capture_output=True, text=True)
        stop.set()
        th.join()
        print("--- placement with %d threads NEXT TO a D2H stream running at %.1f GB/s" % (threads, moved[0] / moved[1] / 1e9))
        print("\n".join(l for l in r.stdout.splitlines() if l.startswith("warm")), flush=True)


if __name__ == "__main__":
    sys.exit(main())
