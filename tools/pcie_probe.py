"""Development aid: what the host<->device link of this box can do (pinned, large copies), and where the GPU sits."""
import os
import subprocess
import sys
import time

import torch


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:
        return "ERR %s" % e


print(sh("nvidia-smi topo -m | head -20"))
print(sh("lscpu | grep -E 'Model name|Socket|NUMA|^CPU\\(s\\)|Thread'"))
print("affinity", len(os.sched_getaffinity(0)), "cpus")
bus = sh("nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i 0").lower()
if bus.startswith("00000000:"):
    bus = bus[4:]
print("bus", bus, "numa_node", sh("cat /sys/bus/pci/devices/%s/numa_node" % bus), "local_cpulist", sh("cat /sys/bus/pci/devices/%s/local_cpulist" % bus))
print(sh("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current --format=csv"))
print(sh("free -g | head -2"))

GB = 1 << 30
n = 4 * GB
a = torch.empty(n, dtype=torch.uint8).pin_memory()
b = torch.empty(n, dtype=torch.uint8).pin_memory()
a.fill_(1)
b.fill_(2)
d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def h2d():
    with torch.cuda.stream(s1):
        d1.copy_(a, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        b.copy_(d2, non_blocking=True)


def both():
    h2d()
    d2h()


def chunked(sz):
    def f():
        for o in range(0, n, sz):
            with torch.cuda.stream(s1):
                d1[o:o + sz].copy_(a[o:o + sz], non_blocking=True)
            with torch.cuda.stream(s2):
                b[o:o + sz].copy_(d2[o:o + sz], non_blocking=True)
    return f


print("H2D  %.1f GB/s" % (n / timed(h2d) / 1e9))
print("D2H  %.1f GB/s" % (n / timed(d2h) / 1e9))
t = timed(both)
print("both %.1f GB/s per direction (%.1f total)" % (n / t / 1e9, 2 * n / t / 1e9))
for sz in (8 << 20, 32 << 20, 128 << 20):
    t = timed(chunked(sz))
    print("both, %d MB chunks: %.1f GB/s per direction" % (sz >> 20, n / t / 1e9))
