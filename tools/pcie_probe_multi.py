"""Development aid: what the host<->device links of this box can do when every GPU copies at once
(pinned, 128 MB chunks, both directions).  Launch: python tools/pcie_probe_multi.py [n_gpus]"""
import multiprocessing as mp
import sys
import time


def worker(rank, n, barrier, q):
    import torch
    torch.cuda.set_device(rank)
    sz = 1 << 30
    a = torch.empty(sz, dtype=torch.uint8).pin_memory()
    b = torch.empty(sz, dtype=torch.uint8).pin_memory()
    a.fill_(1)
    d1 = torch.empty(sz, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(sz, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    chunk = 128 << 20

    def run(h2d, d2h, reps=4):
        for _ in range(reps):
            for o in range(0, sz, chunk):
                if h2d:
                    with torch.cuda.stream(s1):
                        d1[o:o + chunk].copy_(a[o:o + chunk], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2):
                        b[o:o + chunk].copy_(d2[o:o + chunk], non_blocking=True)
        torch.cuda.synchronize()

    out = {}
    for name, h2d, d2h in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        run(h2d, d2h, 1)
        barrier.wait()
        t0 = time.perf_counter()
        run(h2d, d2h)
        dt = time.perf_counter() - t0
        barrier.wait()
        out[name] = 4 * sz / dt / 1e9
    q.put((rank, out))


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(n), ctx.Queue()
    ps = [ctx.Process(target=worker, args=(r, n, barrier, q)) for r in range(n)]
    for p in ps:
        p.start()
    res = sorted(q.get() for _ in range(n))
    for p in ps:
        p.join()
    for name in ("h2d", "d2h", "both"):
        vals = [r[1][name] for r in res]
        print("%d GPUs at once, %-4s: per GPU %s GB/s per direction | sum %.1f GB/s per direction" % (
            n, name, " ".join("%.1f" % v for v in vals), sum(vals)))
