"""Development aid: with-copy throughput of hdb_deblock_batch against the slice size of the host pipeline."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_video_codec_b200 as hdb  # noqa: E402

w, h, n = 3840, 2160, int(os.environ.get("N", 256))
fb = 3 * w * h // 2
nv, nh = hdb.num_vert_bs(w, h), hdb.num_hor_bs(w, h)
g = torch.Generator().manual_seed(0)
h_in = torch.randint(0, 256, (n, fb), dtype=torch.uint8, generator=g).pin_memory()
h_out = torch.empty((n, fb), dtype=torch.uint8).pin_memory()
qp = torch.randint(18, 52, (n,), dtype=torch.int32, generator=g).pin_memory()
v = torch.randint(0, 3, (n, nv), dtype=torch.uint8, generator=g).pin_memory()
hh = torch.randint(0, 3, (n, nh), dtype=torch.uint8, generator=g).pin_memory()
for mb in [int(x) for x in (sys.argv[1:] or ["16", "32", "64", "128", "256", "512"])]:
    os.environ["HDB_CHUNK_MB"] = str(mb)
    dev = hdb.Deblocker(0)
    tm = hdb.Timing()
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter()
        dev.deblock_batch(h_in, w, h, qp, v, hh, out=h_out, timing=tm, n_frames=n)
        best = min(best, time.perf_counter() - t0)
    print("chunk %4d MB: %.1f ms  %.1f Mpixel/s  %.1f GB/s per direction | kernel_ms %.1f copy_ms %.1f launches %d" % (
        mb, best * 1e3, n * w * h / best / 1e6, n * fb / best / 1e9, tm.kernel_ms, tm.copy_ms, tm.kernel_launches), flush=True)
    dev.close()
