"""Development aid: where a launch of DeblockTiledKernel spends its fixed time.  Needs a library built
with -DHDB_TRACE=1 (tools/build_variants.sh "trace -DHDB_TRACE=1") selected with HDB_LIBRARY."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_video_codec_b200 as hdb  # noqa: E402
from tools.quick_time import device_blocky  # noqa: E402

w, h = 3840, 2160
dev = hdb.Deblocker(0)
lib = hdb.load_library()
for n in [int(x) for x in (sys.argv[1:] or ["16", "128"])]:
    g = torch.Generator(device="cuda").manual_seed(1)
    frames = device_blocky(n, w, h)
    out = torch.empty_like(frames)
    qp = torch.randint(18, 52, (n,), device="cuda", generator=g, dtype=torch.int32)
    vbs = torch.randint(0, 3, (n, hdb.num_vert_bs(w, h)), device="cuda", generator=g, dtype=torch.uint8)
    hbs = torch.randint(0, 3, (n, hdb.num_hor_bs(w, h)), device="cuda", generator=g, dtype=torch.uint8)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(4):
        dev.deblock_batch_device(frames, out, n, w, h, qp, vbs, hbs, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dev.deblock_batch_device(frames, out, n, w, h, qp, vbs, hbs, stream=st)
    e1.record()
    torch.cuda.synchronize()
    nw = 148 * 23
    buf = np.zeros((nw, 4), np.uint64)
    rc = lib.hdb_debug_read_trace(buf.ctypes.data_as(C.c_void_p), nw)
    assert rc == 0
    t0 = buf[:, 0].min()
    entry, first, end, items = [(buf[:, i] - t0).astype(np.float64) / 1e3 for i in range(3)] + [buf[:, 3].astype(np.float64)]
    q = lambda a: "min %.1f p10 %.1f median %.1f p90 %.1f max %.1f" % (a.min(), np.percentile(a, 10), np.median(a), np.percentile(a, 90), a.max())
    print("n=%d frames: event time %.1f us" % (n, e0.elapsed_time(e1) * 1e3))
    print("  warp entry      (us after the first warp): " + q(entry))
    print("  first tile ready                          : " + q(first))
    print("  loop exit                                 : " + q(end))
    print("  items per warp                            : " + q(items))
    busy = end - first
    print("  busy time per warp                        : " + q(busy) + "  | us per item: median %.2f" % np.median(busy / np.maximum(items, 1)))
