"""Development aid: single-frame with-copy latency of hdb_deblock_frame on pinned buffers, with the
single-slice zero-copy path on and off."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_video_codec_b200 as hdb  # noqa: E402
from oracle import oracle_py as orc  # noqa: E402
from tests import synth  # noqa: E402

for (w, h) in [(352, 288), (768, 576), (1920, 1080), (3840, 2160), (7680, 4320)]:
    rng = np.random.default_rng(w)
    frame = synth.make_frame(rng, w, h, "blocky")
    v, hh = synth.make_bs(rng, w, h, "uniform012")
    want_def = orc.deblock_frame(frame, w, h, 35)
    want_bs = orc.deblock_frame(frame, w, h, 35, v, hh)
    h_in = torch.from_numpy(frame).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    for zc in ("0", "1"):
        os.environ["HDB_ZERO_COPY"] = zc
        dev = hdb.Deblocker(0)
        tm = hdb.Timing()
        res = []
        for bs in (False, True):
            args = (v, hh) if bs else (None, None)
            for _ in range(5):
                dev.deblock_frame(h_in.numpy(), w, h, 35, *args, out=h_out.numpy(), timing=tm)
            ok = np.array_equal(h_out.numpy(), want_bs if bs else want_def)
            ts = []
            for _ in range(50):
                t0 = time.perf_counter()
                dev.deblock_frame(h_in.numpy(), w, h, 35, *args, out=h_out.numpy(), timing=tm)
                ts.append(time.perf_counter() - t0)
            res.append("%s BS: median %.1f us (kernel %.1f us) %s" % ("caller" if bs else "default", np.median(ts) * 1e6, tm.kernel_ms * 1e3, "ok" if ok else "MISMATCH"))
        print("%dx%d %s | %s | %s" % (w, h, "zero-copy where it applies (<= 4 MB)" if zc == "1" else "plain H2D, kernel, D2H", res[0], res[1]), flush=True)
        dev.close()
