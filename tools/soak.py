"""Development aid: a longer randomised parity soak than the test suite affords (run on the GPU box).
Random geometry, QP, content, boundary strengths; single frames, host batches, device batches (in place and
out of place), pitched surfaces and logical-shard strips, all against the oracle."""
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_video_codec_b200 as hdb  # noqa: E402
from oracle import oracle_py as orc  # noqa: E402
from tests import synth  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
rng = np.random.default_rng(seed)
dev = hdb.Deblocker(0)
multi = hdb.MultiDeblocker([0, 0, 0])
kinds = ["blocky", "random", "steps", "saturated", "ramp", "flat"]
bs_kinds = [None, "uniform012", "wide", "twos", "zeros", "ones"]
t0, n_cases, fails = time.time(), 0, 0
st = torch.cuda.current_stream().cuda_stream
while time.time() - t0 < budget:
    mode = int(rng.integers(5))
    w = 8 * int(rng.integers(1, 200))
    h = 8 * int(rng.integers(1, 80))
    if mode in (2, 3, 4) or rng.integers(2):
        w = 32 * max(1, w // 32)
    n = int(rng.integers(1, 7))
    kind = kinds[int(rng.integers(len(kinds)))]
    bs_kind = bs_kinds[int(rng.integers(len(bs_kinds)))]
    frames = np.stack([synth.make_frame(rng, w, h, kind) for _ in range(n)])
    qps = rng.integers(10, 60, n).astype(np.uint32)
    if bs_kind is None:
        v = hh = None
    else:
        vs, hs = zip(*[synth.make_bs(rng, w, h, bs_kind) for _ in range(n)])
        v, hh = np.stack(vs), np.stack(hs)
    want = orc.deblock_batch(frames, w, h, qps, v, hh)
    try:
        if mode == 0:
            got = np.stack([dev.deblock_frame(frames[i], w, h, int(qps[i]), None if v is None else v[i], None if hh is None else hh[i]) for i in range(n)])
        elif mode == 1:
            got = dev.deblock_batch(frames, w, h, qps, v, hh)
        elif mode == 2:
            d = torch.from_numpy(frames).cuda()
            o = d if rng.integers(2) else torch.empty_like(d)
            dq = torch.from_numpy(qps.astype(np.int32)).cuda()
            dv = None if v is None else torch.from_numpy(v).cuda()
            dh = None if hh is None else torch.from_numpy(hh).cuda()
            dev.deblock_batch_device(d, o, n, w, h, dq, dv, dh, stream=st)
            torch.cuda.synchronize()
            got = o.cpu().numpy()
        elif mode == 3:
            cw, ch = w // 2, h // 2
            py, pc = w + 16 * int(rng.integers(0, 5)), cw + 16 * int(rng.integers(0, 5))
            t = torch.from_numpy(frames).cuda()
            P = [torch.zeros((n, h, py), dtype=torch.uint8, device="cuda"), torch.zeros((n, ch, pc), dtype=torch.uint8, device="cuda"),
                 torch.zeros((n, ch, pc), dtype=torch.uint8, device="cuda")]
            P[0][:, :, :w] = t[:, :w * h].view(n, h, w)
            P[1][:, :, :cw] = t[:, w * h:w * h + cw * ch].view(n, ch, cw)
            P[2][:, :, :cw] = t[:, w * h + cw * ch:].view(n, ch, cw)
            s = hdb.Surfaces(P[0].data_ptr(), P[1].data_ptr(), P[2].data_ptr(), py, pc, h * py, ch * pc)
            dq = torch.from_numpy(qps.astype(np.int32)).cuda()
            dv = None if v is None else torch.from_numpy(v).cuda()
            dh = None if hh is None else torch.from_numpy(hh).cuda()
            dev.deblock_surfaces_device(s, s, n, w, h, dq, dv, dh, stream=st)
            torch.cuda.synchronize()
            got = torch.cat([P[0][:, :, :w].reshape(n, -1), P[1][:, :, :cw].reshape(n, -1), P[2][:, :, :cw].reshape(n, -1)], 1).cpu().numpy()
        else:
            got = np.stack([multi.deblock_frame_strips(frames[i], w, h, int(qps[i]), None if v is None else v[i], None if hh is None else hh[i]) for i in range(n)])
        ok = np.array_equal(got, want)
    except Exception as e:  # noqa: BLE001
        ok = False
        print("EXCEPTION", e)
    n_cases += 1
    if not ok:
        fails += 1
        print("MISMATCH mode=%d %dx%d n=%d kind=%s bs=%s qps=%s" % (mode, w, h, n, kind, bs_kind, qps.tolist()), flush=True)
print("soak seed %d: %d cases, %d failures, %.0f s" % (seed, n_cases, fails, time.time() - t0))
sys.exit(1 if fails else 0)
