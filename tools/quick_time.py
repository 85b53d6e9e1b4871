"""Quick kernel-only timing of the device-resident path (development aid; bench.py is the contract)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_video_codec_b200 as hdb  # noqa: E402


def device_blocky(n, w, h, seed=0):
    """'blocky' content (tests/synth.py recipe) built on the device: uint8[n, 3wh/2]."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    planes = []
    for (pw, ph) in ((w, h), (w // 2, h // 2), (w // 2, h // 2)):
        xx = torch.arange(pw, device="cuda", dtype=torch.float32)[None, None, :] / pw
        yy = torch.arange(ph, device="cuda", dtype=torch.float32)[None, :, None] / ph
        ramp = 96.0 + 64.0 * xx + 32.0 * yy
        bh, bw = (ph + 7) // 8, (pw + 7) // 8
        dc = torch.randint(-12, 13, (n, bh, bw), device="cuda", generator=g, dtype=torch.int16)
        dc = dc.repeat_interleave(8, 1).repeat_interleave(8, 2)[:, :ph, :pw]
        noise = torch.randint(-2, 3, (n, ph, pw), device="cuda", generator=g, dtype=torch.int16)
        planes.append((ramp + dc + noise).clamp_(0, 255).to(torch.uint8).reshape(n, -1))
    return torch.cat(planes, 1).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--w", type=int, default=3840)
    ap.add_argument("--h", type=int, default=2160)
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--inplace", action="store_true")
    ap.add_argument("--nv12", action="store_true", help="the same frames stored as NV12 (DeblockTiledKernel<., true>)")
    a = ap.parse_args()
    dev = hdb.Deblocker(0)
    w, h, n = a.w, a.h, a.n
    g = torch.Generator(device="cuda").manual_seed(1)
    frames = device_blocky(n, w, h)
    if a.nv12:
        ysz, csz = w * h, (w // 2) * (h // 2)
        frames[:, ysz:] = torch.stack([frames[:, ysz:ysz + csz], frames[:, ysz + csz:]], dim=2).reshape(n, -1)
    flags = hdb.FLAG_NV12 if a.nv12 else 0
    out = frames.clone() if a.inplace else torch.empty_like(frames)
    qp = torch.randint(18, 52, (n,), device="cuda", generator=g, dtype=torch.int32)
    nv, nh = hdb.num_vert_bs(w, h), hdb.num_hor_bs(w, h)
    vbs = torch.randint(0, 3, (n, nv), device="cuda", generator=g, dtype=torch.uint8)
    hbs = torch.randint(0, 3, (n, nh), device="cuda", generator=g, dtype=torch.uint8)
    stream = torch.cuda.current_stream().cuda_stream
    src = out if a.inplace else frames
    for _ in range(3):
        dev.deblock_batch_device_ex(src, out, n, w, h, qp, vbs, hbs, flags=flags, stream=stream)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.iters)]
    for e0, e1 in evs:
        e0.record()
        dev.deblock_batch_device_ex(src, out, n, w, h, qp, vbs, hbs, flags=flags, stream=stream)
        e1.record()
    torch.cuda.synchronize()
    ms = np.array([e0.elapsed_time(e1) for e0, e1 in evs])
    if os.environ.get("QT_VERBOSE"):
        print(" ".join("%.3f" % x for x in ms))
    cw, ch = w // 2, h // 2
    alg = 3 * w * h + nv + nh + hdb.num_vert_bs(cw, ch) + hdb.num_hor_bs(cw, ch)
    pix = n * w * h
    chk = int(out.view(torch.int64).sum().item()) & 0xffffffffffff
    print("[%s] chk=%012x " % (os.path.basename(hdb.library_path()), chk), end="")
    print("%s%dx%d n=%d inplace=%s: median %.3f ms  min %.3f ms | %.1f Mpixel/s | %.1f GB/s algorithmic | changed %.1f%%" % (
        "NV12 " if a.nv12 else "", w, h, n, a.inplace, np.median(ms), ms.min(), pix / np.median(ms) / 1e3, n * alg / np.median(ms) / 1e6,
        100.0 * float((out != frames).float().mean()) if not a.inplace else -1))


if __name__ == "__main__":
    main()
