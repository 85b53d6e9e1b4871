"""Development aid: SaoKernel alone on 4K frames resident in HBM, per kind of block -- every block off, band offset,
one edge class, or the bench's random mix -- so that the cost of each path of the kernel is seen by itself.
usage: [HDB_LIBRARY=...] python tools/sao_time.py [--n 256] [--iters 10] [--ctb 64]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import gpu_video_codec_b200 as hdb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--ctb", type=int, default=64)
ap.add_argument("--kinds", default="mix,off,band,eo0,eo1,eo2,eo3")
ap.add_argument("--depth", type=int, default=8, help="> 8: samples in 16-bit containers (Sao16Kernel)")
a = ap.parse_args()
w, h, n, ctb = 3840, 2160, a.n, a.ctb
dev = hdb.Deblocker(0)
st = torch.cuda.current_stream().cuda_stream
frames = bench.workload_frames(torch, 0, n, w, h, 1, "cuda")[0]
flags = 0
if a.depth > 8:
    frames = (frames.to(torch.int16) << (a.depth - 8)) | 1
    flags = hdb.flag_depth(a.depth)
out = torch.empty_like(frames)
rows, cols = (h + ctb - 1) // ctb, (w + ctb - 1) // ctb
g = torch.Generator(device="cuda").manual_seed(17)
R = lambda lo, hi, shape: torch.randint(lo, hi, shape, device="cuda", generator=g, dtype=torch.int32)
res = []
for kind in a.kinds.split(","):
    shape = (n, 3, rows, cols)
    if kind == "mix":
        typ = R(0, 3, shape)
        pos = torch.where(typ == 1, R(0, 32, shape), R(0, 4, shape))
    elif kind == "off":
        typ, pos = torch.zeros(shape, dtype=torch.int32, device="cuda"), torch.zeros(shape, dtype=torch.int32, device="cuda")
    elif kind == "band":
        typ, pos = torch.ones(shape, dtype=torch.int32, device="cuda"), R(0, 32, shape)
    else:
        typ = torch.full(shape, 2, dtype=torch.int32, device="cuda")
        pos = torch.full(shape, int(kind[2:]), dtype=torch.int32, device="cuda")
    off = R(-7, 8, shape + (4,)) & 0xff
    params = torch.cat([typ[..., None], pos[..., None], off], dim=-1).to(torch.uint8).contiguous()
    fn = lambda: dev.sao_batch_device(frames, out, n, w, h, params, ctb, stream=st, flags=flags)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    res.append("%s %.3f" % (kind, e0.elapsed_time(e1) / a.iters))
print("%s n=%d ctb=%d depth=%d ms: %s" % (os.path.basename(os.environ.get("HDB_LIBRARY", "default")), n, ctb, a.depth, "  ".join(res)))
