"""Development aid: one warm launch of every kernel variant on the same 4K frames, for an ncu capture
(profiles/README.md has the command).  Order of the launches ncu sees (each twice: cold, warm):
DeriveEdgesKernel, DeblockTiledKernel<0,0> (plain), <1,0> (edge maps), <0,1> (NV12), <1,1> (NV12 + edge maps)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import gpu_video_codec_b200 as hdb  # noqa: E402

n, w, h = int(os.environ.get("N", "256")), 3840, 2160
dev = hdb.Deblocker(0)
st = torch.cuda.current_stream().cuda_stream
frames, qp, vbs, hbs = bench.device_batch(torch, n, w, h, seed=1)
info = bench.device_block_info(torch, n, w, h, seed=2)
d_v = torch.empty((n, hdb.num_vert_bs(w, h)), dtype=torch.uint8, device="cuda")
d_h = torch.empty((n, hdb.num_hor_bs(w, h)), dtype=torch.uint8, device="cuda")
out = torch.empty_like(frames)
ysz, csz = w * h, (w // 2) * (h // 2)
nv = frames.clone()
nv[:, ysz:] = torch.stack([frames[:, ysz:ysz + csz], frames[:, ysz + csz:]], dim=2).reshape(n, -1)
torch.cuda.synchronize()
for _ in range(2):
    dev.derive_edges_device(info, n, w, h, d_v, d_h, stream=st)
for _ in range(2):
    dev.deblock_batch_device(frames, out, n, w, h, qp, d_v & 3, d_h & 3, stream=st)
for _ in range(2):
    dev.deblock_batch_edges_device(frames, out, n, w, h, qp, d_v, d_h, stream=st)
for _ in range(2):
    dev.deblock_batch_device_ex(nv, out, n, w, h, qp, d_v & 3, d_h & 3, flags=hdb.FLAG_NV12, stream=st)
for _ in range(2):
    dev.deblock_batch_device_ex(nv, out, n, w, h, qp, d_v, d_h, flags=hdb.FLAG_NV12 | hdb.FLAG_EDGE_MAPS, stream=st)
torch.cuda.synchronize()
print("done")
