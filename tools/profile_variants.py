"""Development aid: one warm launch of every kernel of the library on 4K frames (1080x1920 for the cp.async
variant), for an ncu capture (profiles/README.md has the command).  Order of the launches ncu sees (each once warm inside the profiled range, after a cold launch outside it): DeriveEdgesKernel, DeblockTiledKernel<0,0,0> (plain), <1,0,0> (edge maps), <0,1,0> (NV12),
<1,1,0> (NV12 + edge maps), DeblockTiled16Kernel<10,0> (10-bit), SaoKernel, DeblockTiledKernel<0,0,2>
(cp.async loader, 1080x1920, skewed chroma), DeblockTiledKernel<0,0,1> (strip with peer halo rows: two logical
strips on this one device)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import gpu_video_codec_b200 as hdb  # noqa: E402


def twice(fn):
    """One cold launch outside the profiled range, one warm launch inside it (ncu --profile-from-start off)."""
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


n, w, h = int(os.environ.get("N", "256")), 3840, 2160
dev = hdb.Deblocker(0)
st = torch.cuda.current_stream().cuda_stream
frames, qp, vbs, hbs = bench.workload_frames(torch, 0, n, w, h, 1, "cuda")
info = bench.device_block_info(torch, n, w, h, seed=2)
d_v = torch.empty((n, hdb.num_vert_bs(w, h)), dtype=torch.uint8, device="cuda")
d_h = torch.empty((n, hdb.num_hor_bs(w, h)), dtype=torch.uint8, device="cuda")
out = torch.empty_like(frames)
ysz, csz = w * h, (w // 2) * (h // 2)
nv = frames.clone()
nv[:, ysz:] = torch.stack([frames[:, ysz:ysz + csz], frames[:, ysz + csz:]], dim=2).reshape(n, -1)
torch.cuda.synchronize()
twice(lambda: dev.derive_edges_device(info, n, w, h, d_v, d_h, stream=st))
twice(lambda: dev.deblock_batch_device(frames, out, n, w, h, qp, d_v & 3, d_h & 3, stream=st))
twice(lambda: dev.deblock_batch_edges_device(frames, out, n, w, h, qp, d_v, d_h, stream=st))
twice(lambda: dev.deblock_batch_device_ex(nv, out, n, w, h, qp, d_v & 3, d_h & 3, flags=hdb.FLAG_NV12, stream=st))
twice(lambda: dev.deblock_batch_device_ex(nv, out, n, w, h, qp, d_v, d_h, flags=hdb.FLAG_NV12 | hdb.FLAG_EDGE_MAPS, stream=st))
torch.cuda.synchronize()
del nv, info
# 10-bit: half as many frames, the same bytes
n16 = n // 2
f16 = (frames[:n16].to(torch.int16) << 2) | 1
o16 = torch.empty_like(f16)
twice(lambda: dev.deblock_batch_device_ex(f16, o16, n16, w, h, qp[:n16], vbs[:n16], hbs[:n16], flags=hdb.flag_depth(10), stream=st))
torch.cuda.synchronize()
del f16, o16
# sample adaptive offset on the deblocked frames
rows, cols = (h + 63) // 64, (w + 63) // 64
g = torch.Generator(device="cuda").manual_seed(17)
typ = torch.randint(0, 3, (n, 3, rows, cols), device="cuda", generator=g, dtype=torch.int32)
pos = torch.where(typ == 1, torch.randint(0, 32, typ.shape, device="cuda", generator=g, dtype=torch.int32),
                  torch.randint(0, 4, typ.shape, device="cuda", generator=g, dtype=torch.int32))
off = torch.randint(-7, 8, typ.shape + (4,), device="cuda", generator=g, dtype=torch.int32) & 0xff
params = torch.cat([typ[..., None], pos[..., None], off], dim=-1).to(torch.uint8).contiguous()
sao_out = torch.empty_like(out)
twice(lambda: dev.sao_batch_device(out, sao_out, n, w, h, params, 64, stream=st))
torch.cuda.synchronize()
del sao_out, params
# the cp.async loader: portrait 1080x1920 (chroma width 540 = 4 mod 8: skewed planes)
pw, ph = 1080, 1920
pf, pq, pv, phb = bench.workload_frames(torch, 0, n, pw, ph, 2, "cuda")
po = torch.empty_like(pf)
twice(lambda: dev.deblock_batch_device(pf, po, n, pw, ph, pq, pv, phb, stream=st))
torch.cuda.synchronize()
del pf, po
# strips with peer halo rows: two logical strips of 4K frames on this device, device-resident
m = hdb.MultiDeblocker([0, 0])
lay = m.strip_layout(h)
cw = w // 2
fr = frames[0]
Y, U, V = fr[:ysz].view(h, w), fr[ysz:ysz + csz].view(h // 2, cw), fr[ysz + csz:].view(h // 2, cw)
ins = [tuple(p.contiguous() for p in (Y[y0:y0 + r], U[y0 // 2:(y0 + r) // 2], V[y0 // 2:(y0 + r) // 2])) for (y0, r) in lay]
outs = [tuple(torch.empty_like(p) for p in s) for s in ins]
tm = hdb.Timing()
twice(lambda: m.deblock_strips_device(ins, outs, w, h, int(qp[0]), [vbs[0]] * len(lay), [hbs[0]] * len(lay), timing=tm))
torch.cuda.synchronize()
print("done")
