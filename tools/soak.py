"""Development aid: a longer randomised parity soak than the test suite affords (run on the GPU box).
Random geometry, QP, content, boundary strengths; single frames, host batches, device batches (in place and
out of place), pitched surfaces and logical-shard strips, all against the oracle; and the widened entry
points: edge maps (per-edge QP) on host and device batches, NV12 frames with and without edge maps,
pitched NV12 surfaces, boundary-strength derivation, 16-bit samples, sample adaptive offset at 8 .. 12 bits."""
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
only = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else list(range(15))     # modes to draw from
rng = np.random.default_rng(seed)
dev = hdb.Deblocker(0)
multi = hdb.MultiDeblocker([0, 0, 0])
kinds = ["blocky", "random", "steps", "saturated", "ramp", "flat"]
bs_kinds = [None, "uniform012", "wide", "twos", "zeros", "ones"]
t0, n_cases, fails = time.time(), 0, 0
st = torch.cuda.current_stream().cuda_stream
while time.time() - t0 < budget:
    mode = only[int(rng.integers(len(only)))]
    w = 8 * int(rng.integers(1, 200))
    h = 8 * int(rng.integers(1, 80))
    if mode in (3, 4, 6, 7, 9, 11) or rng.integers(2):
        w = 32 * max(1, w // 32)
    if mode == 12:
        w = max(32, 16 * (w // 16))
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
    if 5 <= mode <= 10:
        # edge maps (bs | qp << 2) for modes 5, 6 (odd cases of 6), 7; derivation for mode 8
        edge_kind = ["random", "wide", "smooth", "uniform"][int(rng.integers(4))]
        maps = [synth.make_edge_maps(rng, w, h, edge_kind, picture_edges=bool(rng.integers(2))) for _ in range(n)]
        ev, eh = np.stack([m[0] for m in maps]), np.stack([m[1] for m in maps])
    try:
        if mode == 5:      # edge maps, YV12: host batch or device batch, both kernel paths
            want = np.stack([orc.deblock_frame_edges(frames[i], w, h, int(qps[i]), ev[i], eh[i]) for i in range(n)])
            if rng.integers(2):
                got = dev.deblock_batch_edges(frames, w, h, qps, ev, eh)
            else:
                d = torch.from_numpy(frames).cuda()
                dev.deblock_batch_edges_device(d, d, n, w, h, torch.from_numpy(qps.astype(np.int32)).cuda(),
                                               torch.from_numpy(ev).cuda(), torch.from_numpy(eh).cuda(), stream=st)
                torch.cuda.synchronize()
                got = d.cpu().numpy()
        elif mode == 6:    # NV12 frames, with strengths / defaults / edge maps, host or device
            nv = np.stack([orc.yv12_to_nv12(f, w, h) for f in frames])
            edges = bool(rng.integers(2))
            a, b = (ev, eh) if edges else (v, hh)
            want = np.stack([orc.deblock_frame_nv12(nv[i], w, h, int(qps[i]), None if a is None else a[i],
                                                    None if b is None else b[i], edges=edges) for i in range(n)])
            fl = hdb.FLAG_NV12 | (hdb.FLAG_EDGE_MAPS if edges else 0)
            if rng.integers(2):
                got = dev.deblock_batch_ex(nv, w, h, qps, a, b, flags=fl)
            else:
                d = torch.from_numpy(nv).cuda()
                o = d if rng.integers(2) else torch.empty_like(d)
                dev.deblock_batch_device_ex(d, o, n, w, h, torch.from_numpy(qps.astype(np.int32)).cuda(),
                                            None if a is None else torch.from_numpy(a).cuda(),
                                            None if b is None else torch.from_numpy(b).cuda(), flags=fl, stream=st)
                torch.cuda.synchronize()
                got = o.cpu().numpy()
        elif mode == 7:    # pitched NV12 surfaces, in place
            nv = np.stack([orc.yv12_to_nv12(f, w, h) for f in frames])
            want = np.stack([orc.deblock_frame_nv12(nv[i], w, h, int(qps[i]), ev[i], eh[i], edges=True) for i in range(n)])
            ch = h // 2
            py, puv = w + 16 * int(rng.integers(0, 5)), w + 16 * int(rng.integers(0, 5))
            t = torch.from_numpy(nv).cuda()
            P = [torch.zeros((n, h, py), dtype=torch.uint8, device="cuda"), torch.zeros((n, ch, puv), dtype=torch.uint8, device="cuda")]
            P[0][:, :, :w] = t[:, :w * h].view(n, h, w)
            P[1][:, :, :w] = t[:, w * h:].view(n, ch, w)
            s = hdb.Nv12Surfaces(P[0].data_ptr(), P[1].data_ptr(), py, puv, h * py, ch * puv)
            dev.deblock_nv12_surfaces_device(s, s, n, w, h, torch.from_numpy(qps.astype(np.int32)).cuda(),
                                             torch.from_numpy(ev).cuda(), torch.from_numpy(eh).cuda(),
                                             flags=hdb.FLAG_EDGE_MAPS, stream=st)
            torch.cuda.synchronize()
            got = torch.cat([P[0][:, :, :w].reshape(n, -1), P[1][:, :, :w].reshape(n, -1)], 1).cpu().numpy()
        elif mode == 8:    # derivation: block data -> edge maps, then the filter with them
            ik = ["intra", "random", "cu16", "extreme"][int(rng.integers(4))]
            info = np.stack([synth.make_block_info(rng, w, h, ik) for _ in range(n)])
            gv, gh = dev.derive_edges(info, w, h)
            ov, oh = zip(*[orc.derive_edges(info[i], w, h) for i in range(n)])
            want = np.concatenate([np.stack(ov), np.stack(oh)], 1)
            got = np.concatenate([gv, gh], 1)
        elif mode == 9:    # the strip split with edge maps (logical shards)
            want = np.stack([orc.deblock_frame_edges(frames[i], w, h, int(qps[i]), ev[i], eh[i]) for i in range(n)])
            got = np.stack([multi.deblock_frame_strips(frames[i], w, h, int(qps[i]), ev[i], eh[i], flags=hdb.FLAG_EDGE_MAPS) for i in range(n)])
        elif mode == 10:   # host frames + host block data, maps derived on the device slice by slice
            ik = ["intra", "random", "cu16", "extreme"][int(rng.integers(4))]
            info = np.stack([synth.make_block_info(rng, w, h, ik) for _ in range(n)])
            want = np.stack([orc.deblock_frame_edges(frames[i], w, h, int(qps[i]), *orc.derive_edges(info[i], w, h)) for i in range(n)])
            got = dev.deblock_batch_blocks(frames, w, h, qps, info)
        elif mode == 11:   # a stream of frames, each cut into strips (logical shards), pipelined over three slots
            want = orc.deblock_batch(frames, w, h, qps, v, hh)
            got = multi.deblock_stream_strips(frames, w, h, qps, v, hh)
        elif mode == 12:   # 16-bit containers, depth 8 .. 10, host or device batch
            depth = int(rng.integers(8, 11))
            f16 = (frames.astype(np.uint16) << (depth - 8)) | rng.integers(0, 1 << (depth - 8), frames.shape).astype(np.uint16)
            want = np.stack([orc.deblock_frame16(f16[i], w, h, int(qps[i]), None if v is None else v[i], None if hh is None else hh[i], bit_depth=depth) for i in range(n)])
            if rng.integers(2):
                got = dev.deblock_batch_ex(f16, w, h, qps, v, hh, flags=hdb.flag_depth(depth))
            else:
                d = torch.from_numpy(f16.view(np.int16)).cuda()
                o = d if rng.integers(2) else torch.empty_like(d)
                dev.deblock_batch_device_ex(d, o, n, w, h, torch.from_numpy(qps.astype(np.int32)).cuda(),
                                            None if v is None else torch.from_numpy(v).cuda(), None if hh is None else torch.from_numpy(hh).cuda(),
                                            flags=hdb.flag_depth(depth), stream=st)
                torch.cuda.synchronize()
                got = o.cpu().numpy().view(np.uint16)
        elif mode == 13:   # sample adaptive offset, random parameters per coding tree block
            from tests.test_sao import make_sao_params
            ctb = int(rng.choice([16, 32, 64]))
            pk = ["mixed", "edge", "band", "extreme"][int(rng.integers(4))]
            params = np.stack([make_sao_params(rng, w, h, ctb, pk) for _ in range(n)])
            want = np.stack([orc.sao_frame(frames[i], w, h, params[i], ctb) for i in range(n)])
            got = dev.sao_batch(frames, w, h, params, ctb)
        elif mode == 14:   # the same on samples of 8 .. 12 bits in 16-bit containers
            from tests.test_sao import make_sao_params16
            from tests.test_depth import make_frame16
            ctb = int(rng.choice([16, 32, 64]))
            depth = int(rng.choice([8, 9, 10, 12]))
            pk = ["mixed", "edge", "band", "extreme"][int(rng.integers(4))]
            f16 = np.stack([make_frame16(rng, w, h, kind, depth) for _ in range(n)])
            params = np.stack([make_sao_params16(rng, w, h, ctb, pk, depth) for _ in range(n)])
            want = np.stack([orc.sao_frame16(f16[i], w, h, params[i], ctb, depth) for i in range(n)])
            got = dev.sao_batch(f16, w, h, params, ctb, flags=hdb.flag_depth(depth))
        if mode >= 5:
            pass
        elif mode == 0:
            want = orc.deblock_batch(frames, w, h, qps, v, hh)
            got = np.stack([dev.deblock_frame(frames[i], w, h, int(qps[i]), None if v is None else v[i], None if hh is None else hh[i]) for i in range(n)])
        elif mode == 1:
            want = orc.deblock_batch(frames, w, h, qps, v, hh)
            got = dev.deblock_batch(frames, w, h, qps, v, hh)
        elif mode == 2:
            want = orc.deblock_batch(frames, w, h, qps, v, hh)
            d = torch.from_numpy(frames).cuda()
            o = d if rng.integers(2) else torch.empty_like(d)
            dq = torch.from_numpy(qps.astype(np.int32)).cuda()
            dv = None if v is None else torch.from_numpy(v).cuda()
            dh = None if hh is None else torch.from_numpy(hh).cuda()
            dev.deblock_batch_device(d, o, n, w, h, dq, dv, dh, stream=st)
            torch.cuda.synchronize()
            got = o.cpu().numpy()
        elif mode == 3:
            want = orc.deblock_batch(frames, w, h, qps, v, hh)
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
            want = orc.deblock_batch(frames, w, h, qps, v, hh)
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
