"""GPU parity of the edge-map entry points (per-edge QP; boundary-strength / QP derivation) through
the C ABI: golden hashes that came out of the unmodified reference, the restatement on many sizes,
both kernel paths, host pipeline and device-resident batches, and size-independent properties at
the full 4K size.  Bit-exact: array_equal / md5 only.  Run with  pytest -m gpu."""
import hashlib
import json
import os

import numpy as np
import pytest

import gpu_video_codec_b200 as hdb
from oracle import make_golden_edges as mge
from oracle import oracle_py as orc
from tests import common, synth

pytestmark = pytest.mark.gpu

with open(os.path.join(common.GOLDEN_DIR, "edges.json")) as _f:
    GOLDEN_EDGES = json.load(_f)["edges"]


@pytest.fixture(scope="module")
def dev():
    d = hdb.Deblocker(0)
    yield d
    d.close()


@pytest.mark.parametrize("case", GOLDEN_EDGES, ids=lambda c: "s%d-%dx%d" % (c["seed"], c["w"], c["h"]))
def test_golden_edges(dev, case):
    frame, v, hh = mge.case_inputs(case["seed"], case["w"], case["h"], case["kind"], case["edge_kind"])
    assert hashlib.md5(frame.tobytes() + v.tobytes() + hh.tobytes()).hexdigest() == case["input_md5"]
    t = hdb.Timing()
    y = dev.deblock_frame_edges(frame, case["w"], case["h"], case["qp"], v, hh, timing=t)
    assert common.md5(y) == case["md5"] and common.sha256(y) == case["sha256"]
    assert t.path == common.expected_path(case["w"])


SIZES = [(32, 8), (64, 48), (256, 64), (352, 288), (768, 576), (8192, 16), (1920, 1080),        # tiled kernel
         (8, 8), (24, 40), (72, 40), (200, 120), (1912, 1080)]                                  # generic kernels


@pytest.mark.parametrize("size", SIZES, ids=lambda s: "%dx%d" % s)
def test_edge_sizes_vs_oracle(dev, size):
    w, h = size
    rng = np.random.default_rng(w * 137 + h)
    for kind, edge_kind, pic, qp in [("blocky", "random", False, 37), ("steps", "smooth", True, 30),
                                     ("saturated", "wide", True, 51), ("random", "random", True, 44),
                                     ("blocky", "uniform", False, 22)]:
        frame = synth.make_frame(rng, w, h, kind)
        v, hh = synth.make_edge_maps(rng, w, h, edge_kind, picture_edges=pic)
        got = dev.deblock_frame_edges(frame, w, h, qp, v, hh)
        want = orc.deblock_frame_edges(frame, w, h, qp, v, hh)
        assert np.array_equal(got, want), (size, kind, edge_kind, qp, int((got != want).sum()))


def test_uniform_edge_maps_equal_plain_entry_point(dev):
    rng = np.random.default_rng(6)
    for (w, h) in [(352, 288), (200, 120)]:
        frame = synth.make_frame(rng, w, h, "blocky")
        v, hh = synth.make_bs(rng, w, h, "uniform012")
        for qp in (17, 18, 33, 51):
            a = dev.deblock_frame_edges(frame, w, h, qp, synth.edge_bytes(v, qp), synth.edge_bytes(hh, qp))
            assert np.array_equal(a, dev.deblock_frame(frame, w, h, qp, v, hh))


def test_edge_errors(dev):
    w, h = 64, 48
    frame = np.zeros(3 * w * h // 2, np.uint8)
    v, hh = synth.make_edge_maps(np.random.default_rng(0), w, h)
    with pytest.raises(hdb.DeblockError, match="boundary strenght"):
        dev.deblock_frame_edges(frame, w, h, 30, v[:-1], hh)
    with pytest.raises(hdb.DeblockError):
        dev.deblock_frame_edges(frame, w, h, 30, None, None)        # no default for edge maps
    with pytest.raises(hdb.DeblockError, match="multiplier of sample block size"):
        dev.deblock_frame_edges(np.zeros(3 * 12 * 8 // 2, np.uint8), 12, 8, 30, v, hh)
    assert np.array_equal(dev.deblock_frame_edges(frame, w, h, 30, v, hh), orc.deblock_frame_edges(frame, w, h, 30, v, hh))


def _edge_batch(seed, n, w, h, edge_kind="smooth"):
    rng = np.random.default_rng(seed)
    frames = np.stack([synth.make_frame(rng, w, h, "blocky") for _ in range(n)])
    qps = rng.integers(18, 52, n).astype(np.uint32)
    maps = [synth.make_edge_maps(rng, w, h, edge_kind, picture_edges=bool(f & 1)) for f in range(n)]
    return frames, qps, np.stack([m[0] for m in maps]), np.stack([m[1] for m in maps])


def _oracle_batch(frames, w, h, qps, v, hh):
    return np.stack([orc.deblock_frame_edges(frames[f], w, h, int(qps[f]), v[f], hh[f]) for f in range(frames.shape[0])])


@pytest.mark.parametrize("n,w,h", [(1, 352, 288), (9, 352, 288), (40, 1920, 1080), (5, 24, 40)])
def test_edge_batch_host_pipeline(dev, n, w, h):
    frames, qps, v, hh = _edge_batch(n * 100 + w, n, w, h)
    t = hdb.Timing()
    got = dev.deblock_batch_edges(frames, w, h, qps, v, hh, timing=t)
    want = _oracle_batch(frames, w, h, qps, v, hh)
    for f in range(n):
        assert np.array_equal(got[f], want[f]), f
    assert t.d2h_bytes == frames.nbytes and t.h2d_bytes >= frames.nbytes + v.nbytes + hh.nbytes


@pytest.mark.parametrize("in_place", [False, True])
def test_edge_batch_device_resident(dev, in_place):
    import torch
    n, w, h = 10, 1920, 1080
    frames, qps, v, hh = _edge_batch(9, n, w, h, "random")
    d_in = torch.from_numpy(frames).cuda()
    d_out = d_in if in_place else torch.empty_like(d_in)
    d_qp = torch.from_numpy(qps.astype(np.int32)).cuda()
    d_v, d_h = torch.from_numpy(v).cuda(), torch.from_numpy(hh).cuda()
    dev.deblock_batch_edges_device(d_in, d_out, n, w, h, d_qp, d_v, d_h, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), _oracle_batch(frames, w, h, qps, v, hh))


# ---- derivation ---------------------------------------------------------------------------------
@pytest.mark.parametrize("size", [(8, 8), (16, 24), (64, 48), (352, 288), (1920, 1080), (3840, 2160)], ids=lambda s: "%dx%d" % s)
@pytest.mark.parametrize("kind", ["intra", "random", "cu16", "extreme"])
def test_derive_edges_vs_oracle(dev, size, kind):
    w, h = size
    rng = np.random.default_rng(w * 3 + h)
    n = 3
    info = np.stack([synth.make_block_info(rng, w, h, kind) for _ in range(n)])
    v, hh = dev.derive_edges(info, w, h)
    for f in range(n):
        ov, oh = orc.derive_edges(info[f], w, h)
        assert np.array_equal(v[f], ov) and np.array_equal(hh[f], oh), (size, kind, f)


def test_derive_then_filter_on_device(dev):
    """The decoder-loop shape of the path: block data -> edge maps -> filter, all device-resident, one stream."""
    import torch
    n, w, h = 6, 1920, 1080
    rng = np.random.default_rng(77)
    frames = np.stack([synth.make_frame(rng, w, h, "blocky") for _ in range(n)])
    info = np.stack([synth.make_block_info(rng, w, h, "cu16") for _ in range(n)])
    qps = rng.integers(22, 45, n).astype(np.uint32)
    d_info = torch.from_numpy(info.view(np.uint8).reshape(n, -1)).cuda()
    d_v = torch.empty((n, hdb.num_vert_bs(w, h)), dtype=torch.uint8, device="cuda")
    d_h = torch.empty((n, hdb.num_hor_bs(w, h)), dtype=torch.uint8, device="cuda")
    d_io = torch.from_numpy(frames).cuda()
    d_qp = torch.from_numpy(qps.astype(np.int32)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    dev.derive_edges_device(d_info, n, w, h, d_v, d_h, stream=st)
    dev.deblock_batch_edges_device(d_io, d_io, n, w, h, d_qp, d_v, d_h, stream=st)
    torch.cuda.synchronize()
    got = d_io.cpu().numpy()
    for f in range(n):
        ov, oh = orc.derive_edges(info[f], w, h)
        assert np.array_equal(d_v[f].cpu().numpy(), ov) and np.array_equal(d_h[f].cpu().numpy(), oh)
        assert np.array_equal(got[f], orc.deblock_frame_edges(frames[f], w, h, int(qps[f]), ov, oh)), f
    assert (got != frames).mean() > 0.02


def test_full_size_edge_batch_properties(dev):
    """128 x 4K resident in HBM: replicas of 4 distinct inputs must give replicas of the oracle's 4, and
    maps carrying one QP must reproduce the plain kernel's output byte for byte."""
    import torch
    n, w, h, distinct = 128, 3840, 2160, 4
    frames, qps, v, hh = _edge_batch(5, distinct, w, h)
    reps = n // distinct
    d_in = torch.from_numpy(frames).cuda().repeat(reps, 1)
    d_qp = torch.from_numpy(qps.astype(np.int32)).cuda().repeat(reps)
    d_v, d_h = torch.from_numpy(v).cuda().repeat(reps, 1), torch.from_numpy(hh).cuda().repeat(reps, 1)
    d_out = torch.empty_like(d_in)
    st = torch.cuda.current_stream().cuda_stream
    dev.deblock_batch_edges_device(d_in, d_out, n, w, h, d_qp, d_v, d_h, stream=st)
    torch.cuda.synchronize()
    want = torch.from_numpy(_oracle_batch(frames, w, h, qps, v, hh)).cuda()
    assert bool((d_out.view(reps, distinct, -1) == want.unsqueeze(0)).all())
    # one QP per frame in every entry == the plain entry point with the strengths alone
    bs_v, bs_h = d_v & 3, d_h & 3
    q8 = d_qp.to(torch.uint8).clamp(max=51).view(n, 1)
    dev.deblock_batch_edges_device(d_in, d_out, n, w, h, d_qp, bs_v | (q8 << 2), bs_h | (q8 << 2), stream=st)
    d_ref = torch.empty_like(d_in)
    dev.deblock_batch_device(d_in, d_ref, n, w, h, d_qp, bs_v, bs_h, stream=st)
    torch.cuda.synchronize()
    assert bool((d_out == d_ref).all())


@pytest.mark.parametrize("n,w,h,nv12", [(1, 352, 288, False), (7, 352, 288, True), (40, 1920, 1080, False), (5, 24, 40, False)])
def test_batch_from_block_data(dev, n, w, h, nv12):
    """hdb_deblock_batch_blocks: frames + block data from host memory, edge maps derived on the device slice by
    slice (one slice, several slices, both kernel paths, NV12) == oracle derivation + oracle per-edge-QP filter."""
    rng = np.random.default_rng(n * 31 + w)
    frames = np.stack([synth.make_frame(rng, w, h, "blocky") for _ in range(n)])
    info = np.stack([synth.make_block_info(rng, w, h, ("cu16", "random", "intra")[f % 3]) for f in range(n)])
    qps = rng.integers(18, 52, n).astype(np.uint32)
    stored = np.stack([orc.yv12_to_nv12(f, w, h) for f in frames]) if nv12 else frames
    t = hdb.Timing()
    got = dev.deblock_batch_blocks(stored, w, h, qps, info, flags=hdb.FLAG_NV12 if nv12 else 0, timing=t)
    for f in range(n):
        ov, oh = orc.derive_edges(info[f], w, h)
        want = orc.deblock_frame_edges(frames[f], w, h, int(qps[f]), ov, oh)
        assert np.array_equal(orc.nv12_to_yv12(got[f], w, h) if nv12 else got[f], want), f
    assert t.h2d_bytes >= frames.nbytes + info.nbytes and t.d2h_bytes == frames.nbytes
    with pytest.raises(hdb.DeblockError):
        dev.deblock_batch_blocks(stored, w, h, qps, None)


@pytest.mark.parametrize("n,w,h", [(1, 352, 288), (4, 768, 576)])
def test_batch_from_block_data_on_mapped_pinned_buffers(dev, n, w, h):
    """The single-slice form of hdb_deblock_batch_blocks on page-locked mapped buffers (the kernel reads and writes
    host memory itself; block data and derivation still go through the device) and with per-edge maps likewise."""
    import torch
    rng = np.random.default_rng(n + w)
    frames = np.stack([synth.make_frame(rng, w, h, "blocky") for _ in range(n)])
    info = np.stack([synth.make_block_info(rng, w, h, "cu16") for _ in range(n)])
    qps = rng.integers(18, 52, n).astype(np.uint32)
    t_in = torch.from_numpy(frames).pin_memory()
    t_out = torch.zeros_like(t_in).pin_memory()
    t_qp = torch.from_numpy(qps.astype(np.int32)).pin_memory()
    tm = hdb.Timing()
    dev.deblock_batch_blocks(t_in, w, h, t_qp, info, out=t_out, timing=tm, n_frames=n)
    maps = [orc.derive_edges(info[f], w, h) for f in range(n)]
    want = np.stack([orc.deblock_frame_edges(frames[f], w, h, int(qps[f]), *maps[f]) for f in range(n)])
    assert np.array_equal(t_out.numpy(), want)
    if n == 1:      # one slice: derive + filter, no copy-engine leg for the frames (several frames are cut into ramped slices)
        assert tm.copy_ms == 0.0 and tm.kernel_launches == 2
    assert np.array_equal(t_in.numpy(), frames)
    v, hh = np.stack([m[0] for m in maps]), np.stack([m[1] for m in maps])
    t_out.zero_()
    dev.deblock_batch_edges(t_in, w, h, t_qp, v, hh, out=t_out, n_frames=n)
    assert np.array_equal(t_out.numpy(), want)
