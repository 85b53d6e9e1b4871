"""GPU parity for samples deeper than 8 bits (HDB_FLAG_DEPTH): DeblockTiled16Kernel through the C ABI against
oracle/deblock_oracle.h::hdbo_deblock_frame16.  Depth 8 in 16-bit containers is pinned to the reference (it must
equal the 8-bit entry point and the reference's golden hashes); 9 and 10 are unpinned against the reference and
compared with the oracle's definition only."""
import numpy as np
import pytest

import gpu_video_codec_b200 as hdb
from oracle import oracle_py as orc
from tests import common, synth
from tests.test_depth import make_frame16

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    d = hdb.Deblocker(0)
    yield d
    d.close()


def run16(dev, frames, w, h, qps, v=None, hh=None, depth=10, edges=False, timing=None):
    flags = hdb.flag_depth(depth) | (hdb.FLAG_EDGE_MAPS if edges else 0)
    return dev.deblock_batch_ex(frames, w, h, np.asarray(qps, np.uint32), v, hh, flags=flags, timing=timing)


@pytest.mark.parametrize("case", [c for c in common.golden()["bundled"]], ids=lambda c: "%s-qp%d" % (c["file"][:8], c["qp"]))
def test_reference_frames_in_16bit_containers(dev, case):
    """Depth 8 in 16-bit containers: the reference's own frames must give the reference's golden hashes."""
    x = common.fixture(case["file"])
    w, h = case["w"], case["h"]
    if w % 16:
        pytest.skip("16-bit frames need a width that is a multiple of 16")
    y = run16(dev, x.astype(np.uint16)[None], w, h, [case["qp"]], depth=8)[0]
    assert int(y.max()) <= 255 and common.md5(y.astype(np.uint8)) == case["md5"]


SIZES = [(32, 8), (48, 24), (64, 48), (256, 64), (272, 72), (352, 288), (528, 16), (768, 576), (4096, 16), (1920, 1080), (2000, 1048)]


@pytest.mark.parametrize("depth", [8, 9, 10])
@pytest.mark.parametrize("size", SIZES, ids=lambda s: "%dx%d" % s)
def test_sizes_vs_oracle16(dev, size, depth):
    w, h = size
    rng = np.random.default_rng(w * 17 + h + depth)
    for kind, bs_kind, qp in [("blocky", "uniform012", 37), ("saturated", "wide", 51), ("random", "twos", 44), ("steps", None, 27)]:
        frame = make_frame16(rng, w, h, kind, depth)
        v, hh = (None, None) if bs_kind is None else synth.make_bs(rng, w, h, bs_kind)
        t = hdb.Timing()
        got = run16(dev, frame[None], w, h, [qp], None if v is None else v[None], None if hh is None else hh[None], depth, timing=t)[0]
        want = orc.deblock_frame16(frame, w, h, qp, v, hh, bit_depth=depth)
        assert t.path == 1
        assert np.array_equal(got, want), (size, depth, kind, bs_kind, qp, int((got != want).sum()))


@pytest.mark.parametrize("depth", [8, 10])
def test_edge_maps_16bit(dev, depth):
    rng = np.random.default_rng(depth)
    for (w, h) in [(352, 288), (1920, 1080)]:
        frame = make_frame16(rng, w, h, "blocky", depth)
        v, hh = synth.make_edge_maps(rng, w, h, "random", picture_edges=True)
        got = run16(dev, frame[None], w, h, [33], v[None], hh[None], depth, edges=True)[0]
        assert np.array_equal(got, orc.deblock_frame16(frame, w, h, 33, v, hh, bit_depth=depth, edges=True))


def test_depth8_containers_equal_the_8bit_entry_point(dev):
    rng = np.random.default_rng(5)
    w, h = 768, 576
    frames, qps, v, hh = synth.make_batch(3, 5, w, h)
    a = dev.deblock_batch(frames, w, h, qps, v, hh)
    b = run16(dev, frames.astype(np.uint16), w, h, qps, v, hh, depth=8)
    assert np.array_equal(b, a.astype(np.uint16))


def test_batch16_host_pipeline_and_device_resident(dev):
    """More slices than slots through the pinned pipeline (a 10-bit 1080p frame is 6.2 MB), and the same batch
    resident in HBM between guard bands, out of place and in place."""
    import torch
    rng = np.random.default_rng(11)
    n, w, h, depth = 6, 1920, 1080, 10
    frames = np.stack([make_frame16(rng, w, h, "blocky", depth) for _ in range(n)])
    qps = rng.integers(18, 52, n).astype(np.uint32)
    bs = [synth.make_bs(rng, w, h, "uniform012") for _ in range(n)]
    v, hh = np.stack([a for a, _ in bs]), np.stack([b for _, b in bs])
    want = np.stack([orc.deblock_frame16(frames[f], w, h, int(qps[f]), v[f], hh[f], bit_depth=depth) for f in range(n)])
    reps = 10
    tf = torch.from_numpy(np.tile(frames.view(np.int16), (reps, 1))).pin_memory()
    tq = torch.from_numpy(np.tile(qps, reps).astype(np.int32)).pin_memory()
    tv, th = torch.from_numpy(np.tile(v, (reps, 1))).pin_memory(), torch.from_numpy(np.tile(hh, (reps, 1))).pin_memory()
    out = torch.empty_like(tf).pin_memory()
    t = hdb.Timing()
    dev.deblock_batch_ex(tf, w, h, tq, tv, th, flags=hdb.flag_depth(depth), out=out, n_frames=n * reps, timing=t)
    got = out.numpy().view(np.uint16)
    for f in range(n * reps):
        assert np.array_equal(got[f], want[f % n]), f
    assert t.d2h_bytes == n * reps * frames.shape[1] * 2
    # device-resident
    fb, guard = frames.shape[1] * 2, 1 << 16
    buf = torch.full((guard + n * fb + guard,), 0xA5, dtype=torch.uint8, device="cuda")
    d_in = torch.from_numpy(frames.view(np.int16)).cuda()
    d_qp, d_v, d_h = torch.from_numpy(qps.astype(np.int32)).cuda(), torch.from_numpy(v).cuda(), torch.from_numpy(hh).cuda()
    st = torch.cuda.current_stream().cuda_stream
    o = buf[guard:guard + n * fb]
    dev.deblock_batch_device_ex(d_in, o, n, w, h, d_qp, d_v, d_h, flags=hdb.flag_depth(depth), stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(o.cpu().numpy().view(np.uint16).reshape(n, -1), want)
    assert bool((buf[:guard] == 0xA5).all()) and bool((buf[guard + n * fb:] == 0xA5).all())
    o.copy_(d_in.view(torch.uint8).reshape(-1))
    dev.deblock_batch_device_ex(o, o, n, w, h, d_qp, d_v, d_h, flags=hdb.flag_depth(depth), stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(o.cpu().numpy().view(np.uint16).reshape(n, -1), want)


@pytest.mark.parametrize("size", [(32, 8), (64, 48), (272, 72), (352, 288), (528, 16), (1920, 1080), (2000, 1048)], ids=lambda s: "%dx%d" % s)
def test_p010_sizes_vs_oracle(dev, size):
    """P010: Y plane + interleaved UV, 10-bit samples in the high bits of 16-bit containers (HDB_FLAG_NV12 | HDB_FLAG_DEPTH(10))."""
    w, h = size
    rng = np.random.default_rng(w * 19 + h)
    flags = hdb.FLAG_NV12 | hdb.flag_depth(10)
    for kind, bs_kind, qp in [("blocky", "uniform012", 37), ("saturated", "wide", 51), ("random", None, 44)]:
        frame = orc.planar_to_p010(make_frame16(rng, w, h, kind, 10), w, h)
        v, hh = (None, None) if bs_kind is None else synth.make_bs(rng, w, h, bs_kind)
        got = dev.deblock_batch_ex(frame[None], w, h, np.array([qp], np.uint32), None if v is None else v[None],
                                   None if hh is None else hh[None], flags=flags)[0]
        want = orc.deblock_frame_p010(frame, w, h, qp, v, hh)
        assert np.array_equal(got, want), (size, kind, bs_kind, qp, int((got != want).sum()))
    # edge maps, and a batch on the device in place
    import torch
    n = 3
    frames = np.stack([orc.planar_to_p010(make_frame16(rng, w, h, "blocky", 10), w, h) for _ in range(n)])
    maps = [synth.make_edge_maps(rng, w, h, "random", picture_edges=True) for _ in range(n)]
    ev, eh = np.stack([m[0] for m in maps]), np.stack([m[1] for m in maps])
    qps = rng.integers(18, 52, n).astype(np.uint32)
    d = torch.from_numpy(frames.view(np.int16)).cuda()
    dev.deblock_batch_device_ex(d, d, n, w, h, torch.from_numpy(qps.astype(np.int32)).cuda(), torch.from_numpy(ev).cuda(),
                                torch.from_numpy(eh).cuda(), flags=flags | hdb.FLAG_EDGE_MAPS, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d.cpu().numpy().view(np.uint16)
    for f in range(n):
        assert np.array_equal(got[f], orc.deblock_frame_p010(frames[f], w, h, int(qps[f]), ev[f], eh[f], edges=True)), f


def test_depth_errors(dev):
    w, h = 64, 48
    f = np.zeros((1, 3 * w * h // 2), np.uint16)
    q = np.array([30], np.uint32)
    for bad in (hdb.flag_depth(7), hdb.flag_depth(11), hdb.flag_depth(12), hdb.flag_depth(8) | hdb.FLAG_NV12):
        with pytest.raises(hdb.DeblockError):
            dev.deblock_batch_ex(f, w, h, q, flags=bad)
    with pytest.raises(hdb.DeblockError, match="multiple of 16"):
        dev.deblock_batch_ex(np.zeros((1, 3 * 40 * 24 // 2), np.uint16), 40, 24, q, flags=hdb.flag_depth(10))
