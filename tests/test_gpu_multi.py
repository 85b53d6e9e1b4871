"""The multi-GPU partitioner.  With one GPU the same code runs as N logical shards on device 0
("1 GPU with N logical shards"); with several GPUs it uses them."""
import numpy as np
import pytest

import gpu_video_codec_b200 as hdb
from oracle import oracle_py as orc
from tests import common, synth

pytestmark = pytest.mark.gpu


def _device_sets():
    import torch
    n = torch.cuda.device_count()
    sets = [[0], [0, 0], [0, 0, 0, 0, 0, 0, 0, 0]]
    if n >= 2:
        sets.append(list(range(min(n, 8))))
    return sets


@pytest.mark.parametrize("devices", _device_sets(), ids=lambda d: "dev" + "".join(map(str, d)))
def test_frame_sharded_batch(devices):
    m = hdb.MultiDeblocker(devices)
    assert m.device_count() == len(devices)
    n, w, h = 21, 768, 576  # not a multiple of any device count
    frames, qps, v, hh = synth.make_batch(31, n, w, h)
    got = m.deblock_batch(frames, w, h, qps, v, hh)
    want = orc.deblock_batch(frames, w, h, qps, v, hh)
    assert np.array_equal(got, want)
    # fewer frames than devices
    got = m.deblock_batch(frames[:3], w, h, qps[:3])
    assert np.array_equal(got, orc.deblock_batch(frames[:3], w, h, qps[:3]))
    m.close()


@pytest.mark.parametrize("devices", _device_sets()[1:], ids=lambda d: "dev" + "".join(map(str, d)))
@pytest.mark.parametrize("size", [(7680, 4320), (1920, 1080), (3840, 2160), (256, 136), (64, 64)], ids=lambda s: "%dx%d" % s)
def test_strip_split_equals_unsplit(devices, size):
    """A frame cut into CTU-row strips with P2P halos == the unsplit oracle output, byte-exact."""
    w, h = size
    m = hdb.MultiDeblocker(devices)
    rng = np.random.default_rng(w + h)
    frame = synth.make_frame(rng, w, h, "blocky")
    v, hh = synth.make_bs(rng, w, h, "uniform012")
    t = hdb.Timing()
    got = m.deblock_frame_strips(frame, w, h, 37, v, hh, timing=t)
    want = orc.deblock_frame(frame, w, h, 37, v, hh)
    assert np.array_equal(got, want), int((got != want).sum())
    # default strengths, in place
    buf = frame.copy()
    m.deblock_frame_strips(buf, w, h, 33, out=buf)
    assert np.array_equal(buf, orc.deblock_frame(frame, w, h, 33))
    m.close()


def test_strips_golden_8k():
    case = [c for c in common.golden()["synthetic"] if c["w"] == 7680][0]
    frame, v, hh = common.synth_case(case["seed"], 7680, 4320, case["kind"], case["bs_kind"])
    import torch
    n = torch.cuda.device_count()
    m = hdb.MultiDeblocker(list(range(n)) if n >= 2 else [0] * 8)
    y = m.deblock_frame_strips(frame, 7680, 4320, case["qp"], v, hh)
    assert common.md5(y) == case["md5"]
    m.close()


@pytest.mark.parametrize("devices", _device_sets()[1:], ids=lambda d: "dev" + "".join(map(str, d)))
def test_frame_sharded_batch_nv12_with_edge_maps(devices):
    """The flagged batch form across devices: NV12 frames with per-edge QP maps, sharded by frame."""
    m = hdb.MultiDeblocker(devices)
    n, w, h = 13, 768, 576
    rng = np.random.default_rng(41)
    frames = np.stack([orc.yv12_to_nv12(synth.make_frame(rng, w, h, "blocky"), w, h) for _ in range(n)])
    qps = rng.integers(18, 52, n).astype(np.uint32)
    maps = [synth.make_edge_maps(rng, w, h, "smooth", picture_edges=bool(f & 1)) for f in range(n)]
    v, hh = np.stack([a for a, _ in maps]), np.stack([b for _, b in maps])
    got = m.deblock_batch_ex(frames, w, h, qps, v, hh, flags=hdb.FLAG_NV12 | hdb.FLAG_EDGE_MAPS)
    for f in range(n):
        assert np.array_equal(got[f], orc.deblock_frame_nv12(frames[f], w, h, int(qps[f]), v[f], hh[f], edges=True)), f
    m.close()


@pytest.mark.parametrize("devices", _device_sets()[1:], ids=lambda d: "dev" + "".join(map(str, d)))
@pytest.mark.parametrize("size", [(1920, 1080), (3840, 2160), (256, 136), (200, 120)], ids=lambda s: "%dx%d" % s)
def test_strip_split_with_edge_maps(devices, size):
    """The strip split with per-edge QP: the maps keep their frame-global indexing on every device."""
    w, h = size
    m = hdb.MultiDeblocker(devices)
    rng = np.random.default_rng(w * 5 + h)
    frame = synth.make_frame(rng, w, h, "blocky")
    v, hh = synth.make_edge_maps(rng, w, h, "random", picture_edges=True)
    got = m.deblock_frame_strips(frame, w, h, 37, v, hh, flags=hdb.FLAG_EDGE_MAPS)
    want = orc.deblock_frame_edges(frame, w, h, 37, v, hh)
    assert np.array_equal(got, want), int((got != want).sum())
    with pytest.raises(hdb.DeblockError):
        m.deblock_frame_strips(frame, w, h, 37, flags=hdb.FLAG_EDGE_MAPS)       # maps are required
    m.close()


@pytest.mark.parametrize("devices", _device_sets()[1:], ids=lambda d: "dev" + "".join(map(str, d)))
@pytest.mark.parametrize("size,n", [((1920, 1080), 7), ((7680, 4320), 4), ((256, 136), 11)], ids=lambda s: str(s).replace(" ", ""))
def test_strip_stream_equals_unsplit(devices, size, n):
    """A stream of frames, each cut into strips, pipelined over three slots per device (more frames than slots):
    every frame == the unsplit oracle; per-frame QP and strengths; a second call continues the slot ring."""
    w, h = size
    m = hdb.MultiDeblocker(devices)
    frames, qps, v, hh = synth.make_batch(w + n, min(n, 4), w, h)
    want = orc.deblock_batch(frames, w, h, qps, v, hh)
    idx = [i % frames.shape[0] for i in range(n)]
    t = hdb.Timing()
    got = m.deblock_stream_strips(frames[idx], w, h, qps[idx], v[idx], hh[idx], timing=t)
    for f in range(n):
        assert np.array_equal(got[f], want[idx[f]]), (f, int((got[f] != want[idx[f]]).sum()))
    assert t.kernel_launches == n * len(m.strip_layout(h)) and t.kernel_ms > 0
    # defaults, in place, continuing the ring at an odd position
    buf = frames[idx][:5].copy()
    m.deblock_stream_strips(buf, w, h, qps[idx][:5], out=buf)
    assert np.array_equal(buf, orc.deblock_batch(frames[idx][:5], w, h, qps[idx][:5]))
    m.close()


def test_strip_layout_and_halo_bytes():
    m = hdb.MultiDeblocker([0] * 8)
    lay = m.strip_layout(4320)
    assert [r for _, r in lay] == [576, 576, 576, 576, 512, 512, 512, 480]      # 68 CTU rows: 9,9,9,9,8,8,8,8 (the last one partial)
    assert [y for y, _ in lay] == [0, 576, 1152, 1728, 2304, 2816, 3328, 3840]
    assert m.strip_halo_bytes(7680, 4320) == 7 * 2 * (4 * 7680 + 2 * 4 * 3840)
    assert len(m.strip_layout(64)) == 1
    m.close()


@pytest.mark.parametrize("devices", _device_sets()[1:], ids=lambda d: "dev" + "".join(map(str, d)))
@pytest.mark.parametrize("edges", [False, True])
def test_strips_resident_in_device_memory(devices, edges):
    """Strips that already sit in each GPU's memory: one kernel per GPU, halo rows by peer loads inside it."""
    import torch
    w, h = 3840, 2160
    m = hdb.MultiDeblocker(devices)
    rng = np.random.default_rng(12 + edges)
    frame = synth.make_frame(rng, w, h, "blocky")
    v, hh = synth.make_edge_maps(rng, w, h, "random", picture_edges=True) if edges else synth.make_bs(rng, w, h, "uniform012")
    want = orc.deblock_frame_edges(frame, w, h, 36, v, hh) if edges else orc.deblock_frame(frame, w, h, 36, v, hh)
    cw = w // 2
    Y, U, V = frame[:w * h].reshape(h, w), frame[w * h:w * h * 5 // 4].reshape(h // 2, cw), frame[w * h * 5 // 4:].reshape(h // 2, cw)
    lay = m.strip_layout(h)
    ins, outs, dv, dh = [], [], [], []
    for d, (y0, rows) in enumerate(lay):
        dev = "cuda:%d" % devices[d]
        ins.append(tuple(torch.from_numpy(np.ascontiguousarray(p)).to(dev) for p in (Y[y0:y0 + rows], U[y0 // 2:(y0 + rows) // 2], V[y0 // 2:(y0 + rows) // 2])))
        outs.append(tuple(torch.zeros_like(p) for p in ins[-1]))
        dv.append(torch.from_numpy(v).to(dev))
        dh.append(torch.from_numpy(hh).to(dev))
    for d in set(devices):
        torch.cuda.synchronize(d)
    t = hdb.Timing()
    m.deblock_strips_device(ins, outs, w, h, 36, dv, dh, flags=hdb.FLAG_EDGE_MAPS if edges else 0, timing=t)
    got = np.concatenate([np.concatenate([o[p].cpu().numpy().reshape(-1) for o in outs]) for p in range(3)])
    assert np.array_equal(got, want), int((got != want).sum())
    assert t.kernel_launches == len(lay) and t.kernel_ms > 0
    with pytest.raises(hdb.DeblockError, match="out of place"):
        m.deblock_strips_device(ins, ins, w, h, 36, dv, dh, flags=hdb.FLAG_EDGE_MAPS if edges else 0)
    m.close()


@pytest.mark.parametrize("devices", _device_sets()[1:], ids=lambda d: "dev" + "".join(map(str, d)))
def test_frame_sharded_batch_10bit(devices):
    """The flagged batch form across devices with 16-bit samples: planar 10-bit and P010, sharded by frame."""
    from tests.test_depth import make_frame16
    m = hdb.MultiDeblocker(devices)
    n, w, h = 11, 352, 288
    rng = np.random.default_rng(51)
    planar = np.stack([make_frame16(rng, w, h, "blocky", 10) for _ in range(n)])
    qps = rng.integers(18, 52, n).astype(np.uint32)
    bs = [synth.make_bs(rng, w, h, "uniform012") for _ in range(n)]
    v, hh = np.stack([a for a, _ in bs]), np.stack([b for _, b in bs])
    got = m.deblock_batch_ex(planar, w, h, qps, v, hh, flags=hdb.flag_depth(10))
    for f in range(n):
        assert np.array_equal(got[f], orc.deblock_frame16(planar[f], w, h, int(qps[f]), v[f], hh[f], bit_depth=10)), f
    p010 = np.stack([orc.planar_to_p010(planar[f], w, h) for f in range(n)])
    got = m.deblock_batch_ex(p010, w, h, qps, v, hh, flags=hdb.flag_depth(10) | hdb.FLAG_NV12)
    for f in range(n):
        assert np.array_equal(got[f], orc.deblock_frame_p010(p010[f], w, h, int(qps[f]), v[f], hh[f])), f
    m.close()
