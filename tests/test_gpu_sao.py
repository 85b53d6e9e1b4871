"""GPU parity of sample adaptive offset (SaoKernel) through the C ABI against oracle/deblock_oracle.c::hdbo_sao_frame
(H.265 8.7.3 restated; the reference has no SAO, so the oracle is unpinned by nature -- tests/test_sao.py holds it
against an independent second statement)."""
import numpy as np
import pytest

import gpu_video_codec_b200 as hdb
from oracle import oracle_py as orc
from tests import synth
from tests.test_sao import make_sao_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    d = hdb.Deblocker(0)
    yield d
    d.close()


SIZES = [(8, 8), (16, 16), (64, 64), (72, 40), (200, 120), (352, 288), (1000, 64), (768, 576), (1920, 1080), (4096, 16)]


@pytest.mark.parametrize("ctb", [16, 32, 64])
@pytest.mark.parametrize("size", SIZES, ids=lambda s: "%dx%d" % s)
def test_sao_sizes_vs_oracle(dev, size, ctb):
    w, h = size
    rng = np.random.default_rng(w * 7 + h + ctb)
    for kind, content in [("mixed", "blocky"), ("edge", "random"), ("band", "ramp"), ("extreme", "saturated"), ("off", "steps")]:
        frame = synth.make_frame(rng, w, h, content)
        params = make_sao_params(rng, w, h, ctb, kind)
        got = dev.sao_batch(frame[None], w, h, params[None], ctb)[0]
        want = orc.sao_frame(frame, w, h, params, ctb)
        assert np.array_equal(got, want), (size, ctb, kind, content, int((got != want).sum()))


@pytest.mark.parametrize("cls", [0, 1, 2, 3])
def test_each_edge_class_everywhere(dev, cls):
    """One class in every block of every plane: the picture-border rule of that class on all four sides."""
    w, h, ctb = 200, 120, 32
    rng = np.random.default_rng(cls)
    frame = synth.make_frame(rng, w, h, "random")
    params = np.zeros((3, (h + ctb - 1) // ctb, (w + ctb - 1) // ctb), orc.SAO_PARAMS)
    params["type"] = 2
    params["position"] = cls
    params["offset"] = (5, 3, -3, -5)
    got = dev.sao_batch(frame[None], w, h, params[None], ctb)[0]
    assert np.array_equal(got, orc.sao_frame(frame, w, h, params, ctb))


def test_deblock_then_sao_on_the_device(dev):
    """The in-loop pair chained on one stream == oracle deblocking followed by oracle SAO, frame for frame; the
    stand-alone SAO entry point between guard bands; errors."""
    import torch
    n, w, h, ctb = 5, 1920, 1080, 64
    frames, qps, v, hh = synth.make_batch(9, n, w, h)
    rng = np.random.default_rng(3)
    params = np.stack([make_sao_params(rng, w, h, ctb, "mixed") for _ in range(n)])
    d_in = torch.from_numpy(frames).cuda()
    d_qp = torch.from_numpy(qps.astype(np.int32)).cuda()
    d_v, d_h = torch.from_numpy(v).cuda(), torch.from_numpy(hh).cuda()
    d_p = torch.from_numpy(params.view(np.uint8).reshape(n, -1)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    fb, guard = frames.shape[1], 1 << 16
    buf = torch.full((guard + n * fb + guard,), 0xA5, dtype=torch.uint8, device="cuda")
    out = buf[guard:guard + n * fb].view(n, fb)
    dev.deblock_sao_batch_device(d_in, out, n, w, h, d_qp, d_v, d_h, d_p, ctb, stream=st)
    torch.cuda.synchronize()
    deb = orc.deblock_batch(frames, w, h, qps, v, hh)
    want = np.stack([orc.sao_frame(deb[f], w, h, params[f], ctb) for f in range(n)])
    assert np.array_equal(out.cpu().numpy(), want)
    assert bool((buf[:guard] == 0xA5).all()) and bool((buf[guard + n * fb:] == 0xA5).all())
    assert np.array_equal(d_in.cpu().numpy(), frames)                               # input untouched
    out.fill_(0)
    dev.sao_batch_device(torch.from_numpy(deb).cuda(), out, n, w, h, d_p, ctb, stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), want)
    assert bool((buf[:guard] == 0xA5).all()) and bool((buf[guard + n * fb:] == 0xA5).all())
    with pytest.raises(hdb.DeblockError, match="out of place"):
        dev.sao_batch_device(out, out, n, w, h, d_p, ctb, stream=st)
    with pytest.raises(hdb.DeblockError, match="16, 32 or 64"):
        dev.sao_batch_device(d_in, out, n, w, h, d_p, 48, stream=st)


# --- samples deeper than 8 bits: Sao16Kernel through hdb_sao_batch*_ex ---

@pytest.mark.parametrize("depth", [8, 10, 12])
@pytest.mark.parametrize("ctb", [16, 32, 64])
@pytest.mark.parametrize("size", [(8, 8), (64, 64), (72, 40), (200, 120), (352, 288), (1000, 64), (1920, 1080)], ids=lambda s: "%dx%d" % s)
def test_sao16_sizes_vs_oracle(dev, size, ctb, depth):
    from tests.test_depth import make_frame16
    from tests.test_sao import make_sao_params16
    w, h = size
    rng = np.random.default_rng(w * 11 + h + ctb + depth)
    for kind, content in [("mixed", "blocky"), ("edge", "random"), ("band", "ramp"), ("extreme", "saturated"), ("off", "steps")]:
        frame = make_frame16(rng, w, h, content, depth)
        params = make_sao_params16(rng, w, h, ctb, kind, depth)
        got = dev.sao_batch(frame[None], w, h, params[None], ctb, flags=hdb.flag_depth(depth))[0]
        want = orc.sao_frame16(frame, w, h, params, ctb, depth)
        assert np.array_equal(got, want), (size, ctb, depth, kind, content, int((got != want).sum()))


@pytest.mark.parametrize("cls", [0, 1, 2, 3])
def test_sao16_each_edge_class_everywhere(dev, cls):
    from tests.test_depth import make_frame16
    w, h, ctb, depth = 200, 120, 32, 10
    rng = np.random.default_rng(40 + cls)
    frame = make_frame16(rng, w, h, "random", depth)
    params = np.zeros((3, (h + ctb - 1) // ctb, (w + ctb - 1) // ctb), orc.SAO_PARAMS)
    params["type"] = 2
    params["position"] = cls
    params["offset"] = (31, 9, -9, -31)
    got = dev.sao_batch(frame[None], w, h, params[None], ctb, flags=hdb.flag_depth(depth))[0]
    assert np.array_equal(got, orc.sao_frame16(frame, w, h, params, ctb, depth))


def test_ten_bit_deblock_then_sao_on_the_device(dev):
    """The ten-bit in-loop pair == oracle deblocking at depth 10 followed by oracle SAO at depth 10; unaligned frames and
    depths out of range are refused."""
    import torch
    from tests.test_depth import make_frame16
    from tests.test_sao import make_sao_params16
    n, w, h, ctb, depth = 4, 1920, 1080, 64, 10
    rng = np.random.default_rng(5)
    frames = np.stack([make_frame16(rng, w, h, "blocky", depth) for _ in range(n)])
    qps = rng.integers(20, 52, n).astype(np.uint32)
    v = rng.integers(0, 3, (n, hdb.num_vert_bs(w, h))).astype(np.uint8)
    hh = rng.integers(0, 3, (n, hdb.num_hor_bs(w, h))).astype(np.uint8)
    params = np.stack([make_sao_params16(rng, w, h, ctb, "mixed", depth) for _ in range(n)])
    d_in = torch.from_numpy(frames.view(np.int16)).cuda()
    d_qp = torch.from_numpy(qps.astype(np.int32)).cuda()
    d_v, d_h = torch.from_numpy(v).cuda(), torch.from_numpy(hh).cuda()
    d_p = torch.from_numpy(params.view(np.uint8).reshape(n, -1)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    fs, guard = frames.shape[1], 1 << 15                       # samples per frame, guard band in samples
    buf = torch.full((guard + n * fs + guard,), 0x5A5A, dtype=torch.int16, device="cuda")
    out = buf[guard:guard + n * fs].view(n, fs)
    dev.deblock_sao_batch_device(d_in, out, n, w, h, d_qp, d_v, d_h, d_p, ctb, flags=hdb.flag_depth(depth), stream=st)
    torch.cuda.synchronize()
    deb = [orc.deblock_frame16(frames[f], w, h, int(qps[f]), v[f], hh[f], bit_depth=depth) for f in range(n)]
    want = np.stack([orc.sao_frame16(deb[f], w, h, params[f], ctb, depth) for f in range(n)])
    assert np.array_equal(out.cpu().numpy().view(np.uint16), want)
    assert bool((buf[:guard] == 0x5A5A).all()) and bool((buf[guard + n * fs:] == 0x5A5A).all())
    assert np.array_equal(d_in.cpu().numpy().view(np.uint16), frames)               # input untouched
    # the stage alone on a width off the 16 grid (two 8-byte moves per lane), between the same guard bands
    w2, h2 = 1000, 360
    f2 = np.stack([make_frame16(rng, w2, h2, "random", depth) for _ in range(2)])
    p2 = np.stack([make_sao_params16(rng, w2, h2, ctb, "mixed", depth) for _ in range(2)])
    fs2 = f2.shape[1]
    buf.fill_(0x5A5A)
    out2 = buf[guard:guard + 2 * fs2].view(2, fs2)
    dev.sao_batch_device(torch.from_numpy(f2.view(np.int16)).cuda(), out2, 2, w2, h2, torch.from_numpy(p2.view(np.uint8).reshape(2, -1)).cuda(), ctb,
                         stream=st, flags=hdb.flag_depth(depth))
    torch.cuda.synchronize()
    assert np.array_equal(out2.cpu().numpy().view(np.uint16), np.stack([orc.sao_frame16(f2[i], w2, h2, p2[i], ctb, depth) for i in range(2)]))
    assert bool((buf[:guard] == 0x5A5A).all()) and bool((buf[guard + 2 * fs2:] == 0x5A5A).all())
    with pytest.raises(hdb.DeblockError, match="bit depth"):
        dev.sao_batch_device(d_in, out, n, w, h, d_p, ctb, stream=st, flags=hdb.flag_depth(13))
    with pytest.raises(hdb.DeblockError, match="bit depth"):
        dev.deblock_sao_batch_device(d_in, out, n, w, h, d_qp, d_v, d_h, d_p, ctb, flags=hdb.flag_depth(12), stream=st)
    with pytest.raises(hdb.DeblockError, match="8-byte aligned"):
        dev.sao_batch_device(d_in.view(-1)[2:], out, 1, w, h, d_p, ctb, stream=st, flags=hdb.flag_depth(10))
