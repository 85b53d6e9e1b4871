"""GPU parity of the NV12 entry points through the C ABI: the reference's golden frames stored as NV12,
the oracle on many sizes, host pipeline, device-resident batches, pitched surfaces with canary
padding, with and without edge maps.  Bit-exact.  Run with  pytest -m gpu."""
import numpy as np
import pytest

import gpu_video_codec_b200 as hdb
from oracle import oracle_py as orc
from tests import common, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    d = hdb.Deblocker(0)
    yield d
    d.close()


def _nv12_batch(seed, n, w, h, edges=False):
    rng = np.random.default_rng(seed)
    frames = np.stack([orc.yv12_to_nv12(synth.make_frame(rng, w, h, ("blocky", "steps", "saturated", "random")[f % 4]), w, h)
                       for f in range(n)])
    qps = rng.integers(18, 52, n).astype(np.uint32)
    if edges:
        maps = [synth.make_edge_maps(rng, w, h, "smooth", picture_edges=bool(f & 1)) for f in range(n)]
    else:
        maps = [synth.make_bs(rng, w, h, "uniform012") for _ in range(n)]
    return frames, qps, np.stack([m[0] for m in maps]), np.stack([m[1] for m in maps])


def _oracle(frames, w, h, qps, v, hh, edges=False):
    return np.stack([orc.deblock_frame_nv12(frames[f], w, h, int(qps[f]), None if v is None else v[f],
                                            None if hh is None else hh[f], edges=edges) for f in range(frames.shape[0])])


@pytest.mark.parametrize("case", common.golden()["bundled"][:3] + [c for c in common.golden()["bundled"] if c["w"] == 768][:1],
                         ids=lambda c: "%s-qp%d" % (c["file"][:8], c["qp"]))
def test_reference_frames_as_nv12(dev, case):
    w, h = case["w"], case["h"]
    x = orc.yv12_to_nv12(common.fixture(case["file"]), w, h)
    qp = np.array([case["qp"]], np.uint32)
    y = dev.deblock_batch_ex(x[None, :], w, h, qp, flags=hdb.FLAG_NV12)
    assert common.md5(orc.nv12_to_yv12(y[0], w, h)) == case["md5"]


SIZES = [(32, 8), (32, 32), (64, 48), (256, 64), (288, 72), (352, 288), (480, 272), (512, 16), (544, 32), (768, 576),
         (8192, 16), (1920, 1080), (2048, 1096)]


@pytest.mark.parametrize("size", SIZES, ids=lambda s: "%dx%d" % s)
def test_nv12_sizes_vs_oracle(dev, size):
    w, h = size
    for edges in (False, True):
        frames, qps, v, hh = _nv12_batch(w * 11 + h + edges, 3, w, h, edges)
        got = dev.deblock_batch_ex(frames, w, h, qps, v, hh, flags=hdb.FLAG_NV12 | (hdb.FLAG_EDGE_MAPS if edges else 0))
        want = _oracle(frames, w, h, qps, v, hh, edges)
        for f in range(3):
            assert np.array_equal(got[f], want[f]), (size, edges, f, int((got[f] != want[f]).sum()))
    # default strengths
    got = dev.deblock_batch_ex(frames, w, h, qps, flags=hdb.FLAG_NV12)
    assert np.array_equal(got, _oracle(frames, w, h, qps, None, None))


@pytest.mark.parametrize("size", [(48, 24), (80, 40), (208, 120), (1008, 64), (2000, 1048)], ids=lambda s: "%dx%d" % s)
def test_nv12_widths_off_the_tile_grid(dev, size):
    """Widths that are a multiple of 16 but not of 32: the cp.async loader of the tiled kernel."""
    w, h = size
    for edges in (False, True):
        frames, qps, v, hh = _nv12_batch(w * 13 + h + edges, 2, w, h, edges)
        t = hdb.Timing()
        got = dev.deblock_batch_ex(frames, w, h, qps, v, hh, flags=hdb.FLAG_NV12 | (hdb.FLAG_EDGE_MAPS if edges else 0), timing=t)
        want = _oracle(frames, w, h, qps, v, hh, edges)
        assert t.path == 3
        for f in range(2):
            assert np.array_equal(got[f], want[f]), (size, edges, f, int((got[f] != want[f]).sum()))


def test_nv12_rejects_what_no_kernel_takes(dev):
    """Widths below 32, and widths of 8 mod 16: there the reference's chroma blocks are skewed against the
    samples (chroma width 4 mod 8, cpu.h:450-451 with 469-471) and straddle the interleaving."""
    for (w, h) in [(24, 40), (72, 40), (1080, 1920)]:
        frames = np.zeros((1, 3 * w * h // 2), np.uint8)
        with pytest.raises(hdb.DeblockError, match="multiple of 16"):
            dev.deblock_batch_ex(frames, w, h, np.array([30], np.uint32), flags=hdb.FLAG_NV12)
    w, h = 64, 40
    frames = np.zeros((1, 3 * w * h // 2), np.uint8)
    with pytest.raises(hdb.DeblockError):
        dev.deblock_batch_ex(frames, w, h, np.array([30], np.uint32), flags=64)


def test_nv12_batch_host_pipeline_many_slices(dev):
    import torch
    n, w, h = 48, 1920, 1080
    frames, qps, v, hh = _nv12_batch(21, 6, w, h)
    reps = n // 6
    tf = torch.from_numpy(np.tile(frames, (reps, 1))).pin_memory()
    tq = torch.from_numpy(np.tile(qps, reps).astype(np.int32)).pin_memory()
    tv = torch.from_numpy(np.tile(v, (reps, 1))).pin_memory()
    th = torch.from_numpy(np.tile(hh, (reps, 1))).pin_memory()
    out = torch.empty_like(tf).pin_memory()
    t = hdb.Timing()
    dev.deblock_batch_ex(tf, w, h, tq, tv, th, flags=hdb.FLAG_NV12, out=out, n_frames=n, timing=t)
    want = _oracle(frames, w, h, qps, v, hh)
    got = out.numpy()
    for f in range(n):
        assert np.array_equal(got[f], want[f % 6]), f
    assert t.path == 1


@pytest.mark.parametrize("in_place", [False, True])
def test_nv12_device_resident_with_guards(dev, in_place):
    import torch
    n, w, h = 5, 1920, 1080
    frames, qps, v, hh = _nv12_batch(4, n, w, h)
    fb, guard = frames.shape[1], 1 << 16
    buf = torch.full((guard + n * fb + guard,), 0xA5, dtype=torch.uint8, device="cuda")
    d_out = buf[guard:guard + n * fb]
    if in_place:
        d_out.copy_(torch.from_numpy(frames).reshape(-1))
        d_in = d_out
    else:
        d_in = torch.from_numpy(frames).cuda()
    d_qp = torch.from_numpy(qps.astype(np.int32)).cuda()
    d_v, d_h = torch.from_numpy(v).cuda(), torch.from_numpy(hh).cuda()
    dev.deblock_batch_device_ex(d_in, d_out, n, w, h, d_qp, d_v, d_h, flags=hdb.FLAG_NV12,
                                stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy().reshape(n, fb), _oracle(frames, w, h, qps, v, hh))
    assert bool((buf[:guard] == 0xA5).all()) and bool((buf[guard + n * fb:] == 0xA5).all())


@pytest.mark.parametrize("case", [(1, 352, 288, 64, 32, False), (3, 1920, 1080, 128, 64, True), (2, 3840, 2160, 0, 0, False),
                                  (4, 64, 48, 16, 48, True)], ids=lambda c: "%dx%dx%d+%d" % c[:4])
def test_nv12_pitched_surfaces(dev, case):
    import torch
    n, w, h, pad_y, pad_uv, edges = case
    ch = h // 2
    frames, qps, v, hh = _nv12_batch(w + n, n, w, h, edges)
    want = _oracle(frames, w, h, qps, v, hh, edges)
    py, puv = w + pad_y, w + pad_uv

    def planes(fill):
        return (torch.full((n, h, py), fill, dtype=torch.uint8, device="cuda"),
                torch.full((n, ch, puv), fill, dtype=torch.uint8, device="cuda"))

    def surf(p):
        return hdb.Nv12Surfaces(p[0].data_ptr(), p[1].data_ptr(), py, puv, h * py, ch * puv)

    def gather(p):
        return torch.cat([p[0][:, :, :w].reshape(n, -1), p[1][:, :, :w].reshape(n, -1)], 1).cpu().numpy()

    src, dst = planes(0x5A), planes(0xA5)
    t = torch.from_numpy(frames).cuda()
    src[0][:, :, :w] = t[:, :w * h].view(n, h, w)
    src[1][:, :, :w] = t[:, w * h:].view(n, ch, w)
    d_qp = torch.from_numpy(qps.astype(np.int32)).cuda()
    d_v, d_h = torch.from_numpy(v).cuda(), torch.from_numpy(hh).cuda()
    st = torch.cuda.current_stream().cuda_stream
    fl = hdb.FLAG_EDGE_MAPS if edges else 0
    dev.deblock_nv12_surfaces_device(surf(src), surf(dst), n, w, h, d_qp, d_v, d_h, flags=fl, stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(gather(dst), want)
    assert np.array_equal(gather(src), frames)
    for p in dst:
        assert bool((p[:, :, w:] == 0xA5).all())
    dev.deblock_nv12_surfaces_device(surf(src), surf(src), n, w, h, d_qp, d_v, d_h, flags=fl, stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(gather(src), want)
    for p in src:
        assert bool((p[:, :, w:] == 0x5A).all())


def test_nv12_equals_yv12_at_full_size(dev):
    """64 x 4K resident in HBM: the NV12 kernel's output, de-interleaved, must be the planar kernel's output
    on the same frames stored as YV12 (which the rest of the suite ties to the reference)."""
    import torch
    n, w, h = 64, 3840, 2160
    ysz, cw, ch = w * h, w // 2, h // 2
    g = torch.Generator(device="cuda").manual_seed(5)
    planar = torch.randint(96, 160, (n, 3 * w * h // 2), device="cuda", generator=g, dtype=torch.int16).to(torch.uint8)
    planar[:, :ysz] = (planar[:, :ysz].view(n, h // 8, 8, w // 8, 8)
                       + torch.randint(0, 24, (n, h // 8, 1, w // 8, 1), device="cuda", generator=g, dtype=torch.uint8)).view(n, -1)
    nv = planar.clone()
    nv[:, ysz:] = torch.stack([planar[:, ysz:ysz + cw * ch], planar[:, ysz + cw * ch:]], dim=2).reshape(n, -1)
    d_qp = torch.randint(20, 52, (n,), device="cuda", generator=g, dtype=torch.int32)
    d_v = torch.randint(0, 3, (n, hdb.num_vert_bs(w, h)), device="cuda", generator=g, dtype=torch.uint8)
    d_h = torch.randint(0, 3, (n, hdb.num_hor_bs(w, h)), device="cuda", generator=g, dtype=torch.uint8)
    st = torch.cuda.current_stream().cuda_stream
    out_p, out_n = torch.empty_like(planar), torch.empty_like(nv)
    dev.deblock_batch_device(planar, out_p, n, w, h, d_qp, d_v, d_h, stream=st)
    dev.deblock_batch_device_ex(nv, out_n, n, w, h, d_qp, d_v, d_h, flags=hdb.FLAG_NV12, stream=st)
    torch.cuda.synchronize()
    assert bool((out_n[:, :ysz] == out_p[:, :ysz]).all())
    uv = out_n[:, ysz:].view(n, cw * ch, 2)
    assert bool((uv[:, :, 0] == out_p[:, ysz:ysz + cw * ch]).all()) and bool((uv[:, :, 1] == out_p[:, ysz + cw * ch:]).all())
    assert float((out_p != planar).float().mean()) > 0.02


@pytest.mark.parametrize("nv12", [False, True])
def test_sequence_file(dev, tmp_path, nv12):
    """hdb_execute_gpu_sequence: a raw file with several frames, YV12 and NV12, and the reference's two file errors."""
    n, w, h, qp = 7, 352, 288, 35
    rng = np.random.default_rng(12)
    frames = np.stack([synth.make_frame(rng, w, h, "blocky") for _ in range(n)])
    frames[0] = common.fixture("mother-daughter_352x288_yv12.yuv")
    stored = np.stack([orc.yv12_to_nv12(f, w, h) for f in frames]) if nv12 else frames
    fin, fout = tmp_path / "seq.yuv", tmp_path / "seq_out.yuv"
    stored.tofile(fin)
    t = hdb.Timing()
    assert dev.execute_gpu_sequence(str(fin), str(fout), w, h, qp, flags=hdb.FLAG_NV12 if nv12 else 0, timing=t) == n
    got = np.fromfile(fout, np.uint8).reshape(n, -1)
    for f in range(n):
        want = orc.deblock_frame(frames[f], w, h, qp)
        assert np.array_equal(orc.nv12_to_yv12(got[f], w, h) if nv12 else got[f], want), f
    case = [c for c in common.golden()["bundled"] if c["file"].startswith("mother") and c["qp"] == 35][0]
    assert common.md5(orc.nv12_to_yv12(got[0], w, h) if nv12 else got[0]) == case["md5"]
    assert t.d2h_bytes == frames.nbytes
    stored.reshape(-1)[:-5].tofile(fin)
    with pytest.raises(hdb.DeblockError, match="Incorrect file size"):
        dev.execute_gpu_sequence(str(fin), str(fout), w, h, qp)
    np.zeros(2 * 3 * 12 * 8 // 2, np.uint8).tofile(fin)
    with pytest.raises(hdb.DeblockError, match="multiplier of sample block size"):
        dev.execute_gpu_sequence(str(fin), str(fout), 12, 8, qp)


def test_demo_driver(tmp_path):
    """host/main.cpp (the reference's main.cu with arguments) over the C++ mirror: one frame goes through
    b200::ReadYuvFrame + b200::ExecuteGpu, a longer file through b200::ExecuteGpuSequence."""
    import os
    import subprocess
    demo = os.path.join(os.path.dirname(hdb.library_path()), "hevc_deblock_b200_demo")
    if not os.path.exists(demo):
        pytest.skip("demo driver not built")
    case = [c for c in common.golden()["bundled"] if c["file"].startswith("mother") and c["qp"] == 35][0]
    x = common.fixture(case["file"])
    fin, fout = tmp_path / "one.yuv", tmp_path / "one_out.yuv"
    x.tofile(fin)
    r = subprocess.run([demo, str(fin), "352", "288", "35", str(fout)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    for line in ("Compute Capability: 10.", "Execution Time without copy on GPU:", "Execution Time with copy on GPU:",
                 "Copy Operation Time with GPU buffers:"):                       # main.cu:92-107, gpu.cu:1292, 1302-1303
        assert line in r.stdout, r.stdout
    assert common.md5(np.fromfile(fout, np.uint8)) == case["md5"]
    seq, sout = tmp_path / "seq.yuv", tmp_path / "seq_out.yuv"
    np.tile(x, 5).tofile(seq)
    r = subprocess.run([demo, str(seq), "352", "288", "35", str(sout)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "Frames: 5" in r.stdout, r.stdout + r.stderr
    got = np.fromfile(sout, np.uint8).reshape(5, -1)
    assert all(common.md5(g) == case["md5"] for g in got)
    r = subprocess.run([demo, str(seq), "352", "280", "35", str(sout)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "Incorrect file size" in r.stderr               # cpu.h:43-45


def test_cpp_mirror_selftest(tmp_path):
    """host/selftest.cpp: b200::ReadYuvFrame (ctor / SetBoundaryStrenght / SetEdgeMaps / DeblockingFilter / Save and
    the reference's thrown messages) against the C ABI in-process; its two output files against the oracle here."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(hdb.library_path()), "hevc_deblock_b200_selftest")
    if not os.path.exists(exe):
        pytest.skip("selftest not built")
    w, h, qp = 352, 288, 35
    x = common.fixture("mother-daughter_352x288_yv12.yuv")
    fin, f_plain, f_edges = tmp_path / "in.yuv", tmp_path / "plain.yuv", tmp_path / "edges.yuv"
    x.tofile(fin)
    r = subprocess.run([exe, str(fin), str(w), str(h), str(qp), str(f_plain), str(f_edges)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "selftest ok" in r.stdout, r.stdout + r.stderr
    nv, nh = hdb.num_vert_bs(w, h), hdb.num_hor_bs(w, h)
    maps = np.fromfile(str(f_plain) + ".maps", np.uint8)
    v, hh, ve, he = maps[:nv], maps[nv:nv + nh], maps[nv + nh:2 * nv + nh], maps[2 * nv + nh:]
    assert np.array_equal(np.fromfile(f_plain, np.uint8), orc.deblock_frame(x, w, h, qp, v, hh))
    assert np.array_equal(np.fromfile(f_edges, np.uint8), orc.deblock_frame_edges(x, w, h, qp, ve, he))
