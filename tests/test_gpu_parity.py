"""GPU parity: the sm_100a path, called through the C ABI, against the oracle and the golden
vectors of the unmodified reference.  Bit-exact (integer byte work): every comparison is
array_equal / md5.  Run on the B200 box with  pytest -m gpu."""
import os

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


def test_device_is_b200(dev):
    info = dev.device_info()
    assert info["cc_major"] == 10 and info["sm_count"] >= 100, info


# ---- the reference's own frames (BASELINE configs 1 and 2) -------------------------------
@pytest.mark.parametrize("case", common.golden()["bundled"], ids=lambda c: "%s-qp%d" % (c["file"][:8], c["qp"]))
def test_bundled_frames(dev, case):
    x = common.fixture(case["file"])
    t = hdb.Timing()
    y = dev.deblock_frame(x, case["w"], case["h"], case["qp"], timing=t)
    assert common.md5(y) == case["md5"]
    assert common.sha256(y) == case["sha256"]
    assert t.path == 1 and t.kernel_launches == 1 and t.kernel_ms > 0 and t.total_ms >= t.kernel_ms


@pytest.mark.parametrize("case", common.golden()["synthetic"], ids=lambda c: "s%d-%dx%d" % (c["seed"], c["w"], c["h"]))
def test_golden_synthetic(dev, case):
    frame, v, hh = common.synth_case(case["seed"], case["w"], case["h"], case["kind"], case["bs_kind"])
    assert common.md5(frame) == case["input_md5"]
    y = dev.deblock_frame(frame, case["w"], case["h"], case["qp"], v, hh)
    assert common.md5(y) == case["md5"], "differs in %d bytes from the reference" % -1


# ---- sizes x content x QP x BS against the oracle -----------------------------------------
TILED_SIZES = [(32, 8), (32, 32), (64, 48), (256, 64), (288, 72), (352, 288), (768, 576), (1024, 8), (8192, 16),
               (1920, 1080), (2048, 1096)]
GENERIC_SIZES = [(8, 8), (16, 16), (24, 40)]                                         # padded-scratch kernels
# widths off the 32 grid (cpu.h:46-48 accepts every multiple of 8): the tiled kernel's cp.async loader
# (8 mod 16: the chroma planes are "skewed" -- the reference addresses them with a pitch 4 short of their padded
# width, cpu.h:450-451 with 469-471; 16 mod 32: straight)
LDGSTS_SIZES = [(40, 24), (72, 40), (200, 120), (360, 288), (1000, 64), (1080, 1920), (1912, 1080), (2056, 40),
                (48, 24), (80, 40), (1008, 64), (2000, 1048)]


@pytest.mark.parametrize("size", TILED_SIZES + GENERIC_SIZES + LDGSTS_SIZES, ids=lambda s: "%dx%d" % s)
def test_sizes_vs_oracle(dev, size):
    w, h = size
    rng = np.random.default_rng(w * 131 + h)
    for kind, bs_kind, qp in [("blocky", "uniform012", 37), ("steps", None, 30), ("saturated", "wide", 51),
                              ("random", "twos", 44), ("blocky", None, 22)]:
        frame = synth.make_frame(rng, w, h, kind)
        v, hh = (None, None) if bs_kind is None else synth.make_bs(rng, w, h, bs_kind)
        t = hdb.Timing()
        got = dev.deblock_frame(frame, w, h, qp, v, hh, timing=t)
        want = orc.deblock_frame(frame, w, h, qp, v, hh)
        assert t.path == common.expected_path(w)
        assert np.array_equal(got, want), (size, kind, bs_kind, qp, int((got != want).sum()))


@pytest.mark.parametrize("qp", [0, 15, 16, 17, 18, 26, 27, 35, 37, 47, 51, 52, 255, 2 ** 31])
def test_qp_sweep(dev, qp):
    rng = np.random.default_rng(qp % 1000)
    w, h = 352, 288
    frame = synth.make_frame(rng, w, h, "blocky")
    v, hh = synth.make_bs(rng, w, h, "uniform012")
    got = dev.deblock_frame(frame, w, h, qp, v, hh)
    assert np.array_equal(got, orc.deblock_frame(frame, w, h, qp, v, hh))
    if qp <= 17:
        assert np.array_equal(got, frame)  # no-op range


@pytest.mark.parametrize("bs_kind", ["zeros", "ones", "twos", "uniform012", "wide"])
@pytest.mark.parametrize("kind", ["flat", "ramp", "blocky", "random", "saturated", "steps"])
def test_bs_and_content(dev, kind, bs_kind):
    rng = np.random.default_rng(hash((kind, bs_kind)) & 0xffff)
    w, h = 320, 136
    frame = synth.make_frame(rng, w, h, kind)
    v, hh = synth.make_bs(rng, w, h, bs_kind)
    got = dev.deblock_frame(frame, w, h, 39, v, hh)
    assert np.array_equal(got, orc.deblock_frame(frame, w, h, 39, v, hh))
    if bs_kind == "zeros":
        assert np.array_equal(got[:w * h], frame[:w * h])  # luma untouched; chroma uses its own defaults


def test_planes_entry_point(dev):
    rng = np.random.default_rng(2)
    w, h = 352, 288
    frame = synth.make_frame(rng, w, h, "blocky")
    y, u, v = frame[:w * h], frame[w * h:w * h * 5 // 4], frame[w * h * 5 // 4:]
    yo, uo, vo = dev.deblock_planes(y, u, v, w, h, 35)
    want = orc.deblock_frame(frame, w, h, 35)
    assert np.array_equal(np.concatenate([yo, uo, vo]), want)


def test_execute_gpu_files(dev, tmp_path):
    # ExecuteGpu(input_file_name, output_file_name, width, height, Qp) -- main.cu:138
    case = [c for c in common.golden()["bundled"] if c["file"].startswith("mother") and c["qp"] == 35][0]
    fin, fout = tmp_path / "in.yuv", tmp_path / "out.yuv"
    common.fixture(case["file"]).tofile(fin)
    t = hdb.Timing()
    dev.execute_gpu(str(fin), str(fout), 352, 288, 35, timing=t)
    assert common.md5(np.fromfile(fout, np.uint8)) == case["md5"]
    assert t.h2d_bytes == 152064 and t.d2h_bytes == 152064
    with pytest.raises(hdb.DeblockError, match="Incorrect file size"):
        dev.execute_gpu(str(fin), str(fout), 352, 280, 35)
    with pytest.raises(hdb.DeblockError):
        dev.execute_gpu(str(tmp_path / "missing.yuv"), str(fout), 352, 288, 35)


def test_error_behaviour(dev):
    frame = np.zeros(3 * 16 * 16 // 2, np.uint8)
    with pytest.raises(hdb.DeblockError, match="multiplier of sample block size"):
        dev.deblock_frame(np.zeros(3 * 12 * 8 // 2, np.uint8), 12, 8, 30)
    with pytest.raises(hdb.DeblockError, match="boundary strenght"):
        dev.deblock_frame(frame, 16, 16, 30, np.zeros(5, np.uint8), np.zeros(6, np.uint8))
    with pytest.raises(hdb.DeblockError, match="Incorrect file size"):
        dev.deblock_frame(frame[:-2], 16, 16, 30)
    with pytest.raises(hdb.DeblockError):
        dev.deblock_frame(frame, 16, 16, 30, np.zeros(hdb.num_vert_bs(16, 16), np.uint8), None)
    # the context stays usable after errors
    assert np.array_equal(dev.deblock_frame(frame, 16, 16, 30), orc.deblock_frame(frame, 16, 16, 30))


def test_random_geometry_fuzz(dev):
    """Random frame sizes (every multiple of 8 up to 1280 x 384, both kernel paths), QPs, contents and
    boundary-strength maps against the oracle: chunk, row-group and plane-border handling off the
    beaten sizes."""
    rng = np.random.default_rng(20261018)
    kinds = ["blocky", "random", "steps", "saturated", "ramp"]
    bs_kinds = [None, "uniform012", "wide", "twos", "zeros"]
    for case in range(60):
        w = 8 * int(rng.integers(1, 161))
        h = 8 * int(rng.integers(1, 49))
        if case % 3 == 0:
            w = 32 * max(1, w // 32)          # make sure the tiled path gets its share
        qp = int(rng.integers(16, 56))
        kind = kinds[int(rng.integers(len(kinds)))]
        bs_kind = bs_kinds[int(rng.integers(len(bs_kinds)))]
        frame = synth.make_frame(rng, w, h, kind)
        v, hh = (None, None) if bs_kind is None else synth.make_bs(rng, w, h, bs_kind)
        got = dev.deblock_frame(frame, w, h, qp, v, hh)
        want = orc.deblock_frame(frame, w, h, qp, v, hh)
        assert np.array_equal(got, want), (case, w, h, qp, kind, bs_kind, int((got != want).sum()))


# ---- batches: host pipeline and device-resident -------------------------------------------
@pytest.mark.parametrize("n,w,h", [(1, 352, 288), (9, 352, 288), (40, 1920, 1080), (5, 24, 40), (3, 3840, 2160)])
def test_batch_host_pipeline(dev, n, w, h):
    frames, qps, v, hh = synth.make_batch(n * 1000 + w, n, w, h)
    t = hdb.Timing()
    got = dev.deblock_batch(frames, w, h, qps, v, hh, timing=t)
    want = orc.deblock_batch(frames, w, h, qps, v, hh)
    for f in range(n):
        assert np.array_equal(got[f], want[f]), (f, int(qps[f]))
    assert t.d2h_bytes == frames.nbytes and t.h2d_bytes >= frames.nbytes
    # in place, default strengths
    buf = frames.copy()
    dev.deblock_batch(buf, w, h, qps, out=buf)
    assert np.array_equal(buf, orc.deblock_batch(frames, w, h, qps))


@pytest.mark.parametrize("n,w,h", [(1, 352, 288), (5, 768, 576), (3, 24, 40), (1, 1920, 1080)])
def test_single_slice_on_mapped_pinned_buffers(dev, n, w, h):
    """One small slice in page-locked mapped memory: the kernel reads and writes the host buffers itself
    (no copy-engine legs); same bytes as the oracle, with caller strengths, defaults and in place."""
    import torch
    frames, qps, v, hh = synth.make_batch(n * 7 + w, n, w, h)
    t_in = torch.from_numpy(frames).pin_memory()
    t_out = torch.zeros_like(t_in).pin_memory()
    t_qp = torch.from_numpy(qps.astype(np.int32)).pin_memory()
    tm = hdb.Timing()
    dev.deblock_batch(t_in, w, h, t_qp, v, hh, out=t_out, timing=tm, n_frames=n)
    assert np.array_equal(t_out.numpy(), orc.deblock_batch(frames, w, h, qps, v, hh))
    assert tm.d2h_bytes == frames.nbytes
    if n == 1:
        assert tm.copy_ms == 0.0 and tm.kernel_launches >= 1         # one slice: the transfers happened inside the kernel
    assert np.array_equal(t_in.numpy(), frames)                      # input untouched
    dev.deblock_batch(t_in, w, h, t_qp, out=t_in, n_frames=n)        # defaults, in place
    assert np.array_equal(t_in.numpy(), orc.deblock_batch(frames, w, h, qps))


def test_batch_many_slices_pinned(dev):
    """More slices than pipeline slots, pinned torch buffers (the bench's e2e path)."""
    import torch
    n, w, h = 64, 1920, 1080  # slice plan 5,10,17,17,10,5 frames (128 MiB full slices, ramped): 6 slices over 4 slots
    frames, qps, v, hh = synth.make_batch(77, 8, w, h)
    reps = n // 8
    tf = torch.from_numpy(np.tile(frames, (reps, 1))).pin_memory()
    tq = torch.from_numpy(np.tile(qps, reps).astype(np.int32)).pin_memory()
    tv = torch.from_numpy(np.tile(v, (reps, 1))).pin_memory()
    th = torch.from_numpy(np.tile(hh, (reps, 1))).pin_memory()
    out = torch.empty_like(tf).pin_memory()
    dev.deblock_batch(tf, w, h, tq, tv, th, out=out, n_frames=n)
    want = orc.deblock_batch(frames, w, h, qps, v, hh)
    got = out.numpy()
    for f in range(n):
        assert np.array_equal(got[f], want[f % 8]), f


@pytest.mark.parametrize("in_place", [False, True])
def test_batch_device_resident(dev, in_place):
    import torch
    n, w, h = 12, 1920, 1080
    frames, qps, v, hh = synth.make_batch(5, n, w, h)
    d_in = torch.from_numpy(frames).cuda()
    d_out = d_in if in_place else torch.empty_like(d_in)
    d_qp = torch.from_numpy(qps.astype(np.int32)).cuda()
    d_v, d_h = torch.from_numpy(v).cuda(), torch.from_numpy(hh).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    dev.deblock_batch_device(d_in, d_out, n, w, h, d_qp, d_v, d_h, stream=stream)
    torch.cuda.synchronize()
    want = orc.deblock_batch(frames, w, h, qps, v, hh)
    assert np.array_equal(d_out.cpu().numpy(), want)
    if not in_place:
        assert np.array_equal(d_in.cpu().numpy(), frames)  # input untouched
    # default strengths on the device path
    d_in2 = torch.from_numpy(frames).cuda()
    dev.deblock_batch_device(d_in2, d_in2, n, w, h, d_qp, stream=stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_in2.cpu().numpy(), orc.deblock_batch(frames, w, h, qps))


@pytest.mark.parametrize("size", [(32, 8), (352, 288), (1920, 1080), (24, 40)], ids=lambda s: "%dx%d" % s)
def test_no_write_outside_the_frames(dev, size):
    """compute-sanitizer is closed on this pool, so out-of-bounds stores are hunted with canaries: the
    output batch sits between two 64 KiB guard bands that must come back untouched (tiled path: stores
    the reference would make into its padding must be dropped; generic path: crop must stay inside)."""
    import torch
    w, h = size
    n, guard = 3, 1 << 16
    frames, qps, v, hh = synth.make_batch(w + h, n, w, h)
    fb = frames.shape[1]
    buf = torch.full((guard + n * fb + guard,), 0xA5, dtype=torch.uint8, device="cuda")
    d_out = buf[guard:guard + n * fb]
    d_in = torch.from_numpy(frames).cuda()
    d_qp = torch.from_numpy(qps.astype(np.int32)).cuda()
    d_v, d_h = torch.from_numpy(v).cuda(), torch.from_numpy(hh).cuda()
    dev.deblock_batch_device(d_in, d_out, n, w, h, d_qp, d_v, d_h, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy().reshape(n, fb), orc.deblock_batch(frames, w, h, qps, v, hh))
    assert bool((buf[:guard] == 0xA5).all()) and bool((buf[guard + n * fb:] == 0xA5).all())
    # in place inside the same guarded buffer
    buf[guard:guard + n * fb] = d_in.reshape(-1)
    dev.deblock_batch_device(d_out, d_out, n, w, h, d_qp, d_v, d_h, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy().reshape(n, fb), orc.deblock_batch(frames, w, h, qps, v, hh))
    assert bool((buf[:guard] == 0xA5).all()) and bool((buf[guard + n * fb:] == 0xA5).all())


@pytest.mark.parametrize("case", [(1, 352, 288, 64, 32), (3, 1920, 1080, 128, 64), (2, 3840, 2160, 0, 0), (4, 64, 48, 16, 48)],
                         ids=lambda c: "%dx%dx%d+%d" % c[:4])
def test_pitched_surfaces_in_device_memory(dev, case):
    """Decoder-style surfaces: three separate planes with row pitches above the width.  Same bytes as the
    oracle on the samples, pitch padding untouched, out of place and in place."""
    import torch
    n, w, h, pad_y, pad_c = case
    cw, ch = w // 2, h // 2
    frames, qps, v, hh = synth.make_batch(w + n, n, w, h)
    want = orc.deblock_batch(frames, w, h, qps, v, hh)
    py, pc = w + pad_y, cw + pad_c

    def planes(fill):
        return (torch.full((n, h, py), fill, dtype=torch.uint8, device="cuda"),
                torch.full((n, ch, pc), fill, dtype=torch.uint8, device="cuda"),
                torch.full((n, ch, pc), fill, dtype=torch.uint8, device="cuda"))

    def surf(p):
        return hdb.Surfaces(p[0].data_ptr(), p[1].data_ptr(), p[2].data_ptr(), py, pc, h * py, ch * pc)

    src, dst = planes(0x5A), planes(0xA5)
    t = torch.from_numpy(frames).cuda()
    src[0][:, :, :w] = t[:, :w * h].view(n, h, w)
    src[1][:, :, :cw] = t[:, w * h:w * h + cw * ch].view(n, ch, cw)
    src[2][:, :, :cw] = t[:, w * h + cw * ch:].view(n, ch, cw)
    d_qp = torch.from_numpy(qps.astype(np.int32)).cuda()
    d_v, d_h = torch.from_numpy(v).cuda(), torch.from_numpy(hh).cuda()
    st = torch.cuda.current_stream().cuda_stream

    def gather(p):
        return torch.cat([p[0][:, :, :w].reshape(n, -1), p[1][:, :, :cw].reshape(n, -1), p[2][:, :, :cw].reshape(n, -1)], 1).cpu().numpy()

    dev.deblock_surfaces_device(surf(src), surf(dst), n, w, h, d_qp, d_v, d_h, stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(gather(dst), want)
    assert np.array_equal(gather(src), frames)                                    # input untouched
    for p, width in zip(dst, (w, cw, cw)):
        assert bool((p[:, :, width:] == 0xA5).all())                              # pitch padding untouched
    dev.deblock_surfaces_device(surf(src), surf(src), n, w, h, d_qp, d_v, d_h, stream=st)   # in place
    torch.cuda.synchronize()
    assert np.array_equal(gather(src), want)
    for p, width in zip(src, (w, cw, cw)):
        assert bool((p[:, :, width:] == 0x5A).all())
    with pytest.raises(hdb.DeblockError):
        dev.deblock_surfaces_device(surf(src), surf(dst), n, w + 8, h, d_qp, stream=st)      # width not a multiple of 32


def test_launches_overlapping_on_two_streams(dev):
    """Two streams, one context: launches that overlap on the device must not share a work counter."""
    import torch
    n, w, h = 24, 1920, 1080
    frames, qps, v, hh = synth.make_batch(11, 6, w, h)
    reps = n // 6
    want = torch.from_numpy(orc.deblock_batch(frames, w, h, qps, v, hh)).cuda()
    d_in = torch.from_numpy(frames).cuda().repeat(reps, 1)
    d_qp = torch.from_numpy(qps.astype(np.int32)).cuda().repeat(reps)
    d_v, d_h = torch.from_numpy(v).cuda().repeat(reps, 1), torch.from_numpy(hh).cuda().repeat(reps, 1)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [[torch.empty_like(d_in) for _ in range(6)] for _ in streams]
    torch.cuda.synchronize()
    for i in range(6):
        for st, o in zip(streams, outs):
            dev.deblock_batch_device(d_in, o[i], n, w, h, d_qp, d_v, d_h, stream=st.cuda_stream)
    torch.cuda.synchronize()
    for o in outs:
        for t in o:
            assert bool((t.view(reps, 6, -1) == want.unsqueeze(0)).all())


def test_repeated_launches_reset_scheduler(dev):
    """The work counter is reset by the kernel itself; many back-to-back launches must all be complete."""
    import torch
    n, w, h = 6, 768, 576
    frames, qps, v, hh = synth.make_batch(8, n, w, h)
    want = orc.deblock_batch(frames, w, h, qps, v, hh)
    d_qp = torch.from_numpy(qps.astype(np.int32)).cuda()
    d_v, d_h = torch.from_numpy(v).cuda(), torch.from_numpy(hh).cuda()
    src = torch.from_numpy(frames).cuda()
    outs = [torch.empty_like(src) for _ in range(20)]
    for o in outs:
        dev.deblock_batch_device(src, o, n, w, h, d_qp, d_v, d_h, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for o in outs:
        assert np.array_equal(o.cpu().numpy(), want)


# ---- BASELINE's full sizes through size-independent properties -----------------------------
def test_full_size_batch_properties(dev):
    """256 x 1080p resident in HBM (BASELINE config 3 size): every frame is a replica of one of 8
    distinct (frame, QP, BS) inputs, so the output must be 32 identical copies of the oracle's 8."""
    import torch
    n, w, h, distinct = 256, 1920, 1080, 8
    frames, qps, v, hh = synth.make_batch(123, distinct, w, h)
    reps = n // distinct
    d_in = torch.from_numpy(frames).cuda().repeat(reps, 1)
    d_qp = torch.from_numpy(qps.astype(np.int32)).cuda().repeat(reps)
    d_v, d_h = torch.from_numpy(v).cuda().repeat(reps, 1), torch.from_numpy(hh).cuda().repeat(reps, 1)
    d_out = torch.empty_like(d_in)
    dev.deblock_batch_device(d_in, d_out, n, w, h, d_qp, d_v, d_h, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = torch.from_numpy(orc.deblock_batch(frames, w, h, qps, v, hh)).cuda()
    assert bool((d_out.view(reps, distinct, -1) == want.unsqueeze(0)).all())
    # idempotence of the no-op QPs and determinism of a second run
    d_out2 = torch.empty_like(d_in)
    dev.deblock_batch_device(d_in, d_out2, n, w, h, d_qp, d_v, d_h, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert bool((d_out == d_out2).all())


def test_8k_frame(dev):
    case = [c for c in common.golden()["synthetic"] if c["w"] == 7680][0]
    frame, v, hh = common.synth_case(case["seed"], 7680, 4320, case["kind"], case["bs_kind"])
    y = dev.deblock_frame(frame, 7680, 4320, case["qp"], v, hh)
    assert common.md5(y) == case["md5"]


def test_padded_scratch_kernels_still_match(monkeypatch):
    """The padded-scratch kernels serve widths below 32 only; a context created with HDB_NO_LDGSTS set sends
    every width off the 32 grid through them, so that they stay covered at real sizes."""
    monkeypatch.setenv("HDB_NO_LDGSTS", "1")
    d = hdb.Deblocker(0)
    monkeypatch.delenv("HDB_NO_LDGSTS")
    rng = np.random.default_rng(77)
    for (w, h) in [(72, 40), (200, 120), (1912, 1080)]:
        frame = synth.make_frame(rng, w, h, "blocky")
        v, hh = synth.make_bs(rng, w, h, "uniform012")
        t = hdb.Timing()
        got = d.deblock_frame(frame, w, h, 37, v, hh, timing=t)
        assert t.path == 2
        assert np.array_equal(got, orc.deblock_frame(frame, w, h, 37, v, hh))
    d.close()


def test_ldgsts_batch_device_resident_with_guards(dev):
    """A batch of 1080x1920 (portrait: width off the 32 grid) frames resident in HBM between guard bands,
    out of place and in place."""
    import torch
    n, w, h = 6, 1080, 1920
    frames, qps, v, hh = synth.make_batch(5, n, w, h)
    want = orc.deblock_batch(frames, w, h, qps, v, hh)
    fb, guard = frames.shape[1], 1 << 16
    buf = torch.full((guard + n * fb + guard,), 0xA5, dtype=torch.uint8, device="cuda")
    d_in = torch.from_numpy(frames).cuda()
    d_qp = torch.from_numpy(qps.astype(np.int32)).cuda()
    d_v, d_h = torch.from_numpy(v).cuda(), torch.from_numpy(hh).cuda()
    st = torch.cuda.current_stream().cuda_stream
    out = buf[guard:guard + n * fb].view(n, fb)
    dev.deblock_batch_device(d_in, out, n, w, h, d_qp, d_v, d_h, stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), want)
    assert bool((buf[:guard] == 0xA5).all()) and bool((buf[guard + n * fb:] == 0xA5).all())
    out.copy_(d_in)
    dev.deblock_batch_device(out, out, n, w, h, d_qp, d_v, d_h, stream=st)     # in place
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), want)
    assert bool((buf[:guard] == 0xA5).all()) and bool((buf[guard + n * fb:] == 0xA5).all())


def test_small_frames_in_flight_on_two_streams(dev):
    """Two calls of the padded-scratch path in flight on different streams: each has scratch memory of its own."""
    import torch
    n, w, h = 64, 24, 40
    a, qa, va, ha = synth.make_batch(1, n, w, h)
    b, qb, vb, hb = synth.make_batch(2, n, w, h)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    oa, ob = torch.empty_like(da), torch.empty_like(db)
    dqa, dqb = torch.from_numpy(qa.astype(np.int32)).cuda(), torch.from_numpy(qb.astype(np.int32)).cuda()
    dva, dha, dvb, dhb = (torch.from_numpy(x).cuda() for x in (va, ha, vb, hb))
    torch.cuda.synchronize()
    for _ in range(20):
        dev.deblock_batch_device(da, oa, n, w, h, dqa, dva, dha, stream=s1.cuda_stream)
        dev.deblock_batch_device(db, ob, n, w, h, dqb, dvb, dhb, stream=s2.cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(oa.cpu().numpy(), orc.deblock_batch(a, w, h, qa, va, ha))
    assert np.array_equal(ob.cpu().numpy(), orc.deblock_batch(b, w, h, qb, vb, hb))


def test_planes_entry_point_from_two_threads(dev):
    """hdb_deblock_planes holds the context lock from its first copy to its last: two threads on one context."""
    import threading
    rng = np.random.default_rng(9)
    cases = []
    for (w, h) in [(352, 288), (768, 576)]:
        frame = synth.make_frame(rng, w, h, "blocky")
        cases.append((w, h, frame, orc.deblock_frame(frame, w, h, 35)))
    bad = []

    def work(case):
        w, h, frame, want = case
        y, u, v = frame[:w * h], frame[w * h:w * h * 5 // 4], frame[w * h * 5 // 4:]
        for _ in range(25):
            yo, uo, vo = dev.deblock_planes(y, u, v, w, h, 35)
            if not np.array_equal(np.concatenate([yo, uo, vo]), want):
                bad.append((w, h))
    ths = [threading.Thread(target=work, args=(c,)) for c in cases]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not bad


def test_frame_too_wide_is_refused(dev):
    with pytest.raises(hdb.DeblockError, match="too large"):
        dev.deblock_batch_device(1 << 20, 1 << 21, 1, 262144, 8, 1 << 22)


@pytest.mark.parametrize("size", [(72, 400), (40, 544), (952, 344), (1528, 600), (136, 280)], ids=lambda s: "%dx%d" % s)
def test_skewed_chroma_out_of_place_reaches_the_last_rows(dev, size):
    """Frame widths of 8 mod 16: the reference's chroma blocks tile a virtual plane whose rows are 4 bytes short
    of the buffer's, so its block rows end above the last image rows, which it leaves alone.  Out of place those
    rows must still arrive in the output (found by tools/soak.py: the first version of the skew mapping moved only
    the rows the reference's blocks reach)."""
    import torch
    w, h = size
    n = 3
    frames, qps, v, hh = synth.make_batch(w + h, n, w, h)
    want = orc.deblock_batch(frames, w, h, qps, v, hh)
    d_in = torch.from_numpy(frames).cuda()
    d_out = torch.full_like(d_in, 0xA5)
    d_qp = torch.from_numpy(qps.astype(np.int32)).cuda()
    d_v, d_h = torch.from_numpy(v).cuda(), torch.from_numpy(hh).cuda()
    st = torch.cuda.current_stream().cuda_stream
    dev.deblock_batch_device(d_in, d_out, n, w, h, d_qp, d_v, d_h, stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), want)
    dev.deblock_batch_device(d_in, d_in, n, w, h, d_qp, d_v, d_h, stream=st)      # and in place
    torch.cuda.synchronize()
    assert np.array_equal(d_in.cpu().numpy(), want)
