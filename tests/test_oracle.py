"""The oracle (oracle/deblock_oracle.c) against the golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py), and -- where oracle/_ref is built -- against the reference
itself on fresh random inputs, including the reference's error behaviour."""
import numpy as np
import pytest

from oracle import oracle_py as orc
from tests import common, synth


def test_survey_hashes_present():
    # the three headline hashes of SURVEY.md section 4 / BASELINE.md
    want = {("mother-daughter_352x288_yv12.yuv", 35): "814a18500ea45feb9c73b7f557709529",
            ("image1_352x288_yv12.yuv", 30): "372ebccb963fba860ae53c8275209f10",
            ("image2_768x576.yuv", 30): "cb1123d8396e374af6f780461f6df4ab"}
    have = {(c["file"], c["qp"]): c["md5"] for c in common.golden()["bundled"]}
    for k, v in want.items():
        assert have[k] == v


@pytest.mark.parametrize("case", common.golden()["bundled"], ids=lambda c: "%s-qp%d" % (c["file"][:8], c["qp"]))
def test_oracle_bundled(case):
    x = common.fixture(case["file"])
    assert common.md5(x) == case["input_md5"]
    y = orc.deblock_frame(x, case["w"], case["h"], case["qp"])
    assert common.md5(y) == case["md5"]
    assert common.sha256(y) == case["sha256"]
    assert int((x != y).sum()) == case["changed_bytes"]


@pytest.mark.parametrize("case", [c for c in common.golden()["synthetic"] if c["w"] * c["h"] <= 3840 * 2160],
                         ids=lambda c: "s%d-%dx%d" % (c["seed"], c["w"], c["h"]))
def test_oracle_synthetic(case):
    frame, v, hh = common.synth_case(case["seed"], case["w"], case["h"], case["kind"], case["bs_kind"])
    assert common.md5(frame) == case["input_md5"], "tests/synth.py changed: regenerate tests/golden"
    y = orc.deblock_frame(frame, case["w"], case["h"], case["qp"], v, hh)
    assert common.md5(y) == case["md5"]


def test_tables_and_sizes():
    # cpu.h:1021-1033 spot values and the clamp
    assert [orc.port().hdbo_beta(q) for q in (0, 15, 16, 28, 29, 51, 52, 1000)] == [0, 0, 6, 18, 20, 64, 64, 64]
    assert [orc.port().hdbo_tc(q) for q in (0, 17, 18, 26, 27, 47, 51, 99)] == [0, 0, 1, 1, 2, 13, 20, 20]
    # SURVEY.md section 8 size table
    assert (orc.num_vert_bs(352, 288), orc.num_hor_bs(352, 288)) == (1620, 1628)
    assert (orc.num_vert_bs(1920, 1080), orc.num_hor_bs(1920, 1080)) == (32535, 32640)
    assert (orc.num_vert_bs(960, 540), orc.num_hor_bs(960, 540)) == (8167, 8160)
    assert (orc.num_vert_bs(7680, 4320), orc.num_hor_bs(7680, 4320)) == (518940, 519360)


def test_default_bs_pattern():
    v, hh = orc.default_bs(352, 288)
    assert v.reshape(36, 45)[:, 0].max() == 0 and v.reshape(36, 45)[:, 1:].min() == 2
    assert set(np.flatnonzero(hh == 0)) == set(range(0, hh.size, 288 // 8 + 1))  # scattered, not the top row


def test_errors():
    with pytest.raises(orc.OracleError):
        orc.deblock_frame(np.zeros(3 * 12 * 8 // 2, np.uint8), 12, 8, 30)
    with pytest.raises(orc.OracleError):
        orc.deblock_frame(np.zeros(3 * 16 * 16 // 2, np.uint8), 16, 16, 30, np.zeros(5, np.uint8), np.zeros(6, np.uint8))
    with pytest.raises(orc.OracleError):
        orc.deblock_frame(np.zeros(100, np.uint8), 16, 16, 30)


def test_batch_equals_frames():
    frames, qps, v, hh = synth.make_batch(3, 9, 64, 40)
    out = orc.deblock_batch(frames, 64, 40, qps, v, hh)
    for f in range(9):
        assert np.array_equal(out[f], orc.deblock_frame(frames[f], 64, 40, int(qps[f]), v[f], hh[f]))
    assert np.array_equal(out[3], frames[3]) and np.array_equal(out[5], frames[5])  # QP 12, 17: no-op


ref_only = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@ref_only
@pytest.mark.parametrize("size", [(8, 8), (16, 8), (24, 40), (72, 40), (352, 288), (1920, 1080)])
def test_oracle_equals_reference_random(size):
    w, h = size
    rng = np.random.default_rng(w * 7919 + h)
    for kind in ("blocky", "steps", "saturated", "random"):
        for bs_kind in (None, "uniform012", "wide"):
            frame = synth.make_frame(rng, w, h, kind)
            v, hh = (None, None) if bs_kind is None else synth.make_bs(rng, w, h, bs_kind)
            qp = int(rng.integers(16, 56))
            a = orc.deblock_frame(frame, w, h, qp, v, hh)
            b, _ = orc.ref_deblock_frame(frame, w, h, qp, v, hh)
            assert np.array_equal(a, b), (size, kind, bs_kind, qp)


@ref_only
def test_reference_thread_invariance_and_errors():
    rng = np.random.default_rng(5)
    frame = synth.make_frame(rng, 352, 288, "blocky")
    a, _ = orc.ref_deblock_frame(frame, 352, 288, 37, threads=1)
    b, _ = orc.ref_deblock_frame(frame, 352, 288, 37, threads=8)
    assert np.array_equal(a, b)
    with pytest.raises(orc.OracleError, match="Incorrect file size"):
        orc.ref_deblock_frame(frame[:-1], 352, 288, 37)
    with pytest.raises(orc.OracleError, match="multiplier"):
        orc.ref_deblock_frame(np.zeros(3 * 12 * 8 // 2, np.uint8), 12, 8, 37)
    with pytest.raises(orc.OracleError, match="boundary strenght"):
        orc.ref_deblock_frame(frame, 352, 288, 37, np.zeros(3, np.uint8), np.zeros(3, np.uint8))


@ref_only
def test_reference_batch_harness():
    frames, qps, v, hh = synth.make_batch(9, 4, 64, 48)
    a = orc.deblock_batch(frames, 64, 48, qps, v, hh)
    b, _ = orc.ref_deblock_batch(frames, 64, 48, qps, v, hh, outer_threads=2, inner_threads=1)
    c, _ = orc.ref_deblock_batch(frames, 64, 48, qps, v, hh, outer_threads=1, inner_threads=4)
    assert np.array_equal(a, b) and np.array_equal(a, c)
