"""Samples deeper than 8 bits.  The reference is 8-bit only, so:
  - the 16-bit oracle at depth 8 must BE the pinned 8-bit oracle (and the unmodified reference) on widened samples;
  - depth > 8 is unpinned against the reference: it is defined by oracle/deblock_oracle.h::hdbo_deblock_frame16
    (H.265's beta / tc scaling, sample maximum (1 << depth) - 1), anchored on identities that hold at every depth;
  - the product's 16-bit arithmetic (csrc/deblock_core16.cuh), compiled for the host, must equal that oracle."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle_py as orc
from tests import synth
from tests.test_core_emu import emu  # noqa: F401

_u8p = C.POINTER(C.c_uint8)


def make_frame16(rng, w, h, kind, depth):
    """tests/synth.py content scaled to the depth, with the low bits filled so that every value of the range occurs."""
    f8 = synth.make_frame(rng, w, h, kind).astype(np.uint16)
    shift = depth - 8
    if shift == 0:
        return f8
    low = rng.integers(0, 1 << shift, f8.shape).astype(np.uint16)
    f = (f8 << shift) | low
    if kind == "saturated":
        f = np.where(f8 == 255, (1 << depth) - 1, np.where(f8 == 0, 0, f)).astype(np.uint16)
    return f


def run_emu16(lib, frame, w, h, qp, v, hh, depth, edges=False):
    lib.emu_deblock_frame16.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_uint, _u8p, _u8p, C.c_uint, C.c_int]
    out = np.empty_like(frame)
    p = lambda a: None if a is None else a.ctypes.data_as(_u8p)
    assert lib.emu_deblock_frame16(frame.ctypes.data, out.ctypes.data, w, h, qp, p(v), p(hh), depth, int(edges)) == 0
    return out


@pytest.mark.parametrize("size", [(16, 16), (64, 48), (352, 288), (200, 120), (1920, 1080)], ids=lambda s: "%dx%d" % s)
def test_depth8_in_16bit_containers_is_the_reference_filter(size):
    w, h = size
    rng = np.random.default_rng(w + h)
    for kind, bs_kind, qp in [("blocky", "uniform012", 37), ("saturated", "wide", 51), ("random", None, 44), ("steps", "twos", 27)]:
        frame = synth.make_frame(rng, w, h, kind)
        v, hh = (None, None) if bs_kind is None else synth.make_bs(rng, w, h, bs_kind)
        want = orc.deblock_frame(frame, w, h, qp, v, hh)
        got = orc.deblock_frame16(frame.astype(np.uint16), w, h, qp, v, hh, bit_depth=8)
        assert np.array_equal(got, want.astype(np.uint16))
        if orc.have_ref() and w * h <= 352 * 288:
            assert np.array_equal(got, orc.ref_deblock_frame(frame, w, h, qp, v, hh)[0].astype(np.uint16))
    # per-edge QP at depth 8
    frame = synth.make_frame(rng, w, h, "blocky")
    v, hh = synth.make_edge_maps(rng, w, h, "random", picture_edges=True)
    assert np.array_equal(orc.deblock_frame16(frame.astype(np.uint16), w, h, 30, v, hh, bit_depth=8, edges=True),
                          orc.deblock_frame_edges(frame, w, h, 30, v, hh).astype(np.uint16))


@pytest.mark.parametrize("depth", [9, 10, 12])
def test_identities_at_every_depth(depth):
    rng = np.random.default_rng(depth)
    w, h = 128, 72
    frame = make_frame16(rng, w, h, "blocky", depth)
    v, hh = synth.make_bs(rng, w, h, "uniform012")
    assert int(frame.max()) > 255
    # QP <= 17: tc = 0 and beta < 8 at every depth scale -> nothing changes
    for qp in (0, 15):
        assert np.array_equal(orc.deblock_frame16(frame, w, h, qp, v, hh, bit_depth=depth), frame)
    # all luma strengths 0: luma untouched
    z = orc.deblock_frame16(frame, w, h, 40, np.zeros_like(v), np.zeros_like(hh), bit_depth=depth)
    assert np.array_equal(z[:w * h], frame[:w * h])
    # a flat frame is a fixed point away from the picture border (the border filters against the zero padding)
    flat = np.full(3 * w * h // 2, 600 if depth > 9 else 300, np.uint16)
    y = orc.deblock_frame16(flat, w, h, 45, bit_depth=depth).astype(np.int64)
    assert np.array_equal(y[:w * h].reshape(h, w)[8:-8, 8:-8], flat[:w * h].reshape(h, w)[8:-8, 8:-8])
    # results stay inside the range and the filter does act
    y = orc.deblock_frame16(frame, w, h, 45, v, hh, bit_depth=depth)
    assert int(y.max()) <= (1 << depth) - 1 and (y != frame).mean() > 0.02


@pytest.mark.parametrize("depth", [8, 9, 10])
@pytest.mark.parametrize("kind", ["blocky", "random", "flat", "ramp", "saturated", "steps"])
def test_emu16_matches_oracle16(emu, kind, depth):  # noqa: F811
    for (w, h) in [(8, 8), (24, 40), (64, 48), (352, 288)]:
        rng = np.random.default_rng(hash((kind, w, h, depth)) & 0xffff)
        for qp in (0, 17, 18, 26, 30, 37, 44, 51, 60):
            frame = make_frame16(rng, w, h, kind, depth)
            for bs_kind in (None, "uniform012", "wide"):
                v, hh = (None, None) if bs_kind is None else synth.make_bs(rng, w, h, bs_kind)
                got = run_emu16(emu, frame, w, h, qp, v, hh, depth)
                want = orc.deblock_frame16(frame, w, h, qp, v, hh, bit_depth=depth)
                assert np.array_equal(got, want), (kind, w, h, qp, bs_kind, depth, int((got != want).sum()))


@pytest.mark.parametrize("depth", [8, 10])
def test_emu16_edge_maps_and_extremes(emu, depth):  # noqa: F811
    rng = np.random.default_rng(33 + depth)
    w, h = 200, 120
    for kind, edge_kind in [("blocky", "random"), ("saturated", "wide"), ("random", "smooth")]:
        frame = make_frame16(rng, w, h, kind, depth)
        v, hh = synth.make_edge_maps(rng, w, h, edge_kind, picture_edges=True)
        got = run_emu16(emu, frame, w, h, 33, v, hh, depth, edges=True)
        assert np.array_equal(got, orc.deblock_frame16(frame, w, h, 33, v, hh, bit_depth=depth, edges=True))
    # every sample at an end of the range, in random order: the clips and the biases at their limits
    top = (1 << depth) - 1
    frame = rng.choice(np.array([0, 1, top - 1, top], np.uint16), 3 * w * h // 2)
    v, hh = synth.make_bs(rng, w, h, "twos")
    for qp in (30, 51):
        assert np.array_equal(run_emu16(emu, frame, w, h, qp, v, hh, depth), orc.deblock_frame16(frame, w, h, qp, v, hh, bit_depth=depth))


def test_emu16_1080p(emu):  # noqa: F811
    rng = np.random.default_rng(8)
    w, h = 1920, 1080
    frame = make_frame16(rng, w, h, "blocky", 10)
    v, hh = synth.make_bs(rng, w, h, "uniform012")
    got = run_emu16(emu, frame, w, h, 37, v, hh, 10)
    want = orc.deblock_frame16(frame, w, h, 37, v, hh, bit_depth=10)
    assert np.array_equal(got, want) and (got != frame).mean() > 0.05


def test_p010_layout_conversion_round_trip():
    rng = np.random.default_rng(1)
    w, h = 64, 48
    planar = make_frame16(rng, w, h, "random", 10)
    p010 = orc.planar_to_p010(planar, w, h)
    assert int(p010.max()) > 1023 and not (p010 & 63).any()
    assert np.array_equal(orc.p010_to_planar(p010, w, h), planar)
    # U0 V0 U1 V1 ... after the luma plane
    cw = w // 2
    assert p010[w * h] == planar[w * h] << 6 and p010[w * h + 1] == planar[w * h + cw * (h // 2)] << 6
