"""The product's arithmetic header, compiled for the host with emulated packed primitives,
against the oracle.  Proves the SIMD-in-register formulation of deblock_core.cuh bit-exact
on a machine without a GPU (the GPU parity tests then only have to prove that the device
instructions behave like their emulation and that the data movement is right)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_py as orc
from tests import synth

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu_deblock.cpp")
LIB = os.path.join(HERE, "emu", "libemu_deblock.so")
CORE = os.path.join(os.path.dirname(HERE), "gpu_video_codec_b200", "csrc", "deblock_core.cuh")
_u8p = C.POINTER(C.c_uint8)


@pytest.fixture(scope="module")
def emu():
    if (not os.path.exists(LIB)) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(CORE)):
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-x", "c++",
                        "-o", LIB, SRC], check=True)
    lib = C.CDLL(LIB)
    lib.emu_deblock_frame.argtypes = [_u8p, _u8p, C.c_uint, C.c_uint, C.c_uint, _u8p, _u8p]
    lib.emu_beta.argtypes = [C.c_uint]
    lib.emu_tc.argtypes = [C.c_uint]
    return lib


def run_emu(lib, frame, w, h, qp, v=None, hh=None):
    out = np.empty_like(frame)
    p = lambda a: None if a is None else a.ctypes.data_as(_u8p)
    assert lib.emu_deblock_frame(p(frame), p(out), w, h, qp, p(v), p(hh)) == 0
    return out


def test_tables(emu):
    for qp in list(range(0, 70)) + [200, 2 ** 31]:
        assert emu.emu_beta(qp) == orc.port().hdbo_beta(qp)
        assert emu.emu_tc(qp) == orc.port().hdbo_tc(qp)


@pytest.mark.parametrize("kind", ["blocky", "random", "flat", "ramp", "saturated", "steps"])
@pytest.mark.parametrize("size", [(8, 8), (16, 16), (24, 40), (64, 48), (352, 288), (200, 120)])
def test_emu_matches_oracle(emu, kind, size):
    w, h = size
    rng = np.random.default_rng(hash((kind, w, h)) & 0xffff)
    for qp in (0, 17, 18, 26, 30, 35, 37, 44, 51, 52):
        frame = synth.make_frame(rng, w, h, kind)
        for bs_kind in (None, "uniform012", "wide"):
            v, hh = (None, None) if bs_kind is None else synth.make_bs(rng, w, h, bs_kind)
            got = run_emu(emu, frame, w, h, qp, v, hh)
            want = orc.deblock_frame(frame, w, h, qp, v, hh)
            assert np.array_equal(got, want), (kind, size, qp, bs_kind, int((got != want).sum()))


def test_emu_1080p(emu):
    rng = np.random.default_rng(7)
    w, h = 1920, 1080
    frame = synth.make_frame(rng, w, h, "blocky")
    v, hh = synth.make_bs(rng, w, h, "uniform012")
    for qp in (27, 37, 51):
        got = run_emu(emu, frame, w, h, qp, v, hh)
        want = orc.deblock_frame(frame, w, h, qp, v, hh)
        assert np.array_equal(got, want)
        assert (got != frame).mean() > 0.05
