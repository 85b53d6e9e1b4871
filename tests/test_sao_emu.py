"""The product's sample-adaptive-offset arithmetic (gpu_video_codec_b200/csrc/sao_core.cuh), compiled for the host with
emulated packed primitives and run over whole planes the way the kernels do (tests/emu/emu_sao.cpp), against the oracle.
Proves the one-comparison-per-sample formulation, the byte-permute offset pick and the band folding bit-exact on a machine
without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_py as orc
from tests import synth
from tests.test_depth import make_frame16
from tests.test_sao import make_sao_params, make_sao_params16

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu_sao.cpp")
LIB = os.path.join(HERE, "emu", "libemu_sao.so")
CORE = [os.path.join(os.path.dirname(HERE), "gpu_video_codec_b200", "csrc", f) for f in ("sao_core.cuh", "deblock_core.cuh")]


@pytest.fixture(scope="module")
def emu():
    if (not os.path.exists(LIB)) or os.path.getmtime(LIB) < max([os.path.getmtime(SRC)] + [os.path.getmtime(c) for c in CORE]):
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-x", "c++", "-o", LIB, SRC], check=True)
    lib = C.CDLL(LIB)
    lib.emu_sao_plane8.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    lib.emu_sao_plane8.restype = None
    lib.emu_sao_plane16.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.emu_sao_plane16.restype = None
    return lib


def emu_frame(lib, frame, w, h, params, ctb, depth=None):
    """The three planes of a planar 4:2:0 frame through the emulated core."""
    out = np.empty_like(frame)
    cols = (w + ctb - 1) // ctb
    off = 0
    for plane, (pw, ph, pctb) in enumerate([(w, h, ctb), (w // 2, h // 2, ctb // 2), (w // 2, h // 2, ctb // 2)]):
        src = np.ascontiguousarray(frame[off:off + pw * ph])
        dst = np.empty_like(src)
        par = np.ascontiguousarray(params[plane])
        if depth is None:
            lib.emu_sao_plane8(src.ctypes.data, dst.ctypes.data, pw, ph, par.ctypes.data, cols, pctb)
        else:
            lib.emu_sao_plane16(src.ctypes.data, dst.ctypes.data, pw, ph, par.ctypes.data, cols, pctb, depth)
        out[off:off + pw * ph] = dst
        off += pw * ph
    return out


@pytest.mark.parametrize("size,ctb", [((8, 8), 16), ((64, 64), 64), ((72, 40), 16), ((200, 120), 32), ((352, 288), 64), ((1000, 64), 32)],
                         ids=lambda v: str(v).replace(" ", ""))
def test_emulated_core_equals_oracle(emu, size, ctb):
    w, h = size
    rng = np.random.default_rng(w + 3 * h + ctb)
    for kind, content in [("mixed", "blocky"), ("edge", "random"), ("band", "ramp"), ("extreme", "saturated"), ("off", "steps")]:
        frame = synth.make_frame(rng, w, h, content)
        params = make_sao_params(rng, w, h, ctb, kind)
        got = emu_frame(emu, frame, w, h, params, ctb)
        want = orc.sao_frame(frame, w, h, params, ctb)
        assert np.array_equal(got, want), (size, ctb, kind, content, int((got != want).sum()))


@pytest.mark.parametrize("depth", [8, 10, 12])
@pytest.mark.parametrize("size,ctb", [((8, 8), 16), ((72, 40), 16), ((200, 120), 32), ((352, 288), 64)], ids=lambda v: str(v).replace(" ", ""))
def test_emulated_core16_equals_oracle(emu, size, ctb, depth):
    w, h = size
    rng = np.random.default_rng(w + 3 * h + ctb + depth)
    for kind, content in [("mixed", "blocky"), ("edge", "random"), ("band", "ramp"), ("extreme", "saturated")]:
        frame = make_frame16(rng, w, h, content, depth)
        params = make_sao_params16(rng, w, h, ctb, kind, depth)
        got = emu_frame(emu, frame, w, h, params, ctb, depth)
        want = orc.sao_frame16(frame, w, h, params, ctb, depth)
        assert np.array_equal(got, want), (size, ctb, depth, kind, content, int((got != want).sum()))


@pytest.mark.parametrize("cls", [0, 1, 2, 3])
def test_every_class_on_every_sample_of_a_small_plane(emu, cls):
    """All four picture borders of each class, both sample sizes."""
    w, h, ctb = 40, 24, 16
    rng = np.random.default_rng(cls)
    params = np.zeros((3, (h + ctb - 1) // ctb, (w + ctb - 1) // ctb), orc.SAO_PARAMS)
    params["type"] = 2
    params["position"] = cls
    params["offset"] = (7, 2, -2, -7)
    f8 = synth.make_frame(rng, w, h, "random")
    assert np.array_equal(emu_frame(emu, f8, w, h, params, ctb), orc.sao_frame(f8, w, h, params, ctb))
    f10 = make_frame16(rng, w, h, "random", 10)
    assert np.array_equal(emu_frame(emu, f10, w, h, params, ctb, 10), orc.sao_frame16(f10, w, h, params, ctb, 10))
