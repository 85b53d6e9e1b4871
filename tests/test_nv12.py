"""NV12 on the CPU: the oracle's byte permutations, and the product's interleaved-chroma block code
(SplitUv / ChromaBlock / JoinUv of deblock_core.cuh, compiled for the host) against the oracle."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle_py as orc
from tests import common, synth
from tests.test_core_emu import emu  # noqa: F401

_u8p = C.POINTER(C.c_uint8)


def test_permutations_round_trip():
    rng = np.random.default_rng(0)
    for (w, h) in [(8, 8), (64, 48), (352, 288)]:
        f = rng.integers(0, 256, 3 * w * h // 2, dtype=np.uint8)
        nv = orc.yv12_to_nv12(f, w, h)
        assert np.array_equal(orc.nv12_to_yv12(nv, w, h), f)
        assert np.array_equal(nv[:w * h], f[:w * h])
        assert nv[w * h] == f[w * h] and nv[w * h + 1] == f[w * h + (w // 2) * (h // 2)]


def test_nv12_oracle_on_the_reference_frames():
    # the reference's own frames stored as NV12 give the golden output stored as NV12
    for case in common.golden()["bundled"][:3]:
        x = common.fixture(case["file"])
        y = orc.deblock_frame_nv12(orc.yv12_to_nv12(x, case["w"], case["h"]), case["w"], case["h"], case["qp"])
        assert common.md5(orc.nv12_to_yv12(y, case["w"], case["h"])) == case["md5"]


# widths that are multiples of 16: otherwise the reference addresses its chroma planes with a pitch that is not
# their width (cpu.h:469-471) and rows skew; the product takes NV12 at multiples of 32 only
@pytest.mark.parametrize("size", [(16, 8), (32, 40), (64, 48), (352, 288), (208, 120), (1920, 1080)])
def test_emu_interleaved_chroma_matches_oracle(emu, size):  # noqa: F811
    w, h = size
    emu.emu_deblock_uv_plane.argtypes = [_u8p, _u8p, C.c_uint, C.c_uint, C.c_uint]
    rng = np.random.default_rng(w * 13 + h)
    for kind in ("blocky", "steps", "saturated", "random"):
        for qp in (17, 30, 44, 51):
            nv = orc.yv12_to_nv12(synth.make_frame(rng, w, h, kind), w, h)
            want = orc.deblock_frame_nv12(nv, w, h, qp)
            uv = np.ascontiguousarray(nv[w * h:])
            out = np.empty_like(uv)
            assert emu.emu_deblock_uv_plane(uv.ctypes.data_as(_u8p), out.ctypes.data_as(_u8p), w, h, qp) == 0
            assert np.array_equal(out, want[w * h:]), (size, kind, qp)
