"""Edge maps on the CPU: the per-edge-QP restatement against golden
hashes that came out of the UNMODIFIED reference (oracle/make_golden_edges.py), against chains of
reference runs made here where oracle/_ref is built, and the product's arithmetic header -- edge-QP
luma blocks and the boundary-strength derivation -- compiled for the host against the restatement."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import make_golden_edges as mge
from oracle import oracle_py as orc
from tests import common, synth
from tests.test_core_emu import emu  # noqa: F401  (fixture: builds tests/emu/libemu_deblock.so)

_u8p = C.POINTER(C.c_uint8)
ref_only = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built (needs /root/reference)")

with open(os.path.join(common.GOLDEN_DIR, "edges.json")) as _f:
    GOLDEN_EDGES = json.load(_f)["edges"]


@pytest.mark.parametrize("case", GOLDEN_EDGES, ids=lambda c: "s%d-%dx%d" % (c["seed"], c["w"], c["h"]))
def test_oracle_edges_golden(case):
    frame, v, hh = mge.case_inputs(case["seed"], case["w"], case["h"], case["kind"], case["edge_kind"])
    assert hashlib.md5(frame.tobytes() + v.tobytes() + hh.tobytes()).hexdigest() == case["input_md5"], \
        "tests/synth.py changed: regenerate tests/golden/edges.json"
    y = orc.deblock_frame_edges(frame, case["w"], case["h"], case["qp"], v, hh)
    assert common.md5(y) == case["md5"] and common.sha256(y) == case["sha256"]
    assert int((y != frame).sum()) == case["changed_bytes"]


def test_uniform_edge_maps_equal_plain_filter():
    rng = np.random.default_rng(2)
    for (w, h) in [(8, 8), (64, 48), (352, 288), (200, 120)]:
        for qp in (0, 17, 18, 30, 44, 51, 52, 63):
            frame = synth.make_frame(rng, w, h, "blocky")
            v, hh = synth.make_bs(rng, w, h, "uniform012")
            a = orc.deblock_frame_edges(frame, w, h, qp, synth.edge_bytes(v, qp), synth.edge_bytes(hh, qp))
            assert np.array_equal(a, orc.deblock_frame(frame, w, h, min(qp, 51), v, hh))


def test_edge_errors_and_off_entries():
    w, h = 64, 48
    rng = np.random.default_rng(4)
    frame = synth.make_frame(rng, w, h, "blocky")
    v, hh = synth.make_edge_maps(rng, w, h, "random")
    with pytest.raises(orc.OracleError):
        orc.deblock_frame_edges(frame, w, h, 30, v[:-1], hh)     # cpu.h:122-123
    # strength 0 switches an entry off whatever its qp bits; qp <= 17 is a no-op whatever the strength
    y = orc.deblock_frame_edges(frame, w, h, 10, v & 0xfc, hh & 0xfc)
    assert np.array_equal(y, frame)
    y = orc.deblock_frame_edges(frame, w, h, 10, (v & 3) | (17 << 2), (hh & 3) | (12 << 2))
    assert np.array_equal(y, frame)


@ref_only
@pytest.mark.parametrize("size", [(8, 8), (16, 8), (24, 40), (64, 48), (352, 288)])
def test_oracle_edges_equal_reference_chain(size):
    w, h = size
    rng = np.random.default_rng(w * 31 + h)
    for kind in ("blocky", "steps", "saturated", "random"):
        for edge_kind in ("random", "wide", "smooth"):
            frame = synth.make_frame(rng, w, h, kind)
            v, hh = synth.make_edge_maps(rng, w, h, edge_kind)
            qp = int(rng.integers(16, 56))
            a = orc.deblock_frame_edges(frame, w, h, qp, v, hh)
            b = orc.ref_compose_edges(frame, w, h, qp, v, hh)
            assert np.array_equal(a, b), (size, kind, edge_kind, qp)


@ref_only
def test_oracle_edges_uniform_with_picture_edges_equal_reference():
    # picture edges ON (the reference's own defaults have them on): the padding-dependent paths, one QP
    rng = np.random.default_rng(8)
    for (w, h) in [(64, 48), (352, 288), (1920, 1080)]:
        frame = synth.make_frame(rng, w, h, "blocky")
        v, hh = synth.make_bs(rng, w, h, "twos")
        for qp in (27, 44):
            a = orc.deblock_frame_edges(frame, w, h, qp, synth.edge_bytes(v, qp), synth.edge_bytes(hh, qp))
            b, _ = orc.ref_deblock_frame(frame, w, h, qp, v, hh)
            assert np.array_equal(a, b)


# ---- the product's arithmetic header on the host ----------------------------------------------
def _emu_edges(lib, frame, w, h, qp, v, hh):
    lib.emu_deblock_frame_edges.argtypes = [_u8p, _u8p, C.c_uint, C.c_uint, C.c_uint, _u8p, _u8p]
    out = np.empty_like(frame)
    p = lambda a: a.ctypes.data_as(_u8p)
    assert lib.emu_deblock_frame_edges(p(frame), p(out), w, h, qp, p(v), p(hh)) == 0
    return out


@pytest.mark.parametrize("size", [(8, 8), (24, 40), (64, 48), (352, 288), (200, 120), (1920, 1080)])
def test_emu_edges_match_oracle(emu, size):  # noqa: F811
    w, h = size
    rng = np.random.default_rng(w * 17 + h)
    for kind in (("blocky",) if w > 1000 else ("blocky", "steps", "saturated", "random")):
        for edge_kind, pic in (("random", False), ("wide", True), ("smooth", True)):
            frame = synth.make_frame(rng, w, h, kind)
            v, hh = synth.make_edge_maps(rng, w, h, edge_kind, picture_edges=pic)
            got = _emu_edges(emu, frame, w, h, 33, v, hh)
            want = orc.deblock_frame_edges(frame, w, h, 33, v, hh)
            assert np.array_equal(got, want), (size, kind, edge_kind, int((got != want).sum()))


def _emu_derive(lib, info, w, h):
    lib.emu_derive_edges.argtypes = [C.c_void_p, C.c_uint, C.c_uint, _u8p, _u8p]
    lib.emu_derive_edges.restype = None
    v = np.empty(orc.num_vert_bs(w, h), np.uint8)
    hh = np.empty(orc.num_hor_bs(w, h), np.uint8)
    lib.emu_derive_edges(info.ctypes.data, w, h, v.ctypes.data_as(_u8p), hh.ctypes.data_as(_u8p))
    return v, hh


@pytest.mark.parametrize("size", [(8, 8), (16, 24), (64, 48), (352, 288), (1920, 1080)])
@pytest.mark.parametrize("kind", ["intra", "random", "cu16", "extreme"])
def test_emu_derive_matches_oracle(emu, size, kind):  # noqa: F811
    w, h = size
    rng = np.random.default_rng(w + 7 * h)
    info = synth.make_block_info(rng, w, h, kind)
    assert info.dtype.itemsize == 12
    v, hh = orc.derive_edges(info, w, h)
    v2, h2 = _emu_derive(emu, info, w, h)
    assert np.array_equal(v, v2) and np.array_equal(hh, h2)


def test_derive_rule_cases():
    """Hand-made neighbours: every branch of the rule (H.265 8.7.2.4)."""
    w, h = 16, 8     # two blocks side by side: vert entry [0][1] is the edge between them

    def bs(p, q, flags_q=synth.BLK_TU_LEFT | synth.BLK_PU_LEFT):
        info = np.zeros((1, 2), synth.BLOCK_INFO)
        for i, b in enumerate((p, q)):
            info["mv"][0, i] = b.get("mv", [[0, 0], [0, 0]])
            info["ref"][0, i] = b.get("ref", [0, -1])
            info["flags"][0, i] = b.get("flags", 0)
            info["qp"][0, i] = b.get("qp", 30)
        info["flags"][0, 1] |= flags_q
        v, hh = orc.derive_edges(info, w, h)
        assert v[0] == 0 and v[2] == 0 and not hh.any()     # picture sides off
        return int(v[1] & 3), int(v[1] >> 2)

    assert bs({}, {}) == (0, 30)                                                   # same motion
    assert bs({"flags": synth.BLK_INTRA}, {}) == (2, 30)
    assert bs({}, {"flags": synth.BLK_INTRA}, flags_q=0) == (0, 30)                # not an edge at all
    assert bs({"flags": synth.BLK_CBF}, {}) == (1, 30)
    assert bs({"flags": synth.BLK_CBF}, {}, flags_q=synth.BLK_PU_LEFT) == (0, 30)  # coefficients count on transform edges only
    assert bs({"ref": [0, -1]}, {"ref": [1, -1]}) == (1, 30)                       # different pictures
    assert bs({"ref": [0, -1]}, {"ref": [-1, 0]}) == (0, 30)                       # same picture through the other list
    assert bs({"ref": [0, -1]}, {"ref": [0, 1]}) == (1, 30)                        # different vector count
    assert bs({"mv": [[0, 0], [0, 0]]}, {"mv": [[3, -3], [0, 0]]}) == (0, 30)
    assert bs({"mv": [[0, 0], [0, 0]]}, {"mv": [[4, 0], [0, 0]]}) == (1, 30)
    assert bs({"mv": [[0, 0], [0, 0]]}, {"mv": [[0, -4], [0, 0]]}) == (1, 30)
    two = {"ref": [0, 1], "mv": [[1, 1], [8, 8]]}
    assert bs(two, {"ref": [1, 0], "mv": [[8, 8], [1, 1]]}) == (0, 30)             # lists swapped, same motion
    assert bs(two, {"ref": [1, 0], "mv": [[1, 1], [8, 8]]}) == (1, 30)
    same = {"ref": [2, 2], "mv": [[0, 0], [10, 10]]}
    assert bs(same, {"ref": [2, 2], "mv": [[10, 10], [0, 0]]}) == (0, 30)          # both pairings tried
    assert bs(same, {"ref": [2, 2], "mv": [[10, 10], [5, 0]]}) == (1, 30)
    assert bs({"qp": 20}, {"qp": 27}) == (0, 24) and bs({"qp": 60}, {"qp": 61}) == (0, 51)


def test_all_intra_gives_reference_default_inside():
    """All-intra blocks, every edge a transform edge: strength 2 on every inner edge at the blocks' QP --
    the reference's constant "2" (cpu.h:91-99) away from the picture's sides."""
    w, h = 64, 48
    rng = np.random.default_rng(1)
    info = synth.make_block_info(rng, w, h, "intra")
    q = int(info["qp"][0, 0])
    v, hh = orc.derive_edges(info, w, h)
    v2, h2 = v.reshape(h // 8, w // 8 + 1), hh.reshape(h // 8 + 1, w // 8)
    assert (v2[:, 1:-1] == synth.edge_bytes(2, q)).all() and (h2[1:-1] == synth.edge_bytes(2, q)).all()
    assert not v2[:, 0].any() and not v2[:, -1].any() and not h2[0].any() and not h2[-1].any()
    frame = synth.make_frame(rng, w, h, "blocky")
    dv, dh = orc.default_bs(w, h)
    dv2, dh2 = dv.reshape(v2.shape).copy(), dh.reshape(h2.shape).copy()
    dv2[:, 0] = dv2[:, -1] = 0
    dh2[:] = 2
    dh2[0] = dh2[-1] = 0
    a = orc.deblock_frame_edges(frame, w, h, q, v, hh)
    b = orc.deblock_frame(frame, w, h, q, dv2.reshape(-1), dh2.reshape(-1))
    ysz = w * h
    assert np.array_equal(a[:ysz], b[:ysz])


def test_bench_block_info_generator_packs_the_abi_layout():
    """bench.device_block_info builds hdb_block_info records as three 32-bit words on the device; on the CPU the
    same code must give records the oracle reads as sane block data and derives a mix of strengths from."""
    import torch

    import bench
    n, w, h = 2, 352, 288
    words = bench.device_block_info(torch, n, w, h, seed=1, device="cpu")
    assert words.shape == (n, (h // 8) * (w // 8), 3) and words.dtype == torch.int32
    info = words.numpy().view(np.uint8).reshape(n, h // 8, w // 8, 12).copy().view(orc.BLOCK_INFO).reshape(n, h // 8, w // 8)
    assert info["qp"].min() >= 0 and info["qp"].max() <= 51
    assert set(np.unique(info["ref"])) <= {-1, 0, 1, 2}
    assert np.abs(info["mv"].astype(np.int64)).max() <= 40
    intra = (info["flags"] & synth.BLK_INTRA) != 0
    assert (info["ref"][intra] == -1).all()
    v, hh = orc.derive_edges(info[0], w, h)
    hist = np.bincount(np.concatenate([v, hh]) & 3, minlength=4)
    assert hist[0] > 0 and hist[1] > 0 and hist[2] > 0 and hist[3] == 0
