"""Bounding the unpinned derivation: oracle/derive_spec.py (an independent, set-based restatement of H.265
8.7.2.4) against oracle/deblock_oracle.c::hdbo_derive_edges, against the product's arithmetic compiled for the
host (tests/emu) and, on a GPU, against DeriveEdgesKernel -- exhaustively over a reduced domain."""
import itertools

import numpy as np
import pytest

from oracle import derive_spec as spec
from oracle import oracle_py as orc
from tests.test_core_emu import emu  # noqa: F401
from tests.test_edges import _emu_derive

BLOCK_INFO = orc.BLOCK_INFO
MV_VALUES = [(0, 0), (3, 0), (4, 0), (0, -4), (32767, 3), (-32768, 0)]
REF_PAIRS = [(a, b) for a in (-1, 0, 1) for b in (-1, 0, 1)]          # three states per list: unused, picture 0, picture 1


def _blk(mv0, mv1, ref, flags, qp=30):
    return {"mv": [list(mv0), list(mv1)], "ref": list(ref), "flags": flags, "qp": qp}


def motion_cases():
    """No intra: transform / prediction edge flags x coded flag x every reference pair on both sides x motion
    vectors at the threshold and at the ends of the 16-bit range (sides share the candidates, so every pairing
    of equal / near / far occurs)."""
    cases = []
    for tu, pu, cbf in itertools.product((0, 1), (0, 1), (0, 1)):
        for rp, rq in itertools.product(REF_PAIRS[1:], REF_PAIRS[1:]):          # at least one list used
            for mvs in itertools.product(MV_VALUES, repeat=4):
                cases.append((_blk(mvs[0], mvs[1], rp, cbf * spec.CBF), _blk(mvs[2], mvs[3], rq, 0), tu, pu))
    return cases


def flag_cases():
    """Every combination of intra / coded on both sides and of the two edge kinds x every reference pair
    (including 'no list used') x three motion configurations."""
    cases = []
    for ip, iq, cp, cq, tu, pu in itertools.product((0, 1), repeat=6):
        for rp, rq in itertools.product(REF_PAIRS, REF_PAIRS):
            for mvs in [((0, 0),) * 4, ((0, 0), (9, 9), (4, 0), (9, 9)), ((1, 1), (2, 2), (2, 2), (1, 1))]:
                cases.append((_blk(mvs[0], mvs[1], rp, ip * spec.INTRA | cp * spec.CBF),
                              _blk(mvs[2], mvs[3], rq, iq * spec.INTRA | cq * spec.CBF), tu, pu))
    return cases


def qp_cases():
    return [(_blk((0, 0), (0, 0), (0, -1), 0, a), _blk((8, 0), (0, 0), (0, -1), 0, b), 1, 1)
            for a in range(0, 64) for b in range(0, 64)] + \
           [(_blk((0, 0), (0, 0), (0, -1), 0, a), _blk((0, 0), (0, 0), (0, -1), 0, 255 - a), 1, 0) for a in range(0, 256, 5)]


def layout(cases, vertical, pairs_per_row=512):
    """The cases as one picture of block data: case i is a pair of neighbouring blocks (p left of q for a vertical
    edge, p above q for a horizontal one) whose shared edge is entry idx[i] of the corresponding edge map."""
    n = len(cases)
    rows = (n + pairs_per_row - 1) // pairs_per_row
    if vertical:
        h8, w8 = rows, 2 * pairs_per_row
    else:
        h8, w8 = 2 * rows, pairs_per_row
    info = np.zeros((h8, w8), BLOCK_INFO)
    info["ref"] = -1
    info["flags"] = spec.INTRA
    idx = np.empty(n, np.int64)
    for i, (p, q, tu, pu) in enumerate(cases):
        r, c = divmod(i, pairs_per_row)
        (pr, pc), (qr, qc) = ((r, 2 * c), (r, 2 * c + 1)) if vertical else ((2 * r, c), (2 * r + 1, c))
        for (rr, cc), b, is_q in (((pr, pc), p, False), ((qr, qc), q, True)):
            info["mv"][rr, cc] = b["mv"]
            info["ref"][rr, cc] = b["ref"]
            fl = b["flags"]
            if is_q:        # the edge kinds are properties of the block below / to the right of the edge
                fl |= (tu * (spec.TU_LEFT if vertical else spec.TU_TOP)) | (pu * (spec.PU_LEFT if vertical else spec.PU_TOP))
            info["flags"][rr, cc] = fl
            info["qp"][rr, cc] = b["qp"]
        idx[i] = qr * (w8 + 1) + qc if vertical else qr * w8 + qc
    return info, 8 * w8, 8 * h8, idx


SETS = {"motion": motion_cases, "flags": flag_cases, "qp": qp_cases}


@pytest.fixture(scope="module")
def expected():
    out = {}
    for name, make in SETS.items():
        cases = make()
        out[name] = (cases, np.array([spec.edge_byte(*c) for c in cases], np.uint8))
    return out


@pytest.mark.parametrize("name", sorted(SETS))
@pytest.mark.parametrize("vertical", [True, False], ids=["vertical-edges", "horizontal-edges"])
def test_c_oracle_equals_the_independent_statement(expected, name, vertical):
    cases, want = expected[name]
    info, w, h, idx = layout(cases, vertical)
    v, hh = orc.derive_edges(info, w, h)
    got = (v if vertical else hh)[idx]
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, (name, len(cases), [(cases[i], int(got[i]), int(want[i])) for i in bad[:3]])
    # the domain is not degenerate: all three strengths occur (motion: never 2)
    assert set(np.unique(want & 3)) >= ({0, 1} if name != "flags" else {0, 1, 2})


@pytest.mark.parametrize("name", sorted(SETS))
@pytest.mark.parametrize("vertical", [True, False], ids=["vertical-edges", "horizontal-edges"])
def test_product_arithmetic_on_the_host_equals_the_independent_statement(emu, expected, name, vertical):  # noqa: F811
    cases, want = expected[name]
    info, w, h, idx = layout(cases, vertical)
    v, hh = _emu_derive(emu, info, w, h)
    got = (v if vertical else hh)[idx]
    assert np.array_equal(got, want), int((got != want).sum())


def test_domain_size():
    assert len(motion_cases()) == 8 * 64 * 6 ** 4 and len(flag_cases()) == 64 * 81 * 3


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SETS))
@pytest.mark.parametrize("vertical", [True, False], ids=["vertical-edges", "horizontal-edges"])
def test_kernel_equals_the_independent_statement(expected, name, vertical):
    import gpu_video_codec_b200 as hdb
    cases, want = expected[name]
    info, w, h, idx = layout(cases, vertical)
    dev = hdb.Deblocker(0)
    v, hh = dev.derive_edges(info[None], w, h)
    got = (v[0] if vertical else hh[0])[idx]
    assert np.array_equal(got, want), int((got != want).sum())
    dev.close()
