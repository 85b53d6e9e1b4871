"""Sample adaptive offset (absent from the reference, so unpinned by nature): the C
oracle against the independent numpy statement (oracle/sao_spec.py), and hand-made cases of every category."""
import numpy as np
import pytest

from oracle import oracle_py as orc
from oracle import sao_spec
from tests import synth


def make_sao_params(rng, w, h, ctb=64, kind="mixed"):
    rows, cols = (h + ctb - 1) // ctb, (w + ctb - 1) // ctb
    p = np.zeros((3, rows, cols), orc.SAO_PARAMS)
    if kind == "off":
        return p
    p["type"] = rng.integers(0, 3, p.shape) if kind in ("mixed", "extreme") else {"band": 1, "edge": 2}[kind]
    p["position"] = np.where(p["type"] == 1, rng.integers(0, 32, p.shape), rng.integers(0, 4, p.shape))
    mag = rng.integers(0, 8, p.shape + (4,))
    if kind == "extreme":
        p["type"] = rng.integers(1, 3, p.shape)
        mag = rng.choice(np.array([-128, -100, 100, 127]), p.shape + (4,))
        p["offset"] = mag
        return p
    # edge offsets: positive for the minimum side, negative for the maximum side, as the syntax implies; band: any sign
    sign_eo = np.array([1, 1, -1, -1])
    sign_bo = rng.choice(np.array([-1, 1]), p.shape + (4,))
    p["offset"] = np.where((p["type"] == 2)[..., None], mag * sign_eo, mag * sign_bo)
    return p


@pytest.mark.parametrize("size,ctb", [((64, 64), 64), ((200, 120), 64), ((352, 288), 32), ((352, 288), 16), ((1920, 1080), 64)],
                         ids=lambda v: str(v).replace(" ", ""))
@pytest.mark.parametrize("kind", ["mixed", "band", "edge", "extreme", "off"])
def test_c_oracle_equals_the_independent_statement(size, ctb, kind):
    w, h = size
    rng = np.random.default_rng(w * 3 + h + ctb + len(kind))
    for content in ("blocky", "random", "saturated"):
        frame = synth.make_frame(rng, w, h, content)
        params = make_sao_params(rng, w, h, ctb, kind)
        got = orc.sao_frame(frame, w, h, params, ctb)
        want = sao_spec.sao_frame(frame, w, h, params, ctb)
        assert np.array_equal(got, want), (size, ctb, kind, content, int((got != want).sum()))
        if kind == "off":
            assert np.array_equal(got, frame)
        elif kind in ("mixed", "edge", "band") and content != "saturated":
            assert (got != frame).mean() > (0.0005 if kind == "band" else 0.01)


def test_edge_categories_by_hand():
    """A 3-sample neighbourhood per category, class 0 (horizontal), in the middle of a 16x8 plane."""
    w, h = 16, 8
    params = np.zeros((3, 1, 1), orc.SAO_PARAMS)
    params["type"][0] = 2
    params["position"][0] = 0
    params["offset"][0] = (3, 2, -2, -3)
    for (a, c, b), want in [((9, 5, 9), 5 + 3), ((5, 5, 9), 5 + 2), ((9, 5, 5), 5 + 2), ((1, 5, 9), 5), ((5, 5, 5), 5),
                            ((5, 9, 9), 9 - 2), ((9, 9, 5), 9 - 2), ((5, 9, 5), 9 - 3)]:
        frame = np.full(3 * w * h // 2, 128, np.uint8)
        y = frame[:w * h].reshape(h, w)
        y[4, 7:10] = (a, c, b)
        out = orc.sao_frame(frame, w, h, params, 16)
        assert out[:w * h].reshape(h, w)[4, 8] == want, ((a, c, b), want)
    # the picture's first and last columns have a neighbour outside: untouched, whatever their values
    frame = np.arange(3 * w * h // 2, dtype=np.uint8) * 37
    out = orc.sao_frame(frame, w, h, params, 16)
    y_in, y_out = frame[:w * h].reshape(h, w), out[:w * h].reshape(h, w)
    assert np.array_equal(y_out[:, 0], y_in[:, 0]) and np.array_equal(y_out[:, -1], y_in[:, -1])


def test_band_offset_by_hand():
    w, h = 16, 8
    params = np.zeros((3, 1, 1), orc.SAO_PARAMS)
    params["type"][0] = 1
    params["position"][0] = 30                   # bands 30, 31, 0, 1: wraps
    params["offset"][0] = (1, 2, 3, -4)
    frame = np.zeros(3 * w * h // 2, np.uint8)
    frame[:8] = (240, 247, 248, 255, 0, 7, 8, 16)     # bands 30 30 31 31 0 0 1 2
    out = orc.sao_frame(frame, w, h, params, 16)
    assert list(out[:8]) == [241, 248, 250, 255, 3, 10, 4, 16]


# --- samples deeper than 8 bits (hdbo_sao_frame16) ---

def make_sao_params16(rng, w, h, ctb, kind, depth):
    """Offsets up to the standard's range at the depth: 7 at eight bits, 31 at ten and above."""
    p = make_sao_params(rng, w, h, ctb, kind)
    if kind in ("mixed", "band", "edge") and depth >= 10:
        scale = rng.integers(1, 5, p.shape + (1,))
        p["offset"] = np.clip(p["offset"].astype(np.int32) * scale, -31, 31)
    return p


@pytest.mark.parametrize("depth", [8, 9, 10, 12])
@pytest.mark.parametrize("size,ctb", [((64, 64), 64), ((200, 120), 32), ((352, 288), 16), ((1920, 1080), 64)], ids=lambda v: str(v).replace(" ", ""))
def test_c_oracle16_equals_the_independent_statement(size, ctb, depth):
    from tests.test_depth import make_frame16
    w, h = size
    rng = np.random.default_rng(w + 5 * h + ctb + depth)
    for kind, content in [("mixed", "blocky"), ("edge", "random"), ("band", "ramp"), ("extreme", "saturated")]:
        frame = make_frame16(rng, w, h, content, depth)
        params = make_sao_params16(rng, w, h, ctb, kind, depth)
        got = orc.sao_frame16(frame, w, h, params, ctb, depth)
        want = sao_spec.sao_frame(frame, w, h, params, ctb, depth)
        assert got.dtype == np.uint16 and int(got.max()) < (1 << depth)
        assert np.array_equal(got, want), (size, ctb, depth, kind, int((got != want).sum()))


def test_depth8_in_16bit_containers_is_the_8bit_stage():
    w, h, ctb = 200, 120, 32
    rng = np.random.default_rng(77)
    for kind, content in [("mixed", "blocky"), ("extreme", "saturated"), ("band", "random")]:
        frame = synth.make_frame(rng, w, h, content)
        params = make_sao_params(rng, w, h, ctb, kind)
        assert np.array_equal(orc.sao_frame16(frame.astype(np.uint16), w, h, params, ctb, 8), orc.sao_frame(frame, w, h, params, ctb))


def test_band_of_a_ten_bit_sample_by_hand():
    w, h = 16, 8
    params = np.zeros((3, 1, 1), orc.SAO_PARAMS)
    params["type"][0] = 1
    params["position"][0] = 31                   # bands 31, 0, 1, 2: wraps
    params["offset"][0] = (31, -31, 5, -5)
    frame = np.full(3 * w * h // 2, 512, np.uint16)
    frame[:8] = (1023, 992, 991, 0, 31, 32, 64, 96)      # bands 31 31 30 0 0 1 2 3
    out = orc.sao_frame16(frame, w, h, params, 16, 10)
    assert list(out[:8]) == [1023, 1023, 991, 0, 0, 37, 59, 96]
