import hashlib
import json
import os

import numpy as np

from tests import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(HERE, "golden")


def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


_fixtures = None


def fixture(name):
    """One of the reference's bundled YUV frames (stored under tests/golden/, see oracle/make_golden.py)."""
    global _fixtures
    if _fixtures is None:
        _fixtures = np.load(os.path.join(GOLDEN_DIR, "fixtures.npz"))
    return np.ascontiguousarray(_fixtures[name])


def md5(a):
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def sha256(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def synth_case(seed, w, h, kind, bs_kind):
    rng = np.random.default_rng(seed)
    frame = synth.make_frame(rng, w, h, kind)
    v, hh = (None, None) if bs_kind is None else synth.make_bs(rng, w, h, bs_kind)
    return frame, v, hh


def expected_path(w):
    """hdb_timing.path for a frame of width w: 1 tiled kernel with TMA loads, 3 tiled kernel with cp.async
    loads (width off the 32 grid), 2 padded-scratch kernels (width below 32)."""
    return 1 if w % 32 == 0 else (3 if w >= 32 else 2)
