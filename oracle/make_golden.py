"""Generate tests/golden/ from the UNMODIFIED reference (oracle/_ref), run in this container.

  python -m oracle.make_golden

Writes
  tests/golden/fixtures.npz   the reference's three bundled YUV frames (inputs; /root/reference
                              does not exist on the GPU box) -- data files of the reference repo,
                              MIT-licensed, stored compressed, never its sources;
  tests/golden/golden.json    md5 / sha256 of the reference's output for every case below.
Synthetic inputs are not stored: tests regenerate them from the seed with tests/synth.py.
"""
import hashlib
import json
import os

import numpy as np

from oracle import oracle_py as orc
from tests import synth

REF_DIR = "/root/reference/hevc_deblocking_filter"
OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

BUNDLED = [  # (file, w, h, QPs) -- main.cu:112-133 gives the QP each frame is run with
    ("mother-daughter_352x288_yv12.yuv", 352, 288, [35, 0, 17, 18, 26, 44, 51, 52]),
    ("image1_352x288_yv12.yuv", 352, 288, [30, 22, 40]),
    ("image2_768x576.yuv", 768, 576, [30, 0, 15, 17, 18, 37, 51, 52, 200]),
]

SYNTH = [  # (seed, w, h, kind, bs_kind, qp)
    (11, 8, 8, "blocky", None, 37), (12, 16, 16, "blocky", "uniform012", 37), (13, 24, 40, "steps", None, 30),
    (14, 40, 24, "blocky", "wide", 45), (15, 64, 64, "saturated", "uniform012", 51),
    (16, 352, 288, "blocky", "uniform012", 37), (17, 1920, 1080, "blocky", "uniform012", 37),
    (18, 1920, 1080, "blocky", None, 27), (19, 3840, 2160, "blocky", "uniform012", 32),
    (20, 7680, 4320, "blocky", "uniform012", 41), (21, 96, 1080, "random", "ones", 33),
    (22, 1280, 720, "steps", "twos", 24),
]


def digest(a):
    b = a.tobytes()
    return {"md5": hashlib.md5(b).hexdigest(), "sha256": hashlib.sha256(b).hexdigest()}


def synth_case(seed, w, h, kind, bs_kind):
    rng = np.random.default_rng(seed)
    frame = synth.make_frame(rng, w, h, kind)
    v, hh = (None, None) if bs_kind is None else synth.make_bs(rng, w, h, bs_kind)
    return frame, v, hh


def main():
    os.makedirs(OUT_DIR, exist_ok=True)
    orc.build()
    fixtures, golden = {}, {"bundled": [], "synthetic": []}
    for name, w, h, qps in BUNDLED:
        x = np.fromfile(os.path.join(REF_DIR, name), np.uint8)
        fixtures[name] = x
        for qp in qps:
            y, _ = orc.ref_deblock_frame(x, w, h, qp)
            y8, _ = orc.ref_deblock_frame(x, w, h, qp, threads=8)
            assert np.array_equal(y, y8)
            golden["bundled"].append({"file": name, "w": w, "h": h, "qp": qp, "input_md5": hashlib.md5(x.tobytes()).hexdigest(),
                                      "changed_bytes": int((x != y).sum()), **digest(y)})
    for seed, w, h, kind, bs_kind, qp in SYNTH:
        frame, v, hh = synth_case(seed, w, h, kind, bs_kind)
        y, _ = orc.ref_deblock_frame(frame, w, h, qp, v, hh)
        golden["synthetic"].append({"seed": seed, "w": w, "h": h, "kind": kind, "bs_kind": bs_kind, "qp": qp,
                                    "input_md5": hashlib.md5(frame.tobytes()).hexdigest(),
                                    "changed_bytes": int((frame != y).sum()), **digest(y)})
    np.savez_compressed(os.path.join(OUT_DIR, "fixtures.npz"), **fixtures)
    with open(os.path.join(OUT_DIR, "golden.json"), "w") as f:
        json.dump(golden, f, indent=1)
    print("wrote", OUT_DIR, {k: len(v) for k, v in golden.items()})


if __name__ == "__main__":
    main()
