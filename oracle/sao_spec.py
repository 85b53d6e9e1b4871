"""TEST INFRASTRUCTURE ONLY -- a second, independently written statement of sample adaptive offset.

The reference has no SAO (its in-loop chain ends with deblocking), so oracle/deblock_oracle.c::hdbo_sao_frame is a
restatement of H.265 clause 8.7.3 with nothing in the reference to pin it on.  This file states the same process a
second time in whole-plane numpy terms (shifted planes instead of a per-sample loop, a category array instead of an
if chain), sharing no code with the C oracle; tests/test_sao.py demands agreement.  Status stays "parity unpinned".
"""
import numpy as np

# SaoEoClass -> ((dx, dy) of neighbour a, (dx, dy) of neighbour b): 8.7.3.2, table of hPos / vPos
EO_NEIGHBOURS = {0: ((-1, 0), (1, 0)), 1: ((0, -1), (0, 1)), 2: ((-1, -1), (1, 1)), 3: ((1, -1), (-1, 1))}


def _shifted(plane, dx, dy):
    """plane sampled at (x + dx, y + dy), and the mask of positions where that lies inside the picture."""
    h, w = plane.shape
    out = np.zeros_like(plane)
    ok = np.zeros(plane.shape, bool)
    ys, yd = (slice(dy, h), slice(0, h - dy)) if dy >= 0 else (slice(0, h + dy), slice(-dy, h))
    xs, xd = (slice(dx, w), slice(0, w - dx)) if dx >= 0 else (slice(0, w + dx), slice(-dx, w))
    out[yd, xd] = plane[ys, xs]
    ok[yd, xd] = True
    return out, ok


def sao_plane(plane, params, ctb, bit_depth=8):
    """plane: uint8 / uint16 [h, w]; params: structured [rows, cols] with type / position / offset; ctb: block size in
    this plane; bit_depth: of the samples (band shift bit_depth - 5, clip to 2 ** bit_depth - 1)."""
    h, w = plane.shape
    p = np.repeat(np.repeat(params, ctb, axis=0), ctb, axis=1)[:h, :w]       # per-sample parameters
    c = plane.astype(np.int32)
    delta = np.zeros_like(c)
    # band offset: the four bands from band_position on (mod 32) carry the four offsets
    k = ((c >> (bit_depth - 5)) - p["position"].astype(np.int32)) % 32
    band = (p["type"] == 1) & (k < 4)
    delta = np.where(band, np.take_along_axis(p["offset"].astype(np.int32), np.minimum(k, 3)[..., None], axis=2)[..., 0], delta)
    # edge offset, class by class
    for cls, (na, nb) in EO_NEIGHBOURS.items():
        a, ok_a = _shifted(c, *na)
        b, ok_b = _shifted(c, *nb)
        cat = np.sign(c - a) + np.sign(c - b)                                  # -2 .. 2
        idx = np.select([cat == -2, cat == -1, cat == 1, cat == 2], [0, 1, 2, 3], default=-1)
        here = (p["type"] == 2) & ((p["position"] & 3) == cls) & ok_a & ok_b & (idx >= 0)
        off = np.take_along_axis(p["offset"].astype(np.int32), np.maximum(idx, 0)[..., None], axis=2)[..., 0]
        delta = np.where(here, off, delta)
    return np.clip(c + delta, 0, 2 ** bit_depth - 1).astype(plane.dtype)


def sao_frame(frame, w, h, params, ctb_size=64, bit_depth=8):
    cw, ch = w // 2, h // 2
    y = frame[:w * h].reshape(h, w)
    u = frame[w * h:w * h + cw * ch].reshape(ch, cw)
    v = frame[w * h + cw * ch:].reshape(ch, cw)
    return np.concatenate([sao_plane(y, params[0], ctb_size, bit_depth).reshape(-1), sao_plane(u, params[1], ctb_size // 2, bit_depth).reshape(-1),
                           sao_plane(v, params[2], ctb_size // 2, bit_depth).reshape(-1)])
