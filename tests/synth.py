"""Synthetic YV12 frames, QPs and boundary-strength maps for tests and bench (numpy, CPU).

"blocky" is the recipe of SURVEY.md section 8(d) config 3: smooth ramp + a DC step per 8x8
block in U{-12..12} + noise in U{-2..2}, clipped -- it makes the strong, normal and "off"
paths all fire.  The GPU bench builds the same kind of content with torch on the device.
"""
import numpy as np


def _blocky_plane(rng, w, h):
    yy, xx = np.mgrid[0:h, 0:w]
    ramp = 96.0 + 64.0 * (xx / max(w, 1)) + 32.0 * (yy / max(h, 1))
    bh, bw = (h + 7) // 8, (w + 7) // 8
    dc = rng.integers(-12, 13, size=(bh, bw))
    dc = np.kron(dc, np.ones((8, 8), dtype=np.int64))[:h, :w]
    noise = rng.integers(-2, 3, size=(h, w))
    return np.clip(ramp + dc + noise, 0, 255).astype(np.uint8)


def make_frame(rng, w, h, kind="blocky"):
    """One tight YV12 frame, uint8[3*w*h/2]."""
    cw, ch = w // 2, h // 2
    if kind == "blocky":
        planes = [_blocky_plane(rng, w, h), _blocky_plane(rng, cw, ch), _blocky_plane(rng, cw, ch)]
    elif kind == "random":
        planes = [rng.integers(0, 256, (h, w), dtype=np.uint8),
                  rng.integers(0, 256, (ch, cw), dtype=np.uint8),
                  rng.integers(0, 256, (ch, cw), dtype=np.uint8)]
    elif kind == "flat":
        v = int(rng.integers(0, 256))
        planes = [np.full((h, w), v, np.uint8), np.full((ch, cw), v, np.uint8), np.full((ch, cw), v, np.uint8)]
    elif kind == "ramp":
        planes = [((np.arange(w)[None, :] * 3 + np.arange(hh)[:, None] * 2) % 256).astype(np.uint8).repeat(1, 0)[:hh, :ww]
                  for (ww, hh) in ((w, h), (cw, ch), (cw, ch))]
        planes = [((np.arange(ww)[None, :] * 3 + np.arange(hh)[:, None] * 2) % 256).astype(np.uint8)
                  for (ww, hh) in ((w, h), (cw, ch), (cw, ch))]
    elif kind == "saturated":
        # 0/255 blocks with small noise: drives the 0..255 clips
        planes = []
        for (ww, hh) in ((w, h), (cw, ch), (cw, ch)):
            bh, bw = (hh + 7) // 8, (ww + 7) // 8
            base = rng.integers(0, 2, size=(bh, bw)) * 255
            base = np.kron(base, np.ones((8, 8), dtype=np.int64))[:hh, :ww]
            planes.append(np.clip(base + rng.integers(-3, 4, size=(hh, ww)), 0, 255).astype(np.uint8))
    elif kind == "steps":
        # small steps across every 8x8 edge on a flat field: strong-filter territory
        planes = []
        for (ww, hh) in ((w, h), (cw, ch), (cw, ch)):
            bh, bw = (hh + 7) // 8, (ww + 7) // 8
            dc = rng.integers(-3, 4, size=(bh, bw))
            dc = np.kron(dc, np.ones((8, 8), dtype=np.int64))[:hh, :ww]
            planes.append(np.clip(128 + dc, 0, 255).astype(np.uint8))
    else:
        raise ValueError(kind)
    return np.concatenate([p.reshape(-1) for p in planes])


def num_vert_bs(w, h):
    return (w // 8 + 1) * h // 8


def num_hor_bs(w, h):
    return (h // 8 + 1) * w // 8


def make_bs(rng, w, h, kind="uniform012"):
    nv, nh = num_vert_bs(w, h), num_hor_bs(w, h)
    if kind == "uniform012":
        return rng.integers(0, 3, nv, dtype=np.uint8), rng.integers(0, 3, nh, dtype=np.uint8)
    if kind == "zeros":
        return np.zeros(nv, np.uint8), np.zeros(nh, np.uint8)
    if kind == "ones":
        return np.ones(nv, np.uint8), np.ones(nh, np.uint8)
    if kind == "twos":
        return np.full(nv, 2, np.uint8), np.full(nh, 2, np.uint8)
    if kind == "wide":  # values > 2 still mean "filter" for luma (cpu.h:164)
        return rng.integers(0, 256, nv, dtype=np.uint8), rng.integers(0, 256, nh, dtype=np.uint8)
    raise ValueError(kind)


def make_batch(seed, n, w, h, kind="blocky", bs_kind="uniform012", qp_lo=18, qp_hi=51):
    """(frames[n, 3wh/2], qps[n] uint32, vert_bs[n, nV], hor_bs[n, nH]); a few QPs land on the
    no-op (<18) and clamp (>51) ranges when n is large enough."""
    rng = np.random.default_rng(seed)
    frames = np.stack([make_frame(rng, w, h, kind) for _ in range(n)])
    qps = rng.integers(qp_lo, qp_hi + 1, n).astype(np.uint32)
    if n >= 8:
        qps[3] = 12
        qps[5] = 17
        qps[6] = 60
    bs = [make_bs(rng, w, h, bs_kind) for _ in range(n)]
    return frames, qps, np.stack([b[0] for b in bs]), np.stack([b[1] for b in bs])
