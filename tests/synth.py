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


# ---------------------------------------------------------------------------------------------
# edge maps (edge byte = bs | qp << 2) and the per-8x8-block decoder data they are derived from
# ---------------------------------------------------------------------------------------------
BLOCK_INFO = np.dtype([("mv", np.int16, (2, 2)), ("ref", np.int8, (2,)), ("flags", np.uint8), ("qp", np.uint8)])
BLK_INTRA, BLK_CBF, BLK_TU_LEFT, BLK_TU_TOP, BLK_PU_LEFT, BLK_PU_TOP = 1, 2, 4, 8, 16, 32


def edge_bytes(bs, qp):
    return ((np.asarray(bs, np.uint32) & 3) | (np.asarray(qp, np.uint32) << 2)).astype(np.uint8)


def make_edge_maps(rng, w, h, kind="random", picture_edges=False):
    """(vert_edges[nV], hor_edges[nH]).  kind: "random" (bs U{0..2}, qp U{14..51} per entry),
    "wide" (bs U{0..3}, qp U{0..63}: the six qp bits past 51 act as 51), "smooth" (qp varies slowly, as
    in a rate-controlled picture), "uniform" (one qp everywhere).  Unless picture_edges, the left
    and bottom picture edges are off (the H.265 convention; see oracle_py.ref_compose_edges)."""
    h8, w8 = h // 8, w // 8
    shape_v, shape_h = (h8, w8 + 1), (h8 + 1, w8)
    if kind == "random":
        bs = [rng.integers(0, 3, s) for s in (shape_v, shape_h)]
        qp = [rng.integers(14, 52, s) for s in (shape_v, shape_h)]
    elif kind == "wide":
        bs = [rng.integers(0, 4, s) for s in (shape_v, shape_h)]
        qp = [rng.integers(0, 64, s) for s in (shape_v, shape_h)]
    elif kind == "smooth":
        bs = [rng.integers(0, 3, s) for s in (shape_v, shape_h)]
        base = rng.integers(22, 46)
        qp = []
        for s in (shape_v, shape_h):
            yy, xx = np.mgrid[0:s[0], 0:s[1]]
            qp.append(np.clip(base + np.round(6 * np.sin(xx / 9.0) + 5 * np.cos(yy / 7.0)).astype(np.int64)
                              + rng.integers(-1, 2, s), 0, 51))
    elif kind == "uniform":
        q = int(rng.integers(18, 52))
        bs = [rng.integers(0, 3, s) for s in (shape_v, shape_h)]
        qp = [np.full(s, q) for s in (shape_v, shape_h)]
    else:
        raise ValueError(kind)
    v, hh = edge_bytes(bs[0], qp[0]), edge_bytes(bs[1], qp[1])
    if not picture_edges:
        v[:, 0] &= 0xfc
        hh[-1, :] &= 0xfc
    return v.reshape(-1), hh.reshape(-1)


def make_block_info(rng, w, h, kind="cu16"):
    """BLOCK_INFO[h/8, w/8].  "intra": every block intra, every edge a transform edge; "random":
    every field random per block; "cu16": 16x16 coding units with their own motion, QP and coded
    flag, a fifth of them intra, a third bi-predicted, transform / prediction edges on the CU grid
    plus some inner transform splits -- neighbours that share motion give strength 0."""
    h8, w8 = h // 8, w // 8
    info = np.zeros((h8, w8), BLOCK_INFO)
    if kind == "intra":
        info["flags"] = BLK_INTRA | BLK_TU_LEFT | BLK_TU_TOP | BLK_PU_LEFT | BLK_PU_TOP
        info["ref"] = -1
        info["qp"] = rng.integers(20, 45)
    elif kind == "random":
        info["mv"] = rng.integers(-9, 10, info["mv"].shape)
        info["ref"] = rng.integers(-1, 3, info["ref"].shape)
        info["flags"] = rng.integers(0, 64, info.shape) & np.where(rng.random(info.shape) < 0.15, 63, 62)
        info["qp"] = rng.integers(0, 70, info.shape)
    elif kind == "extreme":
        # vectors at the ends of the 16-bit range and around the threshold of 4; few distinct pictures
        vals = np.array([-32768, -32767, -32765, -32764, -5, -4, -3, 0, 3, 4, 5, 32763, 32764, 32766, 32767])
        info["mv"] = vals[rng.integers(0, vals.size, info["mv"].shape)]
        info["ref"] = rng.integers(-2, 2, info["ref"].shape)
        info["flags"] = rng.integers(0, 64, info.shape) & 60
        info["qp"] = rng.integers(0, 256, info.shape)
    elif kind == "cu16":
        ch, cw = (h8 + 1) // 2, (w8 + 1) // 2
        up = lambda a: np.kron(a, np.ones((2, 2), a.dtype))[:h8, :w8]
        # motion drawn from a few candidates so that neighbouring units often agree
        cand = rng.integers(-24, 25, (6, 2, 2))
        pick = rng.integers(0, 6, (ch, cw))
        mv = cand[pick] + (rng.random((ch, cw, 1, 1)) < 0.3) * rng.integers(-5, 6, (ch, cw, 2, 2))
        for l in range(2):
            for c in range(2):
                info["mv"][:, :, l, c] = up(mv[:, :, l, c])
        ref0 = rng.integers(0, 2, (ch, cw))
        bi = rng.random((ch, cw)) < 0.33
        ref1 = np.where(bi, rng.integers(0, 3, (ch, cw)), -1)
        only_l1 = (~bi) & (rng.random((ch, cw)) < 0.2)
        ref1 = np.where(only_l1, ref0, ref1)
        ref0 = np.where(only_l1, -1, ref0)
        info["ref"][:, :, 0] = up(ref0.astype(np.int64))
        info["ref"][:, :, 1] = up(ref1.astype(np.int64))
        intra = up((rng.random((ch, cw)) < 0.2).astype(np.int64)).astype(bool)
        cbf = up((rng.random((ch, cw)) < 0.5).astype(np.int64)).astype(bool)
        info["qp"] = up(np.clip(rng.integers(24, 40) + rng.integers(-6, 7, (ch, cw)), 0, 51))
        yy, xx = np.mgrid[0:h8, 0:w8]
        split = up((rng.random((ch, cw)) < 0.25).astype(np.int64)).astype(bool)   # units split into four 8x8 transforms
        flags = intra * BLK_INTRA + cbf * BLK_CBF
        flags = flags + ((xx % 2 == 0) | split) * BLK_TU_LEFT + ((yy % 2 == 0) | split) * BLK_TU_TOP
        flags = flags + (xx % 2 == 0) * BLK_PU_LEFT + (yy % 2 == 0) * BLK_PU_TOP
        info["flags"] = flags
        info["ref"][intra] = -1
    else:
        raise ValueError(kind)
    return info
