"""ctypes bindings for the CPU checkers -- TEST INFRASTRUCTURE ONLY.

Two libraries, both built by ``oracle/Makefile``:

* ``liboracle_deblock.so``  -- the plain-C restatement (``deblock_oracle.c``), kind "port";
* ``_ref/libref_deblock.so`` -- the unmodified reference header
  (``hevc_deblocking_filter_cpu.h``) behind ``ref_harness.cpp``, kind "reference".

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` may
import this module.  Nothing under ``gpu_video_codec_b200/`` does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT = os.path.join(_HERE, "liboracle_deblock.so")
_REF = os.path.join(_HERE, "_ref", "libref_deblock.so")

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint)


def build(quiet=True):
    """Compile the checkers (the reference harness only where /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a, ty=_u8p):
    return None if a is None else a.ctypes.data_as(ty)


def _load_port():
    if not os.path.exists(_PORT):
        build()
    lib = C.CDLL(_PORT)
    lib.hdbo_beta.argtypes = [C.c_uint]
    lib.hdbo_tc.argtypes = [C.c_uint]
    lib.hdbo_num_vert_bs.argtypes = [C.c_uint, C.c_uint]
    lib.hdbo_num_vert_bs.restype = C.c_uint
    lib.hdbo_num_hor_bs.argtypes = [C.c_uint, C.c_uint]
    lib.hdbo_num_hor_bs.restype = C.c_uint
    lib.hdbo_default_vert_bs.argtypes = [C.c_uint, C.c_uint, _u8p]
    lib.hdbo_default_hor_bs.argtypes = [C.c_uint, C.c_uint, _u8p]
    lib.hdbo_deblock_frame.argtypes = [_u8p, _u8p, C.c_uint, C.c_uint, C.c_uint,
                                       _u8p, C.c_uint, _u8p, C.c_uint]
    lib.hdbo_deblock_batch.argtypes = [_u8p, _u8p, C.c_uint, C.c_uint, C.c_uint,
                                       _u32p, _u8p, _u8p]
    lib.hdbo_deblock_frame_edges.argtypes = [_u8p, _u8p, C.c_uint, C.c_uint, C.c_uint,
                                             _u8p, C.c_uint, _u8p, C.c_uint]
    lib.hdbo_derive_edges.argtypes = [C.c_void_p, C.c_uint, C.c_uint, _u8p, _u8p]
    lib.hdbo_derive_edges.restype = None
    lib.hdbo_sao_frame.argtypes = [_u8p, _u8p, C.c_uint, C.c_uint, C.c_void_p, C.c_uint]
    lib.hdbo_sao_frame16.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_void_p, C.c_uint, C.c_uint]
    lib.hdbo_deblock_frame16.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_uint,
                                         _u8p, C.c_uint, _u8p, C.c_uint, C.c_uint, C.c_int]
    return lib


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        _port = _load_port()
    return _port


def have_ref():
    return os.path.exists(_REF)


def ref():
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libref_deblock.so not built (needs /root/reference)")
        lib = C.CDLL(_REF)
        lib.ref_deblock_frame.argtypes = [_u8p, C.c_size_t, _u8p, C.c_uint, C.c_uint, C.c_uint,
                                          _u8p, C.c_uint, _u8p, C.c_uint, C.c_uint,
                                          C.POINTER(C.c_double), C.c_char_p, C.c_uint]
        lib.ref_deblock_batch.argtypes = [_u8p, _u8p, C.c_uint, C.c_uint, C.c_uint,
                                          _u32p, _u8p, _u8p, C.c_uint, C.c_uint,
                                          C.POINTER(C.c_double)]
        _ref = lib
    return _ref


def num_vert_bs(w, h):
    return int(port().hdbo_num_vert_bs(w, h))


def num_hor_bs(w, h):
    return int(port().hdbo_num_hor_bs(w, h))


def default_bs(w, h):
    v = np.empty(num_vert_bs(w, h), np.uint8)
    hh = np.empty(num_hor_bs(w, h), np.uint8)
    port().hdbo_default_vert_bs(w, h, _ptr(v))
    port().hdbo_default_hor_bs(w, h, _ptr(hh))
    return v, hh


class OracleError(RuntimeError):
    pass


def _check_frame(frame, w, h):
    frame = np.ascontiguousarray(frame, np.uint8).reshape(-1)
    if frame.size != 3 * w * h // 2:
        raise OracleError("Incorrect file size")  # cpu.h:43-45
    return frame


def deblock_frame(frame, w, h, qp, vert_bs=None, hor_bs=None):
    """Restatement (kind "port").  frame: uint8[3*w*h/2] tight YV12."""
    frame = _check_frame(frame, w, h)
    out = np.empty_like(frame)
    v = None if vert_bs is None else np.ascontiguousarray(vert_bs, np.uint8)
    hh = None if hor_bs is None else np.ascontiguousarray(hor_bs, np.uint8)
    rc = port().hdbo_deblock_frame(_ptr(frame), _ptr(out), w, h, qp,
                                   _ptr(v), 0 if v is None else v.size,
                                   _ptr(hh), 0 if hh is None else hh.size)
    if rc:
        raise OracleError({1: "Width and height of image must be multiplier of sample block size",
                           2: "Incorrect size of input boundary strenght array"}[rc])
    return out


def deblock_batch(frames, w, h, qps, vert_bs=None, hor_bs=None):
    """frames: uint8[n, 3*w*h/2]; qps: uint32[n]; vert_bs/hor_bs: uint8[n, nV] / [n, nH] or None."""
    frames = np.ascontiguousarray(frames, np.uint8)
    n = frames.shape[0]
    out = np.empty_like(frames)
    q = np.ascontiguousarray(qps, np.uint32)
    v = None if vert_bs is None else np.ascontiguousarray(vert_bs, np.uint8)
    hh = None if hor_bs is None else np.ascontiguousarray(hor_bs, np.uint8)
    rc = port().hdbo_deblock_batch(_ptr(frames), _ptr(out), n, w, h, _ptr(q, _u32p), _ptr(v), _ptr(hh))
    if rc:
        raise OracleError("oracle batch failed rc=%d" % rc)
    return out


def ref_deblock_frame(frame, w, h, qp, vert_bs=None, hor_bs=None, threads=1):
    """The unmodified reference (kind "reference").  Returns (out, seconds of DeblockingFilter)."""
    frame = np.ascontiguousarray(frame, np.uint8).reshape(-1)
    out = np.empty(max(frame.size, 3 * w * h // 2, 1), np.uint8)
    v = None if vert_bs is None else np.ascontiguousarray(vert_bs, np.uint8)
    hh = None if hor_bs is None else np.ascontiguousarray(hor_bs, np.uint8)
    sec = C.c_double(0.0)
    err = C.create_string_buffer(256)
    # the harness writes exactly what it is given; the reference checks the file length (cpu.h:43)
    rc = ref().ref_deblock_frame(_ptr(frame), frame.size, _ptr(out), w, h, qp,
                                 _ptr(v), 0 if v is None else v.size,
                                 _ptr(hh), 0 if hh is None else hh.size,
                                 threads, C.byref(sec), err, 256)
    if rc == 1:
        raise OracleError(err.value.decode())
    if rc:
        raise RuntimeError("reference harness I/O failure")
    return out[:frame.size], sec.value


def ref_deblock_batch(frames, w, h, qps, vert_bs=None, hor_bs=None, outer_threads=1, inner_threads=1):
    frames = np.ascontiguousarray(frames, np.uint8)
    n = frames.shape[0]
    out = np.empty_like(frames)
    q = np.ascontiguousarray(qps, np.uint32)
    v = None if vert_bs is None else np.ascontiguousarray(vert_bs, np.uint8)
    hh = None if hor_bs is None else np.ascontiguousarray(hor_bs, np.uint8)
    sec = C.c_double(0.0)
    rc = ref().ref_deblock_batch(_ptr(frames), _ptr(out), n, w, h, _ptr(q, _u32p), _ptr(v), _ptr(hh),
                                 outer_threads, inner_threads, C.byref(sec))
    if rc:
        raise RuntimeError("reference batch failed rc=%d" % rc)
    return out, sec.value


# ---------------------------------------------------------------------------
# edge maps: edge byte = bs | qp << 2
# ---------------------------------------------------------------------------
BLOCK_INFO = np.dtype([("mv", np.int16, (2, 2)), ("ref", np.int8, (2,)), ("flags", np.uint8), ("qp", np.uint8)])
assert BLOCK_INFO.itemsize == 12
BLK_INTRA, BLK_CBF, BLK_TU_LEFT, BLK_TU_TOP, BLK_PU_LEFT, BLK_PU_TOP = 1, 2, 4, 8, 16, 32


def edge_bytes(bs, qp):
    return ((np.asarray(bs, np.uint32) & 3) | (np.asarray(qp, np.uint32) << 2)).astype(np.uint8)


def deblock_frame_edges(frame, w, h, qp, vert_edges, hor_edges):
    """Restatement with per-edge QP for luma (deblock_oracle.h); chroma as the reference (frame qp)."""
    frame = _check_frame(frame, w, h)
    out = np.empty_like(frame)
    v = np.ascontiguousarray(vert_edges, np.uint8)
    hh = np.ascontiguousarray(hor_edges, np.uint8)
    rc = port().hdbo_deblock_frame_edges(_ptr(frame), _ptr(out), w, h, qp, _ptr(v), v.size, _ptr(hh), hh.size)
    if rc:
        raise OracleError({1: "Width and height of image must be multiplier of sample block size",
                           2: "Incorrect size of input boundary strenght array"}[rc])
    return out


def deblock_frame16(frame, w, h, qp, vert=None, hor=None, bit_depth=10, edges=False):
    """16-bit containers holding bit_depth-bit samples (deblock_oracle.h: hdbo_deblock_frame16).  Depth 8 is the
    reference's filter; deeper is the H.265 scaling rule on top of it (unpinned against the reference)."""
    frame = np.ascontiguousarray(frame, np.uint16).reshape(-1)
    if frame.size != 3 * w * h // 2:
        raise OracleError("Incorrect file size")
    out = np.empty_like(frame)
    v = None if vert is None else np.ascontiguousarray(vert, np.uint8)
    hh = None if hor is None else np.ascontiguousarray(hor, np.uint8)
    rc = port().hdbo_deblock_frame16(frame.ctypes.data, out.ctypes.data, w, h, qp,
                                     None if v is None else _ptr(v), 0 if v is None else v.size,
                                     None if hh is None else _ptr(hh), 0 if hh is None else hh.size, bit_depth, int(edges))
    if rc:
        raise OracleError({1: "Width and height of image must be multiplier of sample block size",
                           2: "Incorrect size of input boundary strenght array"}[rc])
    return out


SAO_PARAMS = np.dtype([("type", np.uint8), ("position", np.uint8), ("offset", np.int8, (4,))])


def sao_frame(frame, w, h, params, ctb_size=64):
    """Sample adaptive offset of one 8-bit YV12 frame (deblock_oracle.h: hdbo_sao_frame; H.265 8.7.3, unpinned: the
    reference has no SAO).  params: SAO_PARAMS[3, ceil(h / ctb), ceil(w / ctb)]."""
    frame = _check_frame(frame, w, h)
    params = np.ascontiguousarray(params, SAO_PARAMS)
    assert params.shape == (3, (h + ctb_size - 1) // ctb_size, (w + ctb_size - 1) // ctb_size), params.shape
    out = np.empty_like(frame)
    rc = port().hdbo_sao_frame(_ptr(frame), _ptr(out), w, h, params.ctypes.data, ctb_size)
    if rc:
        raise OracleError("sao: bad size or ctb size")
    return out


def sao_frame16(frame, w, h, params, ctb_size=64, bit_depth=10):
    """The same for bit_depth-bit samples in 16-bit containers (deblock_oracle.h: hdbo_sao_frame16).
    frame: uint16[3wh/2], planar."""
    frame = np.ascontiguousarray(frame, np.uint16).reshape(-1)
    assert frame.size == 3 * w * h // 2, frame.size
    params = np.ascontiguousarray(params, SAO_PARAMS)
    assert params.shape == (3, (h + ctb_size - 1) // ctb_size, (w + ctb_size - 1) // ctb_size), params.shape
    out = np.empty_like(frame)
    rc = port().hdbo_sao_frame16(frame.ctypes.data, out.ctypes.data, w, h, params.ctypes.data, ctb_size, bit_depth)
    if rc:
        raise OracleError("sao: bad size, ctb size or bit depth")
    return out


def p010_to_planar(frame, w, h, depth=10):
    """P010 / P016 (Y plane, then interleaved UV; samples in the high bits of 16-bit containers) -> planar low-aligned."""
    f = np.ascontiguousarray(frame, np.uint16).reshape(-1)
    ysz, csz = w * h, (w // 2) * (h // 2)
    uv = f[ysz:].reshape(csz, 2)
    return (np.concatenate([f[:ysz], uv[:, 0], uv[:, 1]]) >> (16 - depth)).astype(np.uint16)


def planar_to_p010(frame, w, h, depth=10):
    f = np.ascontiguousarray(frame, np.uint16).reshape(-1)
    ysz, csz = w * h, (w // 2) * (h // 2)
    uv = np.stack([f[ysz:ysz + csz], f[ysz + csz:]], axis=1).reshape(-1)
    return (np.concatenate([f[:ysz], uv]) << (16 - depth)).astype(np.uint16)


def deblock_frame_p010(frame, w, h, qp, vert=None, hor=None, edges=False):
    """The 16-bit oracle between the two layout conversions (what a decoder's 10-bit surfaces hold)."""
    return planar_to_p010(deblock_frame16(p010_to_planar(frame, w, h), w, h, qp, vert, hor, bit_depth=10, edges=edges), w, h)


def derive_edges(info, w, h):
    """info: BLOCK_INFO[h/8, w/8] -> (vert_edges, hor_edges) by the H.265 rule (deblock_oracle.h)."""
    info = np.ascontiguousarray(info, BLOCK_INFO)
    assert info.size == (w // 8) * (h // 8)
    v = np.empty(num_vert_bs(w, h), np.uint8)
    hh = np.empty(num_hor_bs(w, h), np.uint8)
    port().hdbo_derive_edges(info.ctypes.data, w, h, _ptr(v), _ptr(hh))
    return v, hh


def ref_compose_edges(frame, w, h, qp, vert_edges, hor_edges):
    """Per-edge QP through the UNMODIFIED reference, as a chain of reference runs over the luma plane,
    each with one QP and with only that QP's edges enabled (all other strengths 0):
      1. one run per distinct QP of the vertical map (vertical segments never touch each other's
         samples, so their order does not matter);
      2. horizontal entries column by column in increasing order, one run per distinct QP of the
         column: entry (by, c) is the right-hand segment of block c and the left-hand segment of
         block c + 1 (cpu.h:287-291, 368-372), so increasing columns reproduce each block's own
         order left segment -> right segment (cpu.h:287-446), and within one column all segments
         lie in different blocks.
    Blocks are independent and every block runs its vertical segments first (cpu.h:158-284), so the
    chain equals the per-edge-QP filter exactly -- provided no segment reads plane padding an earlier
    segment of its block wrote, because padding does not survive from one run to the next.  That
    happens in two places only: the left picture edge (vert[:, 0]; its p samples are padding that
    the right-hand horizontal segment of block column 0 reads as q) and the bottom picture edge
    (hor[H/8, :]).  Maps with those two off -- the H.265 convention, and what derive_edges produces --
    are pinned exactly; for the others the restatement stands on its uniform-QP pin, which covers the
    padding paths.  The chroma planes come from one run at the frame QP."""
    frame = np.ascontiguousarray(frame, np.uint8).reshape(-1)
    v = np.ascontiguousarray(vert_edges, np.uint8)
    hh = np.ascontiguousarray(hor_edges, np.uint8)
    ysz = w * h
    chroma, _ = ref_deblock_frame(frame, w, h, qp)
    cur = frame.copy()
    zero_v, zero_h = np.zeros_like(v), np.zeros_like(hh)

    def run(q, mv, mh):
        out, _ = ref_deblock_frame(cur, w, h, int(q), mv, mh)
        cur[:ysz] = out[:ysz]

    on = (v & 3) > 0
    for q in np.unique(v[on] >> 2):
        run(q, np.where(on & ((v >> 2) == q), v & 3, 0).astype(np.uint8), zero_h)
    h2 = hh.reshape(h // 8 + 1, w // 8)
    for c in range(w // 8):
        col = h2[:, c]
        on = (col & 3) > 0
        for q in np.unique(col[on] >> 2):
            m = np.zeros_like(h2)
            m[:, c] = np.where(on & ((col >> 2) == q), col & 3, 0)
            run(q, zero_v, m.reshape(-1))
    cur[ysz:] = chroma[ysz:]
    return cur


# ---------------------------------------------------------------------------
# NV12: Y plane, then one plane of interleaved U0 V0 U1 V1 ...
# The reference reads planar frames only (cpu.h:66-82) and filters its two chroma planes alike,
# so the NV12 oracle is the YV12 one between two byte permutations.
# ---------------------------------------------------------------------------
def yv12_to_nv12(frame, w, h):
    frame = np.ascontiguousarray(frame, np.uint8).reshape(-1)
    ysz, csz = w * h, (w // 2) * (h // 2)
    out = np.empty_like(frame)
    out[:ysz] = frame[:ysz]
    uv = out[ysz:].reshape(h // 2, w // 2, 2)
    uv[:, :, 0] = frame[ysz:ysz + csz].reshape(h // 2, w // 2)
    uv[:, :, 1] = frame[ysz + csz:].reshape(h // 2, w // 2)
    return out


def nv12_to_yv12(frame, w, h):
    frame = np.ascontiguousarray(frame, np.uint8).reshape(-1)
    ysz, csz = w * h, (w // 2) * (h // 2)
    out = np.empty_like(frame)
    out[:ysz] = frame[:ysz]
    uv = frame[ysz:].reshape(h // 2, w // 2, 2)
    out[ysz:ysz + csz] = uv[:, :, 0].reshape(-1)
    out[ysz + csz:] = uv[:, :, 1].reshape(-1)
    return out


def deblock_frame_nv12(frame, w, h, qp, vert=None, hor=None, edges=False):
    planar = nv12_to_yv12(frame, w, h)
    y = deblock_frame_edges(planar, w, h, qp, vert, hor) if edges else deblock_frame(planar, w, h, qp, vert, hor)
    return yv12_to_nv12(y, w, h)
