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
