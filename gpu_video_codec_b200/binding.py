"""ctypes binding of include/hevc_deblock_b200.h (one-to-one, no logic of its own).

Names follow the reference's interface for this path
(hevc_deblocking_filter_cpu.h: ReadYuvFrame / SetBoundaryStrenght / DeblockingFilter / Save;
hevc_deblocking_filter_gpu.cu: ExecuteGpu) so the parity tests read like calls into it.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
# HDB_LIBRARY selects another build of the same library (kernel-shape tuning runs)
_LIB = os.environ.get("HDB_LIBRARY") or os.path.join(_PKG, "libhevc_deblock_b200.so")
_HEADER = os.path.join(_ROOT, "include", "hevc_deblock_b200.h")

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint)


class DeblockError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("hdb status %d: %s" % (status, message))
        self.status = status
        self.message = message


class Timing(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("copy_ms", C.c_double), ("total_ms", C.c_double),
                ("h2d_bytes", C.c_ulonglong), ("d2h_bytes", C.c_ulonglong),
                ("kernel_launches", C.c_int), ("path", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Surfaces(C.Structure):
    """hdb_surfaces: three device planes with row pitches."""
    _fields_ = [("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p),
                ("pitch_y", C.c_size_t), ("pitch_c", C.c_size_t),
                ("frame_stride_y", C.c_size_t), ("frame_stride_c", C.c_size_t)]


class Nv12Surfaces(C.Structure):
    """hdb_nv12_surfaces: device Y plane and interleaved UV plane with row pitches."""
    _fields_ = [("y", C.c_void_p), ("uv", C.c_void_p), ("pitch_y", C.c_size_t), ("pitch_uv", C.c_size_t),
                ("frame_stride_y", C.c_size_t), ("frame_stride_uv", C.c_size_t)]


class StripPlanes(C.Structure):
    """hdb_strip_planes: the three planes of one strip in one GPU's memory."""
    _fields_ = [("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p)]


FLAG_EDGE_MAPS, FLAG_NV12, FLAG_SAMPLES16 = 1, 2, 4


def flag_depth(bits):
    """HDB_FLAG_DEPTH(bits): planar frames of 16-bit containers holding `bits`-bit samples (8, 9 or 10)."""
    return FLAG_SAMPLES16 | (int(bits) << 8)


class DeviceInfo(C.Structure):
    _fields_ = [("name", C.c_char * 256), ("cc_major", C.c_int), ("cc_minor", C.c_int),
                ("sm_count", C.c_int), ("warp_size", C.c_int), ("max_threads_per_block", C.c_int),
                ("total_global_mem", C.c_ulonglong), ("total_const_mem", C.c_ulonglong),
                ("shared_mem_per_block", C.c_ulonglong)]


def library_path():
    return _LIB


def build_library(quiet=True, **knobs):
    """nvcc-compile csrc/ for sm_100a into the in-tree .so (cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_PKG, "csrc")] + ["%s=%s" % kv for kv in knobs.items()]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL if quiet else None)
    return _LIB


def header_symbols():
    """Every function the public header declares."""
    text = open(_HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hdb_[a-z_0-9]+)\s*\(", text)))


def exported_symbols():
    out = subprocess.run(["nm", "-D", "--defined-only", _LIB], check=True, capture_output=True, text=True).stdout
    return sorted({ln.split()[-1] for ln in out.splitlines() if ln.split() and ln.split()[-1].startswith("hdb_")})


_lib = None


def load_library():
    """Load the C-ABI library.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise RuntimeError("%s is missing: run __graft_entry__.build() (nvcc, sm_100a). "
                           "There is no CPU or PyTorch fallback for the deblocking path." % _LIB)
    lib = C.CDLL(_LIB)
    vp = C.c_void_p
    lib.hdb_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.hdb_destroy.argtypes = [vp]
    lib.hdb_destroy.restype = None
    lib.hdb_last_error.argtypes = [vp]
    lib.hdb_last_error.restype = C.c_char_p
    lib.hdb_status_string.argtypes = [C.c_int]
    lib.hdb_status_string.restype = C.c_char_p
    lib.hdb_get_device_info.argtypes = [vp, C.POINTER(DeviceInfo)]
    lib.hdb_num_vert_bs.argtypes = [C.c_uint, C.c_uint]
    lib.hdb_num_vert_bs.restype = C.c_uint
    lib.hdb_num_hor_bs.argtypes = [C.c_uint, C.c_uint]
    lib.hdb_num_hor_bs.restype = C.c_uint
    lib.hdb_default_boundary_strength.argtypes = [C.c_uint, C.c_uint, _u8p, _u8p]
    lib.hdb_beta.argtypes = [C.c_uint]
    lib.hdb_tc.argtypes = [C.c_uint]
    lib.hdb_host_alloc.argtypes = [C.c_size_t]
    lib.hdb_host_alloc.restype = vp
    lib.hdb_host_free.argtypes = [vp]
    lib.hdb_host_free.restype = None
    frame_args = [C.c_uint, C.c_uint, C.c_uint, vp, C.c_uint, vp, C.c_uint, C.POINTER(Timing)]
    lib.hdb_deblock_frame.argtypes = [vp, vp, vp] + frame_args
    lib.hdb_deblock_planes.argtypes = [vp, vp, vp, vp, vp, vp, vp] + frame_args
    lib.hdb_execute_gpu.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_uint, C.c_uint, C.c_uint, C.POINTER(Timing)]
    batch_args = [vp, vp, C.c_uint, C.c_uint, C.c_uint, vp, vp, vp]
    lib.hdb_deblock_batch.argtypes = [vp] + batch_args + [C.POINTER(Timing)]
    lib.hdb_deblock_batch_device.argtypes = [vp] + batch_args + [vp]
    lib.hdb_deblock_surfaces_device.argtypes = [vp, C.POINTER(Surfaces), C.POINTER(Surfaces), C.c_uint, C.c_uint, C.c_uint,
                                                vp, vp, vp, vp]
    lib.hdb_deblock_frame_edges.argtypes = [vp, vp, vp] + frame_args
    lib.hdb_deblock_batch_edges.argtypes = [vp] + batch_args + [C.POINTER(Timing)]
    lib.hdb_deblock_batch_edges_device.argtypes = [vp] + batch_args + [vp]
    lib.hdb_derive_edges_device.argtypes = [vp, vp, C.c_uint, C.c_uint, C.c_uint, vp, vp, vp]
    lib.hdb_derive_edges.argtypes = [vp, vp, C.c_uint, C.c_uint, C.c_uint, vp, vp]
    lib.hdb_deblock_batch_ex.argtypes = [vp] + batch_args + [C.c_uint, C.POINTER(Timing)]
    lib.hdb_deblock_batch_blocks.argtypes = [vp, vp, vp, C.c_uint, C.c_uint, C.c_uint, vp, vp, C.c_uint, C.POINTER(Timing)]
    lib.hdb_deblock_batch_device_ex.argtypes = [vp] + batch_args + [C.c_uint, vp]
    lib.hdb_deblock_nv12_surfaces_device.argtypes = [vp, C.POINTER(Nv12Surfaces), C.POINTER(Nv12Surfaces), C.c_uint, C.c_uint,
                                                     C.c_uint, vp, vp, vp, C.c_uint, vp]
    lib.hdb_execute_gpu_sequence.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint,
                                             C.POINTER(C.c_uint), C.POINTER(Timing)]
    lib.hdb_sao_batch_device.argtypes = [vp, vp, vp, C.c_uint, C.c_uint, C.c_uint, vp, C.c_uint, vp]
    lib.hdb_sao_batch.argtypes = [vp, vp, vp, C.c_uint, C.c_uint, C.c_uint, vp, C.c_uint]
    lib.hdb_sao_batch_device_ex.argtypes = [vp, vp, vp, C.c_uint, C.c_uint, C.c_uint, vp, C.c_uint, C.c_uint, vp]
    lib.hdb_sao_batch_ex.argtypes = [vp, vp, vp, C.c_uint, C.c_uint, C.c_uint, vp, C.c_uint, C.c_uint]
    lib.hdb_deblock_sao_batch_device.argtypes = [vp, vp, vp, C.c_uint, C.c_uint, C.c_uint, vp, vp, vp, C.c_uint, vp, C.c_uint, vp]
    lib.hdb_kernel_launch_count.argtypes = [vp]
    lib.hdb_kernel_launch_count.restype = C.c_ulonglong
    lib.hdb_multi_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
    lib.hdb_multi_destroy.argtypes = [vp]
    lib.hdb_multi_destroy.restype = None
    lib.hdb_multi_device_count.argtypes = [vp]
    lib.hdb_multi_last_error.argtypes = [vp]
    lib.hdb_multi_last_error.restype = C.c_char_p
    lib.hdb_multi_deblock_batch.argtypes = [vp] + batch_args + [C.POINTER(Timing)]
    lib.hdb_multi_deblock_batch_ex.argtypes = [vp] + batch_args + [C.c_uint, C.POINTER(Timing)]
    lib.hdb_multi_deblock_frame_strips.argtypes = [vp, vp, vp] + frame_args
    lib.hdb_multi_deblock_frame_strips_ex.argtypes = [vp, vp, vp] + frame_args[:-1] + [C.c_uint, C.POINTER(Timing)]
    lib.hdb_multi_deblock_stream_strips.argtypes = [vp] + batch_args + [C.c_uint, C.POINTER(Timing)]
    lib.hdb_multi_strip_layout.argtypes = [vp, C.c_uint, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.hdb_multi_deblock_strips_device.argtypes = [vp, C.POINTER(StripPlanes), C.POINTER(StripPlanes), C.c_uint, C.c_uint, C.c_uint,
                                                    C.POINTER(vp), C.POINTER(vp), C.c_uint, C.POINTER(vp), C.POINTER(Timing)]
    lib.hdb_multi_strip_halo_bytes.argtypes = [vp, C.c_uint, C.c_uint]
    lib.hdb_multi_strip_halo_bytes.restype = C.c_ulonglong
    _lib = lib
    return lib


def num_vert_bs(width, height):
    return int(load_library().hdb_num_vert_bs(width, height))


def num_hor_bs(width, height):
    return int(load_library().hdb_num_hor_bs(width, height))


def default_boundary_strength(width, height):
    """The reference constructor's default luma strengths (cpu.h:92-99)."""
    lib = load_library()
    v = np.empty(num_vert_bs(width, height), np.uint8)
    h = np.empty(num_hor_bs(width, height), np.uint8)
    rc = lib.hdb_default_boundary_strength(width, height, v.ctypes.data_as(_u8p), h.ctypes.data_as(_u8p))
    if rc:
        raise DeblockError(rc, lib.hdb_status_string(rc).decode())
    return v, h


# hdb_sao_params: sample adaptive offset of one coding tree block and colour component (6 bytes)
SAO_PARAMS = np.dtype([("type", np.uint8), ("position", np.uint8), ("offset", np.int8, (4,))])

# hdb_block_info: what a decoder knows about one 8x8 luma block (12 bytes)
BLOCK_INFO = np.dtype([("mv", np.int16, (2, 2)), ("ref", np.int8, (2,)), ("flags", np.uint8), ("qp", np.uint8)])
BLK_INTRA, BLK_CBF, BLK_TU_LEFT, BLK_TU_TOP, BLK_PU_LEFT, BLK_PU_TOP = 1, 2, 4, 8, 16, 32


def edge_bytes(bs, qp):
    """HDB_EDGE(bs, qp) = bs | qp << 2, elementwise."""
    return ((np.asarray(bs, np.uint32) & 3) | (np.minimum(np.asarray(qp, np.uint32), 51) << 2)).astype(np.uint8)


def _addr(a):
    """Address of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    return a.data_ptr()  # torch tensor (host pinned or device)


class Deblocker:
    """One GPU.  Wraps an hdb_context."""

    def __init__(self, device=0):
        self._lib = load_library()
        self._ctx = C.c_void_p()
        rc = self._lib.hdb_create(C.byref(self._ctx), device)
        if rc:
            raise DeblockError(rc, self._lib.hdb_last_error(None).decode())
        self.device = device

    def close(self):
        if self._ctx:
            self._lib.hdb_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise DeblockError(rc, self._lib.hdb_last_error(self._ctx).decode())

    def device_info(self):
        info = DeviceInfo()
        self._check(self._lib.hdb_get_device_info(self._ctx, C.byref(info)))
        return {k: (getattr(info, k).decode() if k == "name" else getattr(info, k)) for k, _ in info._fields_}

    def kernel_launch_count(self):
        return int(self._lib.hdb_kernel_launch_count(self._ctx))

    # --- one frame: ReadYuvFrame(...).SetBoundaryStrenght(...).DeblockingFilter().Save() ---
    def deblock_frame(self, yuv, width, height, qp, vert_bs=None, hor_bs=None, out=None, timing=None):
        yuv = np.ascontiguousarray(yuv, np.uint8).reshape(-1)
        if yuv.size != 3 * width * height // 2:
            raise DeblockError(2, "Incorrect file size")  # cpu.h:43-45 (the length check belongs to the file reader)
        out = np.empty_like(yuv) if out is None else out
        v = None if vert_bs is None else np.ascontiguousarray(vert_bs, np.uint8)
        h = None if hor_bs is None else np.ascontiguousarray(hor_bs, np.uint8)
        self._check(self._lib.hdb_deblock_frame(self._ctx, _addr(yuv), _addr(out), width, height, qp,
                                                _addr(v), 0 if v is None else v.size,
                                                _addr(h), 0 if h is None else h.size,
                                                None if timing is None else C.byref(timing)))
        return out

    def deblock_planes(self, y, u, v, width, height, qp, vert_bs=None, hor_bs=None, timing=None):
        y = np.ascontiguousarray(y, np.uint8)
        u = np.ascontiguousarray(u, np.uint8)
        v = np.ascontiguousarray(v, np.uint8)
        yo, uo, vo = np.empty_like(y), np.empty_like(u), np.empty_like(v)
        vb = None if vert_bs is None else np.ascontiguousarray(vert_bs, np.uint8)
        hb = None if hor_bs is None else np.ascontiguousarray(hor_bs, np.uint8)
        self._check(self._lib.hdb_deblock_planes(self._ctx, _addr(y), _addr(u), _addr(v), _addr(yo), _addr(uo),
                                                 _addr(vo), width, height, qp,
                                                 _addr(vb), 0 if vb is None else vb.size,
                                                 _addr(hb), 0 if hb is None else hb.size,
                                                 None if timing is None else C.byref(timing)))
        return yo, uo, vo

    def execute_gpu(self, input_file_name, output_file_name, width, height, qp, timing=None):
        """ExecuteGpu(input_file_name, output_file_name, width, height, Qp) of main.cu:87-90."""
        self._check(self._lib.hdb_execute_gpu(self._ctx, os.fsencode(input_file_name), os.fsencode(output_file_name),
                                              width, height, qp, None if timing is None else C.byref(timing)))

    def execute_gpu_sequence(self, input_file_name, output_file_name, width, height, qp, flags=0, timing=None):
        """A raw sequence file with any number of frames; returns the frame count."""
        n = C.c_uint(0)
        self._check(self._lib.hdb_execute_gpu_sequence(self._ctx, os.fsencode(input_file_name), os.fsencode(output_file_name),
                                                       width, height, qp, flags, C.byref(n),
                                                       None if timing is None else C.byref(timing)))
        return int(n.value)

    # --- batches ---
    def deblock_batch(self, frames, width, height, qps, vert_bs=None, hor_bs=None, out=None, timing=None,
                      n_frames=None):
        """Host buffers (numpy arrays or pinned torch tensors), pipelined H2D / kernel / D2H."""
        n = n_frames if n_frames is not None else int(frames.shape[0])
        if out is None:
            out = np.empty_like(frames)
        self._check(self._lib.hdb_deblock_batch(self._ctx, _addr(frames), _addr(out), n, width, height,
                                                _addr(qps), _addr(vert_bs), _addr(hor_bs),
                                                None if timing is None else C.byref(timing)))
        return out

    def deblock_batch_device(self, d_in, d_out, n_frames, width, height, d_qp, d_vert_bs=None, d_hor_bs=None,
                             stream=None):
        """Device pointers (torch CUDA tensors or ints); enqueues on `stream` and returns."""
        self._check(self._lib.hdb_deblock_batch_device(self._ctx, _addr(d_in), _addr(d_out), n_frames, width, height,
                                                       _addr(d_qp), _addr(d_vert_bs), _addr(d_hor_bs), stream))
    # --- edge maps: per-edge QP (edge byte = bs | qp << 2) and their derivation from block data ---
    def deblock_frame_edges(self, yuv, width, height, qp, vert_edges, hor_edges, out=None, timing=None):
        yuv = np.ascontiguousarray(yuv, np.uint8).reshape(-1)
        if yuv.size != 3 * width * height // 2:
            raise DeblockError(2, "Incorrect file size")
        out = np.empty_like(yuv) if out is None else out
        v = None if vert_edges is None else np.ascontiguousarray(vert_edges, np.uint8)
        h = None if hor_edges is None else np.ascontiguousarray(hor_edges, np.uint8)
        self._check(self._lib.hdb_deblock_frame_edges(self._ctx, _addr(yuv), _addr(out), width, height, qp,
                                                      _addr(v), 0 if v is None else v.size,
                                                      _addr(h), 0 if h is None else h.size,
                                                      None if timing is None else C.byref(timing)))
        return out

    def deblock_batch_edges(self, frames, width, height, qps, vert_edges, hor_edges, out=None, timing=None,
                            n_frames=None):
        n = n_frames if n_frames is not None else int(frames.shape[0])
        if out is None:
            out = np.empty_like(frames)
        self._check(self._lib.hdb_deblock_batch_edges(self._ctx, _addr(frames), _addr(out), n, width, height,
                                                      _addr(qps), _addr(vert_edges), _addr(hor_edges),
                                                      None if timing is None else C.byref(timing)))
        return out

    def deblock_batch_edges_device(self, d_in, d_out, n_frames, width, height, d_qp, d_vert_edges, d_hor_edges,
                                   stream=None):
        self._check(self._lib.hdb_deblock_batch_edges_device(self._ctx, _addr(d_in), _addr(d_out), n_frames, width,
                                                             height, _addr(d_qp), _addr(d_vert_edges),
                                                             _addr(d_hor_edges), stream))

    def derive_edges(self, info, width, height):
        """info: BLOCK_INFO[n, H/8, W/8] (host) -> (vert_edges[n, nV], hor_edges[n, nH]), computed on the GPU."""
        info = np.ascontiguousarray(info, BLOCK_INFO)
        n = info.size // ((width // 8) * (height // 8))
        v = np.empty((n, num_vert_bs(width, height)), np.uint8)
        h = np.empty((n, num_hor_bs(width, height)), np.uint8)
        self._check(self._lib.hdb_derive_edges(self._ctx, _addr(info), n, width, height, _addr(v), _addr(h)))
        return v, h

    def derive_edges_device(self, d_info, n_frames, width, height, d_vert_edges, d_hor_edges, stream=None):
        self._check(self._lib.hdb_derive_edges_device(self._ctx, _addr(d_info), n_frames, width, height,
                                                      _addr(d_vert_edges), _addr(d_hor_edges), stream))

    # --- flagged general forms: FLAG_EDGE_MAPS, FLAG_NV12 ---
    def deblock_batch_ex(self, frames, width, height, qps, vert=None, hor=None, flags=0, out=None, timing=None,
                         n_frames=None):
        n = n_frames if n_frames is not None else int(frames.shape[0])
        if out is None:
            out = np.empty_like(frames)
        self._check(self._lib.hdb_deblock_batch_ex(self._ctx, _addr(frames), _addr(out), n, width, height,
                                                   _addr(qps), _addr(vert), _addr(hor), flags,
                                                   None if timing is None else C.byref(timing)))
        return out

    def deblock_batch_blocks(self, frames, width, height, qps, info, flags=0, out=None, timing=None, n_frames=None):
        """Host frames + host block data (BLOCK_INFO[n, H/8, W/8]); the edge maps are derived on the device."""
        n = n_frames if n_frames is not None else int(frames.shape[0])
        if out is None:
            out = np.empty_like(frames)
        if isinstance(info, np.ndarray):
            info = np.ascontiguousarray(info, BLOCK_INFO)
        self._check(self._lib.hdb_deblock_batch_blocks(self._ctx, _addr(frames), _addr(out), n, width, height,
                                                       _addr(qps), _addr(info), flags,
                                                       None if timing is None else C.byref(timing)))
        return out

    def deblock_batch_device_ex(self, d_in, d_out, n_frames, width, height, d_qp, d_vert=None, d_hor=None, flags=0,
                                stream=None):
        self._check(self._lib.hdb_deblock_batch_device_ex(self._ctx, _addr(d_in), _addr(d_out), n_frames, width, height,
                                                          _addr(d_qp), _addr(d_vert), _addr(d_hor), flags, stream))

    def deblock_nv12_surfaces_device(self, src, dst, n_frames, width, height, d_qp, d_vert=None, d_hor=None, flags=0,
                                     stream=None):
        self._check(self._lib.hdb_deblock_nv12_surfaces_device(self._ctx, C.byref(src), C.byref(dst), n_frames, width,
                                                               height, _addr(d_qp), _addr(d_vert), _addr(d_hor), flags,
                                                               stream))

    # --- sample adaptive offset (H.265 8.7.3; absent from the reference) ---
    def sao_batch(self, frames, width, height, params, ctb_size=64, out=None, flags=0):
        """Host frames uint8[n, 3wh/2] (uint16 with flags = flag_depth(d)) and
        SAO_PARAMS[n, 3, ceil(h / ctb), ceil(w / ctb)] -> offset frames."""
        n = int(frames.shape[0])
        params = np.ascontiguousarray(params, SAO_PARAMS)
        if out is None:
            out = np.empty_like(frames)
        if flags:
            self._check(self._lib.hdb_sao_batch_ex(self._ctx, _addr(frames), _addr(out), n, width, height, _addr(params), ctb_size, flags))
        else:
            self._check(self._lib.hdb_sao_batch(self._ctx, _addr(frames), _addr(out), n, width, height, _addr(params), ctb_size))
        return out

    def sao_batch_device(self, d_in, d_out, n_frames, width, height, d_params, ctb_size=64, stream=None, flags=0):
        if flags:
            self._check(self._lib.hdb_sao_batch_device_ex(self._ctx, _addr(d_in), _addr(d_out), n_frames, width, height,
                                                          _addr(d_params), ctb_size, flags, stream))
        else:
            self._check(self._lib.hdb_sao_batch_device(self._ctx, _addr(d_in), _addr(d_out), n_frames, width, height,
                                                       _addr(d_params), ctb_size, stream))

    def deblock_sao_batch_device(self, d_in, d_out, n_frames, width, height, d_qp, d_vert, d_hor, d_params, ctb_size=64,
                                 flags=0, stream=None):
        self._check(self._lib.hdb_deblock_sao_batch_device(self._ctx, _addr(d_in), _addr(d_out), n_frames, width, height,
                                                           _addr(d_qp), _addr(d_vert), _addr(d_hor), flags, _addr(d_params),
                                                           ctb_size, stream))

    def deblock_surfaces_device(self, src, dst, n_frames, width, height, d_qp, d_vert_bs=None, d_hor_bs=None,
                                stream=None):
        """src / dst: Surfaces (device planes with pitches); enqueues on `stream` and returns."""
        self._check(self._lib.hdb_deblock_surfaces_device(self._ctx, C.byref(src), C.byref(dst), n_frames, width, height,
                                                          _addr(d_qp), _addr(d_vert_bs), _addr(d_hor_bs), stream))


class MultiDeblocker:
    """Several GPUs of one box, one process (hdb_multi)."""

    def __init__(self, devices=None, n_devices=0):
        self._lib = load_library()
        self._m = C.c_void_p()
        arr = None
        if devices is not None:
            n_devices = len(devices)
            arr = (C.c_int * n_devices)(*devices)
        rc = self._lib.hdb_multi_create(C.byref(self._m), arr, n_devices)
        if rc:
            raise DeblockError(rc, self._lib.hdb_last_error(None).decode())

    def close(self):
        if self._m:
            self._lib.hdb_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_count(self):
        return int(self._lib.hdb_multi_device_count(self._m))

    def _check(self, rc):
        if rc:
            raise DeblockError(rc, self._lib.hdb_multi_last_error(self._m).decode())

    def deblock_batch(self, frames, width, height, qps, vert_bs=None, hor_bs=None, out=None, timing=None):
        n = int(frames.shape[0])
        if out is None:
            out = np.empty_like(frames)
        self._check(self._lib.hdb_multi_deblock_batch(self._m, _addr(frames), _addr(out), n, width, height,
                                                      _addr(qps), _addr(vert_bs), _addr(hor_bs),
                                                      None if timing is None else C.byref(timing)))
        return out

    def deblock_batch_ex(self, frames, width, height, qps, vert=None, hor=None, flags=0, out=None, timing=None):
        n = int(frames.shape[0])
        if out is None:
            out = np.empty_like(frames)
        self._check(self._lib.hdb_multi_deblock_batch_ex(self._m, _addr(frames), _addr(out), n, width, height,
                                                         _addr(qps), _addr(vert), _addr(hor), flags,
                                                         None if timing is None else C.byref(timing)))
        return out

    def deblock_frame_strips(self, yuv, width, height, qp, vert_bs=None, hor_bs=None, out=None, timing=None, flags=0):
        if out is None:
            out = np.empty_like(yuv)
        v = None if vert_bs is None else np.ascontiguousarray(vert_bs, np.uint8)
        h = None if hor_bs is None else np.ascontiguousarray(hor_bs, np.uint8)
        self._check(self._lib.hdb_multi_deblock_frame_strips_ex(self._m, _addr(yuv), _addr(out), width, height, qp,
                                                                _addr(v), 0 if v is None else v.size,
                                                                _addr(h), 0 if h is None else h.size, flags,
                                                                None if timing is None else C.byref(timing)))
        return out

    def deblock_stream_strips(self, frames, width, height, qps, vert=None, hor=None, flags=0, out=None, timing=None,
                              n_frames=None):
        """A stream of frames, each split into CTU-row strips over the devices; pipelined across frames."""
        n = n_frames if n_frames is not None else int(frames.shape[0])
        if out is None:
            out = np.empty_like(frames)
        self._check(self._lib.hdb_multi_deblock_stream_strips(self._m, _addr(frames), _addr(out), n, width, height,
                                                              _addr(qps), _addr(vert), _addr(hor), flags,
                                                              None if timing is None else C.byref(timing)))
        return out

    def strip_layout(self, height):
        """[(first luma row, luma rows)] per device in use."""
        n = self.device_count()
        a, b = (C.c_int * n)(), (C.c_int * n)()
        g = int(self._lib.hdb_multi_strip_layout(self._m, height, a, b))
        return [(int(a[i]), int(b[i])) for i in range(g)]

    def strip_halo_bytes(self, width, height):
        return int(self._lib.hdb_multi_strip_halo_bytes(self._m, width, height))

    def deblock_strips_device(self, strips_in, strips_out, width, height, qp, d_vert=None, d_hor=None, flags=0,
                              streams=None, timing=None):
        """strips_in / strips_out: per device (y, u, v) device tensors or addresses holding that device's rows."""
        g = len(strips_in)
        mk = lambda strips: (StripPlanes * g)(*[StripPlanes(_addr(y), _addr(u), _addr(v)) for (y, u, v) in strips])
        arr = lambda xs: None if xs is None else (C.c_void_p * g)(*[_addr(x) for x in xs])
        self._check(self._lib.hdb_multi_deblock_strips_device(self._m, mk(strips_in), mk(strips_out), width, height, qp,
                                                              arr(d_vert), arr(d_hor), flags, arr(streams),
                                                              None if timing is None else C.byref(timing)))
