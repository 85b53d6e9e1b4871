"""gpu_video_codec_b200 -- B200 (sm_100a) HEVC-style deblocking filter.

The product is the C-ABI shared library ``libhevc_deblock_b200.so`` built from
``csrc/`` (CUDA kernels + C++ host code, boundary declared in
``include/hevc_deblock_b200.h``).  This Python package is only the thin ctypes
binding the tests and ``bench.py`` call it through; it adds no computation of its
own and has no CPU path: without the built library, or without a compute
capability 10.x device, every filtering call raises.
"""
from .binding import (  # noqa: F401
    Deblocker,
    DeblockError,
    FLAG_EDGE_MAPS,
    FLAG_NV12,
    FLAG_SAMPLES16,
    MultiDeblocker,
    Nv12Surfaces,
    SAO_PARAMS,
    StripPlanes,
    Surfaces,
    Timing,
    build_library,
    default_boundary_strength,
    exported_symbols,
    flag_depth,
    header_symbols,
    library_path,
    load_library,
    num_hor_bs,
    num_vert_bs,
)
