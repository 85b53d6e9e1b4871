"""CPU-side checks of the C-ABI library: it loads, exports exactly what include/*.h declares,
its host-side data-contract helpers agree with the oracle, and it refuses to run without a GPU
(no CPU fallback).  No compute call is made here."""
import ctypes as C
import os

import numpy as np
import pytest

import gpu_video_codec_b200 as hdb
from oracle import oracle_py as orc


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(hdb.library_path()):
        hdb.build_library()
    return hdb.load_library()


def test_exports_every_declared_symbol(lib):
    declared = hdb.header_symbols()
    exported = hdb.exported_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    for s in declared:
        assert getattr(lib, s) is not None
    # nothing CPU-filter-like leaks out of the product: only hdb_* is exported, none of the oracle's names
    assert not [s for s in exported if "oracle" in s or s.startswith("hdbo_")]


def test_product_does_not_link_oracle():
    import subprocess
    needed = subprocess.run(["ldd", hdb.library_path()], capture_output=True, text=True).stdout
    assert "oracle" not in needed and "ref_deblock" not in needed


def test_sizes_and_defaults_match_oracle(lib):
    for (w, h) in [(8, 8), (352, 288), (768, 576), (1920, 1080), (3840, 2160), (7680, 4320), (24, 40)]:
        assert hdb.num_vert_bs(w, h) == orc.num_vert_bs(w, h)
        assert hdb.num_hor_bs(w, h) == orc.num_hor_bs(w, h)
        v, hh = hdb.default_boundary_strength(w, h)
        ov, oh = orc.default_bs(w, h)
        assert np.array_equal(v, ov) and np.array_equal(hh, oh)
    with pytest.raises(hdb.DeblockError):
        hdb.default_boundary_strength(12, 8)


def test_tables_match_oracle(lib):
    for qp in list(range(0, 64)) + [200, 2 ** 31 + 5]:
        assert lib.hdb_beta(qp) == orc.port().hdbo_beta(qp)
        assert lib.hdb_tc(qp) == orc.port().hdbo_tc(qp)


def test_status_strings(lib):
    assert lib.hdb_status_string(0) == b"ok"
    assert lib.hdb_status_string(2) == b"Incorrect file size"
    assert b"multiplier of sample block size" in lib.hdb_status_string(3)
    assert lib.hdb_status_string(4) == b"Incorrect size of input boundary strenght array"


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = C.c_void_p()
    rc = lib.hdb_create(C.byref(ctx), 0)
    assert rc == 5 and not ctx.value  # HDB_ERR_CUDA
    assert b"CUDA" in lib.hdb_last_error(None) or b"device" in lib.hdb_last_error(None)
    with pytest.raises(hdb.DeblockError):
        hdb.Deblocker(0)


def test_struct_layouts_and_constants_match_the_header(tmp_path):
    """The ctypes mirrors in binding.py against the C compiler's view of include/hevc_deblock_b200.h."""
    import subprocess
    from gpu_video_codec_b200 import binding
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "hevc_deblock_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %u %u %u\\n", sizeof(hdb_timing), sizeof(hdb_surfaces),'
                   ' sizeof(hdb_nv12_surfaces), sizeof(hdb_block_info), sizeof(hdb_device_info), offsetof(hdb_timing, kernel_launches),'
                   ' offsetof(hdb_nv12_surfaces, frame_stride_uv), offsetof(hdb_block_info, flags), offsetof(hdb_device_info, total_global_mem),'
                   ' HDB_FLAG_EDGE_MAPS, HDB_FLAG_NV12, (unsigned)HDB_EDGE(2, 37)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(binding.Timing), C.sizeof(binding.Surfaces), C.sizeof(binding.Nv12Surfaces), binding.BLOCK_INFO.itemsize,
            C.sizeof(binding.DeviceInfo), binding.Timing.kernel_launches.offset, binding.Nv12Surfaces.frame_stride_uv.offset,
            binding.BLOCK_INFO.fields["flags"][1], binding.DeviceInfo.total_global_mem.offset,
            binding.FLAG_EDGE_MAPS, binding.FLAG_NV12, int(binding.edge_bytes(2, 37))]
    assert got == want, (got, want)
    # this round's additions: strip planes, SAO parameters, the depth flags
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "hevc_deblock_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %u %u\\n", sizeof(hdb_strip_planes), offsetof(hdb_strip_planes, v),'
                   ' sizeof(hdb_sao_params), offsetof(hdb_sao_params, offset), HDB_FLAG_SAMPLES16, HDB_FLAG_DEPTH(10)); return 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(binding.StripPlanes), binding.StripPlanes.v.offset, binding.SAO_PARAMS.itemsize,
            binding.SAO_PARAMS.fields["offset"][1], binding.FLAG_SAMPLES16, binding.flag_depth(10)]
    assert got == want, (got, want)
    assert (binding.BLK_INTRA, binding.BLK_CBF, binding.BLK_TU_LEFT, binding.BLK_TU_TOP, binding.BLK_PU_LEFT, binding.BLK_PU_TOP) \
        == (1, 2, 4, 8, 16, 32)
