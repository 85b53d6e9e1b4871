"""Zero-copy interop with a decoder-side producer.  No hardware decoder can be
driven here (no bitstream source, no Video Codec SDK headers), so a stand-in plays NVDEC's part exactly where it
touches this library: ANOTHER PROCESS owns a pool of pitched NV12 surfaces in device memory -- what
cuvidMapVideoFrame hands out: a device pointer and a pitch -- writes decoded pictures into it, and exports it over
CUDA IPC.  This process imports the allocation and runs the deblocking filter on it IN PLACE through
hdb_deblock_nv12_surfaces_device; the producer then reads its own memory back.  No sample crosses the host link
on the way through the filter: the PCIe leg the reference's README shows at 84-88 % of the with-copy time is gone."""
import numpy as np
import pytest

import gpu_video_codec_b200 as hdb
from oracle import oracle_py as orc
from tests import synth

pytestmark = pytest.mark.gpu

W, H, N, PITCH = 1920, 1080, 4, 2048          # decoder-style pitch: wider than the picture, a multiple of 256


def _producer(q_out, q_in):
    """The stand-in decoder: allocates the surface pool, 'decodes' N pictures into it, shares it, waits for the
    filter, and reports what its memory holds afterwards."""
    import torch
    torch.cuda.set_device(0)
    rng = np.random.default_rng(123)
    frames = np.stack([orc.yv12_to_nv12(synth.make_frame(rng, W, H, "blocky"), W, H) for _ in range(N)])
    pool = torch.full((N, H * 3 // 2, PITCH), 0x5A, dtype=torch.uint8, device="cuda")       # Y rows, then UV rows, pitched
    pool[:, :, :W] = torch.from_numpy(frames.reshape(N, H * 3 // 2, W)).cuda()
    torch.cuda.synchronize()
    q_out.put((pool, frames))                 # the tensor travels as a CUDA IPC handle
    q_in.get()                                # the consumer has filtered in place
    torch.cuda.synchronize()
    q_out.put((pool[:, :, :W].cpu().numpy().reshape(N, -1), bool((pool[:, :, W:] == 0x5A).all())))
    q_in.get()


def test_filter_in_place_on_surfaces_owned_by_another_process():
    import torch
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q_out, q_in = ctx.Queue(), ctx.Queue()
    p = ctx.Process(target=_producer, args=(q_out, q_in))
    p.start()
    try:
        pool, frames = q_out.get(timeout=120)
        assert pool.is_cuda and pool.shape == (N, H * 3 // 2, PITCH)
        dev = hdb.Deblocker(0)
        qps = np.array([37, 22, 45, 30], np.uint32)
        d_qp = torch.from_numpy(qps.astype(np.int32)).cuda()
        base = pool.data_ptr()
        surf = hdb.Nv12Surfaces(base, base + H * PITCH, PITCH, PITCH, (H * 3 // 2) * PITCH, (H * 3 // 2) * PITCH)
        launches0 = dev.kernel_launch_count()
        dev.deblock_nv12_surfaces_device(surf, surf, N, W, H, d_qp, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert dev.kernel_launch_count() == launches0 + 1          # one kernel, no copies of any kind
        q_in.put("filtered")
        got, padding_untouched = q_out.get(timeout=120)
        q_in.put("done")
        want = np.stack([orc.deblock_frame_nv12(frames[f], W, H, int(qps[f])) for f in range(N)])
        assert np.array_equal(got, want)
        assert padding_untouched
        del pool
        dev.close()
    finally:
        p.join(timeout=60)
        if p.is_alive():
            p.kill()
