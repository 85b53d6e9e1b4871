"""Host-side logic of the N>1 path on CPU: world_size-2 gloo processes exercise the frame
sharding rule and the max-over-ranks reduction bench.py uses; the strip partition rule of the
multi-GPU front-end is checked against its specification.  No GPU."""
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = bench.shard(total, world, rank)
    mine = torch.zeros(total, dtype=torch.int32)
    mine[first:first + n] = 1
    dist.all_reduce(mine)                     # every frame owned by exactly one rank
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # timing = max over ranks
    dist.barrier()
    if rank == 0:
        q.put((bool((mine == 1).all()), float(t.item())))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [1024, 21, 2, 1])
def test_frame_sharding_gloo_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + total % 97
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    covered, tmax = q.get(timeout=10)
    assert covered and tmax == 11.0


def test_shard_rule_matches_library_rule():
    # hdb_multi_deblock_batch: n/g frames each, remainder to the first devices, contiguous
    for total in (1, 7, 21, 1024, 1025):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                first, n = bench.shard(total, world, r)
                assert n in (total // world, total // world + 1)
                seen += list(range(first, first + n))
            assert seen == list(range(total))


def test_reference_arm_prints_only_on_rank0():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_bytes_match_survey_table():
    # SURVEY.md section 8(d)
    assert bench.alg_bytes_per_frame(352, 288) == 308208
    assert bench.alg_bytes_per_frame(768, 576) == 1344636
    assert bench.alg_bytes_per_frame(1920, 1080) == 6302302
    assert bench.alg_bytes_per_frame(3840, 2160) == 25208325
    assert bench.alg_bytes_per_frame(7680, 4320) == 100831050


def test_workload_frames_depend_on_the_global_frame_number_only():
    """Every rank of our arm and the reference arm must draw the same frames whatever their shard (bench.py:
    workload_frames): any range equals the matching slice of a larger one."""
    import torch
    w, h = 64, 48
    full = bench.workload_frames(torch, 0, 40, w, h, 123, "cpu")
    for first, n in [(0, 16), (16, 16), (24, 13), (5, 35), (39, 1)]:
        part = bench.workload_frames(torch, first, n, w, h, 123, "cpu")
        for a, b in zip(part, full):
            assert torch.equal(a, b[first:first + n]), (first, n)
    other = bench.workload_frames(torch, 0, 16, w, h, 124, "cpu")
    assert not torch.equal(other[0], full[0][:16])
    # the QP specials ride on the global frame number too
    q = bench.workload_frames(torch, 64, 8, w, h, 123, "cpu")[1]
    assert int(q[3]) == 12 and int(q[5]) == 17 and int(q[6]) == 60


def test_both_arms_print_the_same_config():
    for name in ("4k_batch1024", "1080p_batch256", "4k_yuv420p10_512"):
        assert bench.workload_config(name, 8, "cuda") == bench.workload_config(name, 8, "cuda")
        c = bench.workload_config(name, 4, "cuda")
        assert c["frames_per_gpu"] * 4 >= c["frames"] and c["workload"] == name


def test_exact_block_shapes_divide_the_block_counts():
    for nbx, nby in [(481, 271), (241, 136), (97, 73), (49, 37), (45, 37), (961, 541)]:
        dx, dy = bench.exact_block_shape(nbx, nby)
        assert nbx % dx == 0 and nby % dy == 0 and dx * dy <= 1024
