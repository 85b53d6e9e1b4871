#!/usr/bin/env python
"""bench.py -- the measurement contract.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" is one pass of the deblocking path over the whole batch of the workload.  Default
workload = BASELINE.json configs[3], the configuration the metric is quoted on: a synthetic
3840x2160 YV12 batch of 1024 frames with random per-frame QP and luma boundary-strength maps,
sharded by contiguous frame ranges across the N GPUs (total work fixed => strong scaling; no
data-path collective, NCCL only carries the barrier and the max-over-ranks of the timings).

  value : Mpixel/s, whole job, frames resident in HBM when the timed region starts (one kernel
          launch per step per GPU), CUDA events on the launching stream, max over ranks.
  e2e   : Mpixel/s through the public C ABI (hdb_deblock_batch) with pinned HOST buffers:
          H2D copy, kernels and D2H copy of every step inside the timed region.
  roofline   : algorithmic bytes of one launch / its mean duration vs the measured HBM peak.
  cpu_baseline: the reference's own CPU filter (oracle/_ref, else the oracle port) on the box's
          host cores, rank 0 at N=1, on a bounded sample of the same frames.

--impl reference times only that CPU implementation (the reference arm).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, total frames, description)  -- BASELINE.json configs[3], [2], [1]
    "4k_batch1024": (3840, 2160, 1024, "synthetic 3840x2160 YV12 batch of 1024 frames, random QP/BS maps, sharded by frame"),
    "1080p_batch256": (1920, 1080, 256, "synthetic 1920x1080 YV12 batch of 256 frames, random QP/BS maps"),
    "image2_single": (768, 576, 1, "single 768x576 frame, QP 30, default BS (latency case)"),
    # BASELINE.json configs[4]: one process drives all GPUs (hdb_multi_deblock_frame_strips); not a torchrun workload
    "8k_strips": (7680, 4320, 16, "stream of single 7680x4320 frames, each split into CTU-row strips across the GPUs, P2P halos"),
    # the same kind of batch stored as NV12 (Y + interleaved UV), one GPU
    "4k_nv12_512": (3840, 2160, 512, "synthetic 3840x2160 NV12 batch of 512 frames, random QP/BS maps"),
    # 10-bit samples in 16-bit containers, planar (yuv420p10le); one GPU
    "4k_yuv420p10_512": (3840, 2160, 512, "synthetic 3840x2160 planar 10-bit (16-bit containers) batch of 512 frames, random QP/BS maps"),
    # the same in the layout a hardware decoder writes: P010 (Y + interleaved UV, samples in the high 10 bits); one GPU
    "4k_p010_512": (3840, 2160, 512, "synthetic 3840x2160 P010 (semi-planar 10-bit, high-aligned) batch of 512 frames, random QP/BS maps"),
    # the in-loop pair, deblocking then sample adaptive offset; one GPU
    "4k_sao256": (3840, 2160, 256, "synthetic 3840x2160 YV12 batch of 256 frames: deblocking filter, then sample adaptive offset with random per-CTB parameters (a third off, a third band, a third edge)"),
    # block data -> edge maps (bs | qp << 2) -> per-edge-QP filter; one GPU
    "4k_edges512": (3840, 2160, 512, "synthetic 3840x2160 YV12 batch of 512 frames, per-8x8-block decoder data -> derived edge maps -> per-edge-QP filter"),
}
METRIC = "deblocked Mpixel/s"
UNIT = "Mpixel/s"


def alg_bytes_per_frame(w, h, sample_bytes=1):
    """every sample read once and written once + every BS byte read once."""
    cw, ch = w // 2, h // 2
    nv, nh = (w // 8 + 1) * h // 8, (h // 8 + 1) * w // 8
    ncv, nch = (cw // 8 + 1) * ch // 8, (ch // 8 + 1) * cw // 8
    return 3 * w * h * sample_bytes + nv + nh + ncv + nch


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload, frames_in_launch):
    """(DRAM bytes of one launch over `frames_in_launch` frames, where the figure comes from).  STATIC: ncu
    cannot run inside the bench, so this is dram__bytes_read.sum + dram__bytes_write.sum per frame from the
    committed ncu --set full capture of this workload's kernel (profiles/traffic.json names the capture),
    scaled by the frames in the launch; (None, reason) without one."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            t = json.load(open(p))
            per_frame = t.get(workload + "_per_frame")
            if per_frame is None:
                return None, "no ncu capture of this workload"
            return int(per_frame * frames_in_launch), "static: %s, per frame x frames in the launch" % (t.get("source") or "profiles/traffic.json")
        except Exception:
            return None, "profiles/traffic.json unreadable"
    return None, "no profiles/traffic.json"


class ClockSampler:
    """SM clock and throttle reasons of one GPU while the timed region runs: NVML polled from a
    thread every ~2 ms (nvidia-smi -lms as the fallback)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.samples = []
        self.stop_flag = False
        self.paused = False     # samples are only taken inside the timed regions
        self.interval = 0.002   # seconds between NVML polls
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            except Exception:
                pass
            handle = None
            if uuid:
                try:
                    handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
                except Exception:
                    handle = None
            if handle is None:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml = (pynvml, handle)
            self.thr = threading.Thread(target=self._poll, daemon=True)
            self.thr.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nvml
        while not self.stop_flag:
            if self.paused:
                time.sleep(0.002)
                continue
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.samples.append((sm, reasons, pw))
            except Exception:
                pass
            time.sleep(self.interval)

    def _read(self):
        for ln in self.proc.stdout:
            if not self.paused:
                self.lines.append(ln.strip())

    def pause(self):
        self.paused = True

    def resume(self, interval=0.002):
        self.interval = interval
        self.paused = False

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.thr.join(timeout=1)
            nv, h = self.nvml
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
            try:
                mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            except Exception:
                mx = None
            bits = 0
            for _, r, _ in self.samples:
                bits |= int(r)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                     "hw_power_brake_slowdown": 0x80}
            reasons = sorted(k for k, v in names.items() if bits & v)
            return {"sm_mhz": float(np.median([s[0] for s in self.samples])), "sm_max_mhz": mx,
                    "power_w_max": round(max(s[2] for s in self.samples), 1), "samples": len(self.samples),
                    "reasons": reasons, "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def shard(total, world, rank):
    """Contiguous frame range of `rank`: remainder to the first ranks (same rule as hdb_multi_deblock_batch)."""
    base, rem = divmod(total, world)
    n = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n


GEN_CHUNK = 16      # frames per generator chunk: chunk c of a workload is a function of (workload seed, c) alone


def workload_frames(torch, first, n, w, h, seed, device="cuda"):
    """Frames [first, first + n) of a workload ('blocky' recipe of tests/synth.py), with their QPs and luma BS
    maps.  Content depends on the GLOBAL frame number only (chunks of GEN_CHUNK frames, each with its own
    generator seed), so every rank of our arm and the reference arm draw the same frames whatever their
    shard.  device "cuda" (Philox) and "cpu" (Mersenne) give different streams: both arms use "cuda" when a
    GPU is visible, and the line says which was used."""
    fb = 3 * w * h // 2
    nv, nh = (w // 8 + 1) * h // 8, (h // 8 + 1) * w // 8
    frames = torch.empty((n, fb), dtype=torch.uint8, device=device)
    qp = torch.empty((n,), dtype=torch.int32, device=device)
    vbs = torch.empty((n, nv), dtype=torch.uint8, device=device)
    hbs = torch.empty((n, nh), dtype=torch.uint8, device=device)
    c0, c1 = first // GEN_CHUNK, (first + n + GEN_CHUNK - 1) // GEN_CHUNK
    for c in range(c0, c1):
        g = torch.Generator(device=device).manual_seed(seed * 100003 + c)
        m = GEN_CHUNK
        chunk = torch.empty((m, fb), dtype=torch.uint8, device=device)
        off = 0
        for (pw, ph) in ((w, h), (w // 2, h // 2), (w // 2, h // 2)):
            xx = torch.arange(pw, device=device, dtype=torch.float32)[None, None, :] / pw
            yy = torch.arange(ph, device=device, dtype=torch.float32)[None, :, None] / ph
            ramp = 96.0 + 64.0 * xx + 32.0 * yy
            bh, bw = (ph + 7) // 8, (pw + 7) // 8
            dc = torch.randint(-12, 13, (m, bh, bw), device=device, generator=g, dtype=torch.int16)
            dc = dc.repeat_interleave(8, 1).repeat_interleave(8, 2)[:, :ph, :pw]
            noise = torch.randint(-2, 3, (m, ph, pw), device=device, generator=g, dtype=torch.int16)
            chunk[:, off:off + pw * ph] = (ramp + dc + noise).clamp_(0, 255).to(torch.uint8).reshape(m, -1)
            off += pw * ph
        cq = torch.randint(18, 52, (m,), device=device, generator=g, dtype=torch.int32)
        gidx = torch.arange(c * GEN_CHUNK, (c + 1) * GEN_CHUNK, device=device)
        # a few frames on the no-op (QP <= 17) and clamp (QP > 51) ranges
        cq = torch.where(gidx % 64 == 3, torch.full_like(cq, 12), cq)
        cq = torch.where(gidx % 64 == 5, torch.full_like(cq, 17), cq)
        cq = torch.where(gidx % 64 == 6, torch.full_like(cq, 60), cq)
        cv = torch.randint(0, 3, (m, nv), device=device, generator=g, dtype=torch.uint8)
        ch = torch.randint(0, 3, (m, nh), device=device, generator=g, dtype=torch.uint8)
        lo, hi = max(first, c * GEN_CHUNK), min(first + n, (c + 1) * GEN_CHUNK)
        a, b = lo - c * GEN_CHUNK, hi - c * GEN_CHUNK
        frames[lo - first:hi - first] = chunk[a:b]
        qp[lo - first:hi - first] = cq[a:b]
        vbs[lo - first:hi - first] = cv[a:b]
        hbs[lo - first:hi - first] = ch[a:b]
    return frames, qp, vbs, hbs


WORKLOAD_SEED = {"4k_p010_512": 5000, "4k_sao256": 6000, "4k_yuv420p10_512": 5000, "4k_batch1024": 1000, "1080p_batch256": 2000, "4k_nv12_512": 3000, "4k_edges512": 4242, "8k_strips": 77}


def workload_config(name, world, gen_device):
    """The `config` object of a line: facts about the workload only, identical in both arms."""
    w, h, total, desc = WORKLOADS[name]
    fb = 3 * w * h // 2 * (2 if ("p10" in name or "p010" in name) else 1)
    per = (total + world - 1) // world
    return {"workload": name, "description": desc, "width": w, "height": h, "frames": total,
            "frames_per_gpu": per, "sharding": "contiguous frame ranges, no collective",
            "l2": "inputs larger than L2 (%.1f GB in + %.1f GB out per GPU)" % (per * fb / 1e9, per * fb / 1e9),
            "generator": "bench.workload_frames, seed %d, torch %s generator" % (WORKLOAD_SEED.get(name, 0), gen_device)}


# ------------------------------------------------------------------------------------------------
# the reference arm / cpu_baseline: the reference's own CPU filter on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(frames, w, h, qps, vbs, hbs, cores):
    """Times the CPU filter on `frames` (numpy).  Returns (Mpixel/s best, detail dict)."""
    from oracle import oracle_py as orc
    n = frames.shape[0]
    mpix = n * w * h / 1e6
    detail = {}
    if orc.have_ref():
        kind = "reference"
        _, t1 = orc.ref_deblock_batch(frames[:max(1, n // 4)], w, h, qps, vbs, hbs, 1, 1)
        detail["single_thread_mpixel_s"] = round(max(1, n // 4) * w * h / 1e6 / t1, 1)
        _, t_omp = orc.ref_deblock_batch(frames, w, h, qps, vbs, hbs, 1, cores)          # the reference's OpenMP path
        detail["openmp_%d_threads_mpixel_s" % cores] = round(mpix / t_omp, 1)
        _, t_fp = orc.ref_deblock_batch(frames, w, h, qps, vbs, hbs, min(cores, n), 1)   # one frame per thread
        detail["frame_parallel_%d_threads_mpixel_s" % min(cores, n)] = round(mpix / t_fp, 1)
        best = mpix / min(t_omp, t_fp)
        used = cores
    else:
        kind = "port"
        t0 = time.perf_counter()
        orc.deblock_batch(frames, w, h, qps, vbs, hbs)
        t = time.perf_counter() - t0
        detail["single_thread_mpixel_s"] = round(mpix / t, 1)
        best, used = mpix / t, 1
    return best, kind, used, detail


def run_reference_arm(args, rank):
    if rank != 0:
        return
    w, h, total, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    n = max(1, min(total, args.cpu_sample or 256))
    gen_device = "cpu"
    if args.workload == "image2_single":
        from tests import common
        frames = common.fixture("image2_768x576.yuv")[None, :]
        qps, vbs, hbs = np.array([30], np.uint32), None, None
    else:
        # the first n frames of the very workload our arm times (same generator, same seed)
        import torch
        gen_device = "cuda" if torch.cuda.is_available() else "cpu"
        t_f, t_q, t_v, t_h = workload_frames(torch, 0, n, w, h, WORKLOAD_SEED.get(args.workload, 0), gen_device)
        frames, qps = t_f.cpu().numpy(), t_q.cpu().numpy().astype(np.uint32)
        vbs, hbs = t_v.cpu().numpy(), t_h.cpu().numpy()
        del t_f, t_q, t_v, t_h
    vals, kind, used, detail = [], None, None, None
    for i in range(args.warmup + args.steps):
        v, kind, used, detail = cpu_reference_run(frames, w, h, qps, vbs, hbs, cores)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.median(vals))
    sample = "first %d of the %d frames of the workload per step, best of OpenMP-in-frame and frame-parallel on %d threads" % (n, total, used)
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(n * w * h / 1e3 / value, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args.workload, max(args.gpus, 1), gen_device),
            "timing": "wall clock around the reference's DeblockingFilter calls on the host cores",
            "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": used, "kind": kind, "sample": sample, "detail": detail},
            "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def exact_block_shape(nbx, nby, limit=1024):
    """A thread-block shape (dx, dy) with dx | nbx and dy | nby: the reference's kernels have no bounds guard
    (gpu.cu:540-541), so its rounded-up grid (gpu.cu:1268) must not round.  The most threads under `limit`,
    wider first."""
    best = (1, 1)
    for dx in range(1, min(nbx, limit) + 1):
        if nbx % dx:
            continue
        for dy in range(1, min(nby, limit // dx) + 1):
            if nby % dy == 0 and (dx * dy, dx) > (best[0] * best[1], best[0]):
                best = (dx, dy)
    return best


def run_reference_gpu_arm(args, rank):
    """Courtesy datapoint: the reference's OWN CUDA kernels (hevc_deblocking_filter_gpu.cu,
    unmodified, compiled for sm_100a by oracle/Makefile into oracle/_ref/ref_gpu_demo) on one frame of the workload,
    timed by the reference's own protocol (gpu.cu:1246-1303: wall clock around synchronous copies and three launches).
    Not an oracle and not bit-exact: its kernels differ from its CPU filter."""
    if rank != 0:
        return
    import re
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_gpu_demo")
    w, h, total, desc = WORKLOADS[args.workload]
    if not os.path.exists(exe):
        print(json.dumps({"impl": "reference_gpu", "unavailable": "oracle/_ref/ref_gpu_demo has not been built (needs /root/reference and nvcc)"}), flush=True)
        return
    import torch
    from oracle import oracle_py as orc
    gen_device = "cuda" if torch.cuda.is_available() else "cpu"
    if args.workload == "image2_single":
        from tests import common
        frame, qp = common.fixture("image2_768x576.yuv"), 30
    else:
        t_f, t_q, _, _ = workload_frames(torch, 0, 1, w, h, WORKLOAD_SEED.get(args.workload, 0), gen_device)
        frame, qp = t_f[0].cpu().numpy(), int(t_q[0])
    d1 = exact_block_shape((w + 8) // 8, (h + 8) // 8)
    d2 = exact_block_shape((w // 2 + 8) // 8, (h // 2 + 8) // 8)
    # One ExecuteGpu call per process, as main.cu makes it (main.cu:138): repeated calls inside one process corrupt the
    # reference's state at 4K (the second call took 77 ms, the third died with SIGSEGV; profiles/r02*_refgpu*).
    k, c, got, log = [], [], None, []
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, "in.yuv"), os.path.join(tmp, "out.yuv")
        frame.tofile(fin)
        for i in range(args.warmup + args.steps):
            p = subprocess.run([exe, fin, fout, str(w), str(h), str(qp), str(d1[0]), str(d1[1]), str(d2[0]), str(d2[1]), "1"],
                               capture_output=True, text=True, timeout=300)
            log.append("rc=%s\n%s%s" % (p.returncode, p.stdout, p.stderr))
            if p.returncode != 0 or not os.path.exists(fout):
                print(json.dumps({"impl": "reference_gpu", "unavailable": "ref_gpu_demo exited with %s: %s" % (
                    p.returncode, (p.stderr.strip() or p.stdout.strip())[-200:])}), flush=True)
                return
            if i >= args.warmup:
                k += [float(x) for x in re.findall(r"Execution Time without copy on GPU: ([0-9.eE+-]+)s", p.stdout)]
                c += [float(x) for x in re.findall(r"Execution Time with copy on GPU: ([0-9.eE+-]+)s", p.stdout)]
            got = np.fromfile(fout, np.uint8)
    if os.environ.get("HDB_REFGPU_LOG"):
        open(os.environ["HDB_REFGPU_LOG"], "w").write("\n".join(log))
    want = orc.deblock_frame(frame, w, h, qp)
    diff = int((got != want).sum()) if got.size == want.size else -1
    mpix = w * h / 1e6
    value, e2e = mpix / float(np.median(k)), mpix / float(np.median(c))
    line = {"impl": "reference_gpu", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(float(np.median(k)) * 1e3, 4), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args.workload, 1, gen_device),
            "timing": "the reference's own protocol (gpu.cu:1246-1303): host wall clock around three launches + cudaDeviceSynchronize, one frame per call",
            "label": "courtesy datapoint: the reference's own CUDA kernels recompiled for sm_100a; NOT an oracle and NOT bit-exact",
            "sample": "frame 0 of the workload, default boundary strengths (ExecuteGpu has no setter), one frame per ExecuteGpu call, one call per process (as main.cu:138)",
            "block_shapes": {"luma": list(d1), "chroma": list(d2), "why": "exact divisors of the block counts: the reference's kernels have no bounds guard"},
            "bytes_differing_from_the_reference_cpu_filter": diff, "frame_bytes": int(want.size),
            "e2e": {"value": round(e2e, 1), "unit": UNIT, "h2d_bytes_per_step": int(want.size), "d2h_bytes_per_step": int(want.size),
                    "what": "the reference's 'Execution Time with copy on GPU' (padded planes + 4 BS arrays in, padded planes out, synchronous cudaMemcpy)"},
            "gpu_launches": 3 * args.steps}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def quiet_stdout():
    """Send everything libraries write to fd 1 (e.g. NCCL's version banner) to stderr; the one JSON
    line is written to the original stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


class Dist:
    """torch.distributed plumbing of one rank: NCCL carries the barrier and the max-over-ranks of the timings
    (nothing else -- the path has no data-path collective); a gloo group lets ranks idle on the CPU while rank 0
    drives all devices for the strip-split sub-record."""

    def __init__(self, torch, world, local_rank):
        self.torch, self.world, self.dist, self.cpu_group = torch, world, None, None
        if world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist
            self.cpu_group = dist.new_group(backend="gloo")

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def cpu_barrier(self):
        if self.dist is not None:
            self.dist.barrier(group=self.cpu_group)

    def max(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_ints(self, xs):
        if self.dist is None:
            return [int(x) for x in xs]
        t = self.torch.tensor(list(xs), dtype=self.torch.int64, device="cuda")
        self.dist.all_reduce(t)
        return [int(x) for x in t.tolist()]

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def measure_workload(args, torch, hdb, dev, dd, rank, world, workload, steps, warmup, sampler=None, full_parity=True):
    """Kernel-only, e2e, link probe and parity of one workload on this rank's shard.  Returns the pieces of the
    JSON line (rank 0) -- the caller assembles them."""
    w, h, total, desc = WORKLOADS[workload]
    first, n = shard(total, world, rank)
    fb = 3 * w * h // 2
    stream = torch.cuda.current_stream().cuda_stream
    if workload == "image2_single":
        from tests import common
        frames = torch.from_numpy(common.fixture("image2_768x576.yuv")[None, :].copy()).cuda()
        qp = torch.tensor([30], dtype=torch.int32, device="cuda")
        vbs = hbs = None
        n = 1
    else:
        frames, qp, vbs, hbs = workload_frames(torch, first, n, w, h, WORKLOAD_SEED.get(workload, 0), "cuda")
    nv12 = "nv12" in workload
    depth = 10 if ("p10" in workload or "p010" in workload) else 0
    p010 = "p010" in workload
    flags = hdb.FLAG_NV12 if nv12 else (hdb.flag_depth(depth) | (hdb.FLAG_NV12 if p010 else 0) if depth else 0)
    ysz, csz = w * h, (w // 2) * (h // 2)
    if depth:
        # the same frames scaled to the depth, the low bits filled at random: int16 tensors (torch has no uint16 arithmetic)
        g16 = torch.Generator(device="cuda").manual_seed(WORKLOAD_SEED[workload] + 7 * first)
        frames = (frames.to(torch.int16) << (depth - 8)) | torch.randint(0, 1 << (depth - 8), frames.shape, device="cuda", generator=g16, dtype=torch.int16)
        fb *= 2
        if p010:      # interleave the chroma planes and move the samples to the high bits
            uv = torch.stack([frames[:, ysz:ysz + csz], frames[:, ysz + csz:]], dim=2).reshape(n, -1)
            frames = (torch.cat([frames[:, :ysz], uv], dim=1) << 6).contiguous()

    def to_nv12(t):      # [k, 3wh/2] planar -> NV12, on whatever device t lives
        r = t.clone()
        r[:, ysz:] = torch.stack([t[:, ysz:ysz + csz], t[:, ysz + csz:]], dim=2).reshape(t.shape[0], -1)
        return r

    def to_yv12(t):
        r = t.clone()
        uv = t[:, ysz:].reshape(t.shape[0], csz, 2)
        r[:, ysz:ysz + csz] = uv[:, :, 0]
        r[:, ysz + csz:] = uv[:, :, 1]
        return r

    if nv12:
        frames = to_nv12(frames)
    out = torch.empty_like(frames)

    # ---- kernel-only: frames resident in HBM ------------------------------------------------
    def step():
        dev.deblock_batch_device_ex(frames, out, n, w, h, qp, vbs, hbs, flags=flags, stream=stream)

    for _ in range(max(warmup, 3)):
        step()
    dd.barrier()
    if sampler is not None:
        sampler.resume(0.002)
    launches0 = dev.kernel_launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    dd.barrier()
    evs[0].record()
    for i in range(steps):
        step()
        evs[i + 1].record()
    dd.barrier()
    if sampler is not None:
        sampler.pause()             # resumed for the e2e timed region; both regions feed the clocks record
    kernel_launches = dev.kernel_launch_count() - launches0
    total_ms = dd.max(evs[0].elapsed_time(evs[-1]))
    launch_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    mean_launch_ms = dd.max(float(np.mean(launch_ms)))
    ms_per_step = total_ms / steps
    value = total * w * h / 1e6 / (ms_per_step / 1e3)

    # ---- parity of what was just timed: a strided sample of this rank's frames vs the oracle ---
    parity = None
    if rank == 0:
        from oracle import oracle_py as orc
        idx = sorted(set([0, n // 3, (2 * n) // 3, n - 1] + ([3, 5, 6] if n > 6 else [])))
        sel = torch.tensor(idx, device="cuda")
        # NV12: the checkers take planar frames; the two layouts are byte permutations of each other
        got = (to_yv12(out[sel]) if nv12 else out[sel]).cpu().numpy()
        src = (to_yv12(frames[sel]) if nv12 else frames[sel]).cpu().numpy()
        q = qp[sel].cpu().numpy().astype(np.uint32)
        v = None if vbs is None else vbs[sel].cpu().numpy()
        hh = None if hbs is None else hbs[sel].cpu().numpy()
        if depth:
            got, src = got.view(np.uint16), src.view(np.uint16)
            chk16 = (lambda f, qq, vv, hv: orc.deblock_frame_p010(f, w, h, qq, vv, hv)) if p010 else \
                (lambda f, qq, vv, hv: orc.deblock_frame16(f, w, h, qq, vv, hv, bit_depth=depth))
            want = np.stack([chk16(src[i], int(q[i]), v[i], hh[i]) for i in range(len(idx))])
        else:
            want = orc.deblock_batch(src, w, h, q, v, hh)
        ok = bool(np.array_equal(got, want))
        parity = {"bit_exact": ok, "frames_checked": len(idx), "changed_fraction": round(float((got != src).mean()), 4)}
        if not ok:
            raise SystemExit("PARITY FAILURE: GPU output differs from the oracle; no number is reported")

    # ---- end to end: pinned host buffers through hdb_deblock_batch -----------------------------
    h_in = torch.empty(frames.shape, dtype=frames.dtype).pin_memory()
    h_out = torch.empty(frames.shape, dtype=frames.dtype).pin_memory()
    h_in.copy_(frames)
    h_qp = qp.cpu().pin_memory()
    h_v = None if vbs is None else vbs.cpu().pin_memory()
    h_h = None if hbs is None else hbs.cpu().pin_memory()
    tm = hdb.Timing()

    def e2e_step():
        dev.deblock_batch_ex(h_in, w, h, h_qp, h_v, h_h, flags=flags, out=h_out, timing=tm, n_frames=n)

    for _ in range(max(1, min(warmup, 2))):
        e2e_step()
    dd.barrier()
    l0 = dev.kernel_launch_count()
    if sampler is not None:
        sampler.resume(0.05)   # a sparse poll: NVML queries take driver locks the copy submissions also need
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()     # returns after the last D2H copy has completed
    torch.cuda.synchronize()
    e2e_wall = time.perf_counter() - t0
    if sampler is not None:
        sampler.pause()
    e2e_s = dd.max(e2e_wall)
    e2e_launches = dev.kernel_launch_count() - l0
    e2e_value = total * w * h / 1e6 / (e2e_s / steps)
    h2d = int(tm.h2d_bytes)
    d2h = int(tm.d2h_bytes)

    # ---- link probe: what the host links of this box carry when every rank copies the SAME pinned buffers
    #      both ways at once with plain cudaMemcpyAsync in 128 MB pieces and nothing else runs ----
    link = None
    if n * fb >= (64 << 20):
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        piece = max(1, (128 << 20) // fb)

        def probe():
            for f0 in range(0, n, piece):
                f1 = min(n, f0 + piece)
                with torch.cuda.stream(s_up):
                    frames[f0:f1].copy_(h_in[f0:f1], non_blocking=True)      # the bytes already there: nothing is disturbed
                with torch.cuda.stream(s_dn):
                    h_out[f0:f1].copy_(out[f0:f1], non_blocking=True)        # likewise
            s_up.synchronize(); s_dn.synchronize()

        probe()
        dd.barrier()
        reps = max(1, min(steps, 3))
        t0 = time.perf_counter()
        for _ in range(reps):
            probe()
        probe_s = dd.max((time.perf_counter() - t0) / reps)
        link = {"ms_per_step": round(probe_s * 1e3, 3), "gb_s_per_direction_all_gpus": round(world * n * fb / probe_s / 1e9, 2),
                "what": "all ranks at once: cudaMemcpyAsync H2D + D2H of the e2e leg's pinned buffers, 128 MB pieces, two streams, no kernel"}

    if rank == 0 and full_parity:
        # every frame of this rank's shard, as delivered to the host by the e2e path, against the CPU
        # checker (the unmodified reference on all host cores when oracle/_ref is there, else the port)
        from oracle import oracle_py as orc
        cores = os.cpu_count() or 1
        np_in, np_out = (to_yv12(h_in).numpy(), to_yv12(h_out).numpy()) if nv12 else (h_in.numpy(), h_out.numpy())
        np_qp = h_qp.numpy().astype(np.uint32)
        np_v = None if h_v is None else h_v.numpy()
        np_h = None if h_h is None else h_h.numpy()
        bad, t_chk = 0, time.perf_counter()
        e2e_checked = n
        if depth:
            # the 16-bit checker is the single-threaded port: a strided 32 frames of the e2e output, the rest through
            # the device-resident output, which is compared with the e2e output frame for frame below
            chk = list(range(0, n, max(1, n // 32)))
            e2e_checked = len(chk)
            a_in, a_out = h_in.numpy().view(np.uint16), h_out.numpy().view(np.uint16)
            for f in chk:
                wnt = orc.deblock_frame_p010(a_in[f], w, h, int(np_qp[f]), np_v[f], np_h[f]) if p010 else \
                    orc.deblock_frame16(a_in[f], w, h, int(np_qp[f]), np_v[f], np_h[f], bit_depth=depth)
                bad += 0 if np.array_equal(a_out[f], wnt) else 1
        for f0 in range(0, 0 if depth else n, 64):
            f1 = min(n, f0 + 64)
            vv = None if np_v is None else np_v[f0:f1]
            hv = None if np_h is None else np_h[f0:f1]
            if orc.have_ref():
                want, _ = orc.ref_deblock_batch(np_in[f0:f1], w, h, np_qp[f0:f1], vv, hv, min(cores, f1 - f0), 1)
            else:
                want = orc.deblock_batch(np_in[f0:f1], w, h, np_qp[f0:f1], vv, hv)
            bad += sum(0 if np.array_equal(np_out[f0 + i], want[i]) else 1 for i in range(f1 - f0))
        # and the output of the kernel-only (timed) launch must be that same output, frame for frame
        for f0 in range(0, n, 64):
            f1 = min(n, f0 + 64)
            if not torch.equal(out[f0:f1], h_out[f0:f1].cuda(non_blocking=False)):
                raise SystemExit("PARITY FAILURE: device-resident output differs from the e2e output in frames %d..%d" % (f0, f1))
        parity["e2e_frames_checked"] = e2e_checked
        parity["e2e_frames_differing"] = bad
        parity["checker"] = "port at depth %d (oracle/deblock_oracle.c: hdbo_deblock_frame16; unpinned against the reference above depth 8)" % depth if depth \
            else ("reference (oracle/_ref)" if orc.have_ref() else "port (oracle/deblock_oracle.c)")
        parity["check_seconds"] = round(time.perf_counter() - t_chk, 1)
        parity["bit_exact"] = parity["bit_exact"] and bad == 0
        if bad:
            raise SystemExit("PARITY FAILURE: %d of %d e2e frames differ from the CPU checker; no number is reported" % (bad, n))
    h2d, d2h, kernel_launches, e2e_launches = dd.sum_ints([h2d, d2h, kernel_launches, e2e_launches])

    res = None
    if rank == 0:
        peak, peak_src = measured_peak()
        alg = alg_bytes_per_frame(w, h, 2 if depth else 1) * n          # one launch processes this rank's n frames
        achieved = alg / (mean_launch_ms / 1e3) / 1e9
        traffic, traffic_src = ncu_traffic(workload, n)
        e2e = {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": round(e2e_s / steps * 1e3, 3), "launches_per_step": e2e_launches // max(steps, 1),
               "api": "hdb_deblock_batch%s (pinned host buffers, 3-stream pipeline)" % ("_ex, HDB_FLAG_NV12" if nv12 else ("_ex, HDB_FLAG_DEPTH(%d)%s" % (depth, " | HDB_FLAG_NV12" if p010 else "") if depth else ""))}
        if link is not None:
            e2e["frac_of_link"] = round(link["ms_per_step"] / e2e["ms_per_step"], 4)
        res = {"value": round(value, 1), "ms_per_step": round(ms_per_step, 4),
               "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                            "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                            "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                            "kernel": "DeblockTiled16Kernel<%d, 0, %d>" % (depth, int(p010)) if depth else ("DeblockTiledKernel<0, 1, 0>" if nv12 else "DeblockTiledKernel<0, 0, 0>"),
                            "mean_launch_ms": round(mean_launch_ms, 4)},
               "e2e": e2e, "link_probe": link, "gpu_launches": kernel_launches, "parity": parity}
        if world == 1 and not args.no_cpu and full_parity:
            cores = os.cpu_count() or 1
            ns = max(1, min(n, args.cpu_sample or 256))
            sel = torch.arange(ns, device="cuda")
            src = (to_yv12(frames[sel]) if nv12 else frames[sel]).cpu().numpy()
            q = qp[sel].cpu().numpy().astype(np.uint32)
            v = None if vbs is None else vbs[sel].cpu().numpy()
            hh = None if hbs is None else hbs[sel].cpu().numpy()
            if depth:
                src16 = src.view(np.uint16)[:8]
                t0 = time.perf_counter()
                for i in range(src16.shape[0]):
                    (orc.deblock_frame_p010(src16[i], w, h, int(q[i]), v[i], hh[i]) if p010 else
                     orc.deblock_frame16(src16[i], w, h, int(q[i]), v[i], hh[i], bit_depth=depth))
                dt = time.perf_counter() - t0
                best, kind, used, detail = src16.shape[0] * w * h / 1e6 / dt, "port", 1, {"frames": int(src16.shape[0])}
                ns = int(src16.shape[0])
            else:
                best, kind, used, detail = cpu_reference_run(src, w, h, q, v, hh, cores)
            res["cpu_baseline"] = {"value": round(best, 1), "unit": UNIT, "cores": used, "kind": kind,
                                   "sample": "first %d of the %d frames of the workload, best of OpenMP-in-frame and "
                                             "frame-parallel on %d threads" % (ns, total, used), "detail": detail}
    del frames, out, h_in, h_out
    torch.cuda.empty_cache()
    return res


def run_ours(args, rank, world, local_rank):
    import torch
    import gpu_video_codec_b200 as hdb
    quiet_stdout()
    torch.cuda.set_device(local_rank)
    dd = Dist(torch, world, local_rank)
    dev = hdb.Deblocker(local_rank)
    w, h, total, desc = WORKLOADS[args.workload]
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler is not None:
        sampler.paused = True
        sampler.start()
    res = measure_workload(args, torch, hdb, dev, dd, rank, world, args.workload, args.steps, args.warmup, sampler)
    clocks = sampler.stop() if sampler is not None else None
    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u16" if ("p10" in args.workload or "p010" in args.workload) else "u8", "data": "synthetic",
                "config": workload_config(args.workload, world, "cuda"),
                "timing": "CUDA events on the launching stream, max over ranks",
                "roofline": res["roofline"], "e2e": res["e2e"], "link_probe": res["link_probe"],
                "gpu_launches": res["gpu_launches"], "clocks": clocks, "parity": res["parity"]}
        if "cpu_baseline" in res:
            line["cpu_baseline"] = res["cpu_baseline"]
    # ---- the other single-GPU configurations of BASELINE.json (configs[2], configs[1]) ride along at N = 1 ----
    if world == 1 and args.workload == "4k_batch1024" and not args.no_sub:
        sub = {}
        for name, st in (("1080p_batch256", max(3, args.steps)), ("image2_single", 50)):
            r = measure_workload(args, torch, hdb, dev, dd, rank, world, name, st, 3, None, full_parity=(name != "image2_single"))
            r.pop("link_probe", None)
            r.pop("cpu_baseline", None)
            sub[name] = dict(r, config=workload_config(name, 1, "cuda"), steps=st)
        line["other_configs"] = sub
    # ---- BASELINE.json configs[4] rides along at N > 1: rank 0 alone drives all N devices, the others idle ----
    if world > 1 and args.workload == "4k_batch1024" and not args.no_sub:
        dev.close()
        dev = None
        torch.cuda.empty_cache()
        dd.barrier()
        if rank == 0:
            try:
                line["strips_8k"] = strips_record(args, torch, hdb, list(range(world)), steps=max(2, min(args.steps, 3)))
            except SystemExit:
                raise
            except Exception as e:      # the main line stands on its own
                line["strips_8k"] = {"error": "%s: %s" % (type(e).__name__, e)}
        dd.cpu_barrier()
    if rank == 0:
        emit(line)
    if dev is not None:
        dev.close()
    dd.close()


def nvlink_counters(devices):
    """Sum of the NVLink data bytes (tx, rx) of the devices from NVML's field values, or None."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        tx = rx = 0
        for d in devices:
            hnd = nv.nvmlDeviceGetHandleByIndex(d)
            vals = nv.nvmlDeviceGetFieldValues(hnd, [(nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, 0xffffffff),
                                                     (nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xffffffff)])
            if vals[0].nvmlReturn != 0 or vals[1].nvmlReturn != 0:
                return None
            tx += int(vals[0].value.ullVal) * 1024      # KiB
            rx += int(vals[1].value.ullVal) * 1024
        return tx, rx
    except Exception:
        pass
    try:        # the same counters through nvidia-smi: "Link n: Data Tx: x KiB" / "Data Rx: y KiB" per link
        import re
        tx = rx = 0
        for d in devices:
            out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(d)], capture_output=True, text=True, timeout=20).stdout
            t = [int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out)]
            r = [int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out)]
            if not t or not r:
                return None
            tx += sum(t) * 1024
            rx += sum(r) * 1024
        return tx, rx
    except Exception:
        return None


def strips_record(args, torch, hdb, devices, steps):
    """BASELINE configs[4]: a stream of 8K frames, each cut into CTU-row strips over `devices` (one process
    drives them all).  Legs: the pipelined stream (hdb_multi_deblock_stream_strips), frame-at-a-time calls
    (hdb_multi_deblock_frame_strips: the latency of one frame), the same stream through ONE GPU unsplit
    (hdb_deblock_batch), and strips resident in each GPU's memory (kernel + NVLink only).  Parity: every
    distinct frame of every leg against the unsplit CPU checker."""
    from oracle import oracle_py as orc
    from tests import synth
    w, h, total, desc = WORKLOADS["8k_strips"]
    fb = 3 * w * h // 2
    m = hdb.MultiDeblocker(devices)
    one = hdb.Deblocker(devices[0])
    distinct = 4
    frames, qps, v, hh = synth.make_batch(WORKLOAD_SEED["8k_strips"], distinct, w, h)
    idx = [i % distinct for i in range(total)]
    h_in = torch.from_numpy(frames[idx]).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    h_qp = torch.from_numpy(qps[idx].astype(np.int32)).pin_memory()
    h_v, h_h = torch.from_numpy(v[idx]).pin_memory(), torch.from_numpy(hh[idx]).pin_memory()
    cores = os.cpu_count() or 1
    want, _ = orc.ref_deblock_batch(frames, w, h, qps, v, hh, min(cores, distinct), 1) if orc.have_ref() else (orc.deblock_batch(frames, w, h, qps, v, hh), 0)
    tm = hdb.Timing()
    np_in, np_out = h_in.numpy(), h_out.numpy()

    def check(name):
        for f in range(total):
            if not np.array_equal(np_out[f], want[idx[f]]):
                raise SystemExit("PARITY FAILURE (%s): frame %d differs from the CPU checker" % (name, f))
        np_out[:] = 0

    def leg_stream():
        m.deblock_stream_strips(h_in, w, h, h_qp, h_v, h_h, out=h_out, timing=tm, n_frames=total)
        return tm.kernel_ms

    def leg_frames():
        k = 0.0
        for f in range(total):
            m.deblock_frame_strips(np_in[f], w, h, int(qps[idx[f]]), v[idx[f]], hh[idx[f]], out=np_out[f], timing=tm)
            k += tm.kernel_ms
        return k

    def leg_unsplit():
        one.deblock_batch(h_in, w, h, h_qp, h_v, h_h, out=h_out, timing=tm, n_frames=total)
        return tm.kernel_ms

    res = {}
    nvl0 = nvlink_counters(sorted(set(devices)))
    for name, fn in (("stream", leg_stream), ("frame_at_a_time", leg_frames), ("unsplit_1gpu_stream", leg_unsplit)):
        for _ in range(2):
            fn()
        check(name)
        t0 = time.perf_counter()
        kms = 0.0
        for _ in range(steps):
            kms += fn()
        dt = time.perf_counter() - t0
        check(name)
        res[name] = {"mpixel_s": round(steps * total * w * h / 1e6 / dt, 1), "ms_per_frame": round(dt * 1e3 / (steps * total), 4),
                     "kernel_ms_per_frame": round(kms / (steps * total), 4)}
    nvl1 = nvlink_counters(sorted(set(devices)))
    # strips resident in each GPU's memory: kernel + NVLink only
    lay = m.strip_layout(h)
    cw = w // 2
    Y, U, V = (frames[0][:w * h].reshape(h, w), frames[0][w * h:w * h * 5 // 4].reshape(h // 2, cw),
               frames[0][w * h * 5 // 4:].reshape(h // 2, cw))
    ins, outs, dv, dh = [], [], [], []
    for d, (y0, rows) in enumerate(lay):
        where = "cuda:%d" % devices[d]
        ins.append(tuple(torch.from_numpy(np.ascontiguousarray(p)).to(where) for p in (Y[y0:y0 + rows], U[y0 // 2:(y0 + rows) // 2], V[y0 // 2:(y0 + rows) // 2])))
        outs.append(tuple(torch.zeros_like(p) for p in ins[-1]))
        dv.append(torch.from_numpy(v[0]).to(where))
        dh.append(torch.from_numpy(hh[0]).to(where))
    for d in set(devices):
        torch.cuda.synchronize(d)
    ks = []
    for i in range(3 + 20):
        m.deblock_strips_device(ins, outs, w, h, int(qps[0]), dv, dh, timing=tm)
        if i >= 3:
            ks.append(tm.kernel_ms)
    got = np.concatenate([np.concatenate([o[p].cpu().numpy().reshape(-1) for o in outs]) for p in range(3)])
    if not np.array_equal(got, want[0]):
        raise SystemExit("PARITY FAILURE (device-resident strips): differs from the CPU checker")
    halo = m.strip_halo_bytes(w, h)
    res["device_resident"] = {"kernel_ms_per_frame_slowest_gpu": round(float(np.median(ks)), 4),
                              "mpixel_s": round(w * h / 1e6 / (float(np.median(ks)) / 1e3), 1),
                              "nvlink_halo_bytes_per_frame": halo,
                              "what": "strips already in each GPU's HBM; one kernel per GPU, halo rows read by peer loads inside it"}
    if nvl0 is None or nvl1 is None:
        res["nvlink_counters"] = {"unavailable": "NVML and nvidia-smi report N/A for the NVLink data counters on this box; nvlink_halo_bytes_per_frame is the analytic figure (4 rows of each plane per cut and direction)"}
    else:
        frames_run = 3 * (2 + steps) * total
        res["nvlink_counters"] = {"tx_bytes": nvl1[0] - nvl0[0], "rx_bytes": nvl1[1] - nvl0[1], "frames": frames_run,
                                  "expected_halo_bytes": halo * (2 * (2 + steps) * total) if len(set(devices)) > 1 else 0,
                                  "source": "NVML NVLINK_THROUGHPUT_DATA_TX/RX, all links, all devices, over the three host legs"}
    res.update({"devices": devices, "strips": len(lay), "frames_in_stream": total, "width": w, "height": h,
                "h2d_bytes_per_frame": fb + v.shape[1] + hh.shape[1], "d2h_bytes_per_frame": fb,
                "parity": {"bit_exact": True, "frames_checked_per_leg": total, "checker": "reference (oracle/_ref)" if orc.have_ref() else "port"},
                "speedup_stream_vs_unsplit": round(res["unsplit_1gpu_stream"]["ms_per_frame"] / res["stream"]["ms_per_frame"], 3)})
    m.close()
    one.close()
    return res


def run_strips(args):
    """`--workload 8k_strips`: the strip-split record as a line of its own (one process drives --gpus devices;
    logical shards on device 0 when the box has fewer)."""
    import torch
    import gpu_video_codec_b200 as hdb
    quiet_stdout()
    w, h, total, desc = WORKLOADS[args.workload]
    have = torch.cuda.device_count()
    devices = list(range(args.gpus)) if have >= args.gpus else [i % have for i in range(args.gpus)]
    res = strips_record(args, torch, hdb, devices, steps=args.steps)
    fb = 3 * w * h // 2
    line = {"metric": METRIC, "value": res["stream"]["mpixel_s"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": 2, "ms_per_step": round(res["stream"]["ms_per_frame"] * total, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args.workload, 1, "numpy (tests/synth.py)"),
            "timing": "wall clock around the calls (host buffers in and out)",
            "e2e": {"value": res["stream"]["mpixel_s"], "unit": UNIT, "h2d_bytes_per_step": total * res["h2d_bytes_per_frame"],
                    "d2h_bytes_per_step": total * fb, "ms_per_step": round(res["stream"]["ms_per_frame"] * total, 3),
                    "api": "hdb_multi_deblock_stream_strips"},
            "strips_8k": res, "gpu_launches": args.steps * total * res["strips"], "parity": res["parity"]}
    emit(line)


def device_block_info(torch, n, w, h, seed, device="cuda"):
    """hdb_block_info[n, h/8, w/8] on the device as int32[n, h/8 * w/8, 3] words (deblock_core.cuh: BlockWords):
    16x16 coding units with their own motion, QP, coded flag; a fifth intra; a third bi-predicted."""
    g = torch.Generator(device=device).manual_seed(seed)
    h8, w8 = h // 8, w // 8
    ch, cw = (h8 + 1) // 2, (w8 + 1) // 2
    R = lambda lo, hi, shape: torch.randint(lo, hi, shape, device=device, generator=g, dtype=torch.int32)
    up = lambda a: a.repeat_interleave(2, 1).repeat_interleave(2, 2)[:, :h8, :w8]
    cand = R(-24, 25, (n, 6, 4))
    pick = R(0, 6, (n, ch, cw)).long()
    mv = torch.gather(cand, 1, pick.view(n, -1, 1).expand(-1, -1, 4)).view(n, ch, cw, 4)
    mv = mv + (R(0, 10, (n, ch, cw, 1)) < 3) * R(-5, 6, (n, ch, cw, 4))
    ref0 = R(0, 2, (n, ch, cw))
    bi = R(0, 3, (n, ch, cw)) == 0
    ref1 = torch.where(bi, R(0, 3, (n, ch, cw)), torch.full_like(ref0, -1))
    intra = R(0, 5, (n, ch, cw)) == 0
    cbf = R(0, 2, (n, ch, cw)) == 1
    split = R(0, 4, (n, ch, cw)) == 0
    qp = (R(24, 40, (n, 1, 1)) + R(-6, 7, (n, ch, cw))).clamp_(0, 51)
    ref0 = torch.where(intra, torch.full_like(ref0, -1), ref0)
    ref1 = torch.where(intra, torch.full_like(ref1, -1), ref1)
    yy = torch.arange(h8, device=device).view(1, h8, 1)
    xx = torch.arange(w8, device=device).view(1, 1, w8)
    sp = up(split.int()).bool()
    flags = up(intra.int()) * 1 + up(cbf.int()) * 2 + ((xx % 2 == 0) | sp).int() * 4 + ((yy % 2 == 0) | sp).int() * 8 \
        + (xx % 2 == 0).int() * 16 + (yy % 2 == 0).int() * 32
    m = [up(mv[..., i]) & 0xffff for i in range(4)]
    w0 = m[0] | (m[1] << 16)
    w1 = m[2] | (m[3] << 16)
    w2 = (up(ref0) & 0xff) | ((up(ref1) & 0xff) << 8) | (flags << 16) | (up(qp) << 24)
    return torch.stack([w0, w1, w2], dim=-1).reshape(n, h8 * w8, 3).contiguous()


def run_edges(args):
    """SURVEY.md section 8(f) rank 1 on one GPU: DeriveEdgesKernel and DeblockTiledKernel<true> (per-edge QP),
    with the plain kernel on the same frames beside them; e2e through hdb_deblock_batch_edges."""
    import torch
    import gpu_video_codec_b200 as hdb
    from oracle import oracle_py as orc
    quiet_stdout()
    torch.cuda.set_device(0)
    w, h, n, desc = WORKLOADS[args.workload]
    fb = 3 * w * h // 2
    dev = hdb.Deblocker(0)
    stream = torch.cuda.current_stream().cuda_stream
    frames, qp, vbs, hbs = workload_frames(torch, 0, n, w, h, WORKLOAD_SEED[args.workload], "cuda")
    info = device_block_info(torch, n, w, h, seed=99)
    nv, nh = hdb.num_vert_bs(w, h), hdb.num_hor_bs(w, h)
    d_v = torch.empty((n, nv), dtype=torch.uint8, device="cuda")
    d_h = torch.empty((n, nh), dtype=torch.uint8, device="cuda")
    out = torch.empty_like(frames)
    out_plain = torch.empty_like(frames)

    def timed(fn, reps):
        for _ in range(max(args.warmup, 3)):
            fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[-1]) / reps

    sampler = ClockSampler(0)
    sampler.start()
    l0 = dev.kernel_launch_count()
    ms_derive = timed(lambda: dev.derive_edges_device(info, n, w, h, d_v, d_h, stream=stream), args.steps)
    ms_edges = timed(lambda: dev.deblock_batch_edges_device(frames, out, n, w, h, qp, d_v, d_h, stream=stream), args.steps)
    # the plain kernel on the same frames with the derived strengths alone and one QP per frame
    bs_v, bs_h = d_v & 3, d_h & 3
    ms_plain = timed(lambda: dev.deblock_batch_device(frames, out_plain, n, w, h, qp, bs_v, bs_h, stream=stream), args.steps)
    ms_both = timed(lambda: (dev.derive_edges_device(info, n, w, h, d_v, d_h, stream=stream),
                             dev.deblock_batch_edges_device(frames, out, n, w, h, qp, d_v, d_h, stream=stream)), args.steps)
    launches = dev.kernel_launch_count() - l0
    sampler.pause()

    # parity: derivation and filter of a strided sample against the oracle
    idx = [0, n // 3, (2 * n) // 3, n - 1]
    np_info = info[idx].cpu().numpy().view(np.uint8).reshape(len(idx), h // 8, w // 8, 12).copy().view(orc.BLOCK_INFO).reshape(len(idx), h // 8, w // 8)
    for k, f in enumerate(idx):
        ov, oh = orc.derive_edges(np_info[k], w, h)
        if not (np.array_equal(d_v[f].cpu().numpy(), ov) and np.array_equal(d_h[f].cpu().numpy(), oh)):
            raise SystemExit("PARITY FAILURE: derived edge maps of frame %d differ from the oracle" % f)
        want = orc.deblock_frame_edges(frames[f].cpu().numpy(), w, h, int(qp[f]), ov, oh)
        if not np.array_equal(out[f].cpu().numpy(), want):
            raise SystemExit("PARITY FAILURE: per-edge-QP output of frame %d differs from the oracle" % f)
    hist = torch.bincount((d_v[idx].reshape(-1) & 3).long(), minlength=4).tolist()

    # e2e: host buffers through hdb_deblock_batch_edges (edge maps travel with the frames)
    h_in = torch.empty((n, fb), dtype=torch.uint8).pin_memory()
    h_out = torch.empty((n, fb), dtype=torch.uint8).pin_memory()
    h_in.copy_(frames)
    h_qp, h_v, h_h = qp.cpu().pin_memory(), d_v.cpu().pin_memory(), d_h.cpu().pin_memory()
    tm = hdb.Timing()
    e2e = lambda: dev.deblock_batch_edges(h_in, w, h, h_qp, h_v, h_h, out=h_out, timing=tm, n_frames=n)
    e2e()
    sampler.resume(0.05)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if not torch.equal(h_out[::37].cuda(), out[::37]):
        raise SystemExit("PARITY FAILURE: e2e output differs from the device-resident output")
    # the decoder-loop form: block data travels with the frames, the edge maps are derived on the device
    h_info = info.cpu().pin_memory()
    h_out.zero_()
    tm2 = hdb.Timing()
    e2e_blocks = lambda: dev.deblock_batch_blocks(h_in, w, h, h_qp, h_info, out=h_out, timing=tm2, n_frames=n)
    e2e_blocks()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_blocks()
    e2e_blocks_s = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    if not torch.equal(h_out[::37].cuda(), out[::37]):
        raise SystemExit("PARITY FAILURE: output of hdb_deblock_batch_blocks differs from the device-resident output")

    peak, peak_src = measured_peak()
    alg = alg_bytes_per_frame(w, h) * n
    blocks = (w // 8) * (h // 8) * n
    alg_derive = 12 * blocks + (nv + nh) * n
    mpix = n * w * h / 1e6
    line = {"metric": METRIC, "value": round(mpix / (ms_edges / 1e3), 1), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_edges, 4), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "width": w, "height": h, "frames": n,
                       "l2": "inputs larger than L2 (%.1f GB in + %.1f GB out)" % (n * fb / 1e9, n * fb / 1e9),
                       "timing": "CUDA events on the launching stream", "derived_bs_histogram_0123": hist},
            "roofline": {"bound": "hbm", "achieved": round(alg / (ms_edges / 1e3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(alg / (ms_edges / 1e3) / 1e9 / peak, 4), "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg, "kernel": "DeblockTiledKernel<true>", "mean_launch_ms": round(ms_edges, 4)},
            "plain_kernel_same_frames": {"ms": round(ms_plain, 4), "mpixel_s": round(mpix / (ms_plain / 1e3), 1),
                                         "achieved_gbs": round(alg / (ms_plain / 1e3) / 1e9, 1)},
            "derive_kernel": {"ms": round(ms_derive, 4), "algorithmic_bytes": alg_derive,
                              "achieved_gbs": round(alg_derive / (ms_derive / 1e3) / 1e9, 1), "frac": round(alg_derive / (ms_derive / 1e3) / 1e9 / peak, 4)},
            "derive_plus_filter": {"ms": round(ms_both, 4), "mpixel_s": round(mpix / (ms_both / 1e3), 1)},
            "e2e": {"value": round(mpix / e2e_s, 1), "unit": UNIT, "h2d_bytes_per_step": int(tm.h2d_bytes), "d2h_bytes_per_step": int(tm.d2h_bytes),
                    "ms_per_step": round(e2e_s * 1e3, 3), "api": "hdb_deblock_batch_edges (pinned host buffers, 3-stream pipeline)"},
            "e2e_from_block_data": {"value": round(mpix / e2e_blocks_s, 1), "unit": UNIT, "h2d_bytes_per_step": int(tm2.h2d_bytes),
                                    "d2h_bytes_per_step": int(tm2.d2h_bytes), "ms_per_step": round(e2e_blocks_s * 1e3, 3),
                                    "api": "hdb_deblock_batch_blocks (frames + 12 B of block data per 8x8 block in, edge maps derived on the device)"},
            "gpu_launches": launches, "clocks": clocks,
            "parity": {"bit_exact": True, "frames_checked": len(idx), "changed_fraction": round(float((out[idx] != frames[idx]).float().mean()), 4)}}
    emit(line)
    dev.close()


def run_sao(args):
    """SURVEY.md section 8(f) rank 3 on one GPU: SaoKernel alone, DeblockTiledKernel alone and the chained pair
    (hdb_deblock_sao_batch_device) on frames resident in HBM; parity of a strided sample against the oracles."""
    import torch
    import gpu_video_codec_b200 as hdb
    from oracle import oracle_py as orc
    quiet_stdout()
    torch.cuda.set_device(0)
    w, h, n, desc = WORKLOADS[args.workload]
    fb, ctb = 3 * w * h // 2, 64
    dev = hdb.Deblocker(0)
    stream = torch.cuda.current_stream().cuda_stream
    frames, qp, vbs, hbs = workload_frames(torch, 0, n, w, h, WORKLOAD_SEED[args.workload], "cuda")
    rows, cols = (h + ctb - 1) // ctb, (w + ctb - 1) // ctb
    g = torch.Generator(device="cuda").manual_seed(17)
    R = lambda lo, hi, shape: torch.randint(lo, hi, shape, device="cuda", generator=g, dtype=torch.int32)
    typ = R(0, 3, (n, 3, rows, cols))
    pos = torch.where(typ == 1, R(0, 32, typ.shape), R(0, 4, typ.shape))
    mag = R(0, 8, typ.shape + (4,))
    sign = torch.where((typ == 2)[..., None], torch.tensor([1, 1, -1, -1], device="cuda", dtype=torch.int32).expand(mag.shape),
                       2 * R(0, 2, mag.shape) - 1)
    params = torch.cat([typ[..., None], pos[..., None], (mag * sign) & 0xff], dim=-1).to(torch.uint8).contiguous()   # [n, 3, rows, cols, 6]
    deb = torch.empty_like(frames)
    out = torch.empty_like(frames)
    out_pair = torch.empty_like(frames)

    def timed(fn, reps):
        for _ in range(max(args.warmup, 3)):
            fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[-1]) / reps

    sampler = ClockSampler(0)
    sampler.start()
    l0 = dev.kernel_launch_count()
    ms_deb = timed(lambda: dev.deblock_batch_device(frames, deb, n, w, h, qp, vbs, hbs, stream=stream), args.steps)
    ms_sao = timed(lambda: dev.sao_batch_device(deb, out, n, w, h, params, ctb, stream=stream), args.steps)
    ms_pair = timed(lambda: dev.deblock_sao_batch_device(frames, out_pair, n, w, h, qp, vbs, hbs, params, ctb, stream=stream), args.steps)
    launches = dev.kernel_launch_count() - l0
    clocks = sampler.stop()
    if not torch.equal(out, out_pair):
        raise SystemExit("PARITY FAILURE: the chained pair differs from the two stages run one after the other")
    idx = [0, n // 3, (2 * n) // 3, n - 1]
    np_par = params[idx].cpu().numpy().reshape(len(idx), 3, rows, cols, 6).copy().view(orc.SAO_PARAMS).reshape(len(idx), 3, rows, cols)
    for k, f in enumerate(idx):
        d = orc.deblock_frame(frames[f].cpu().numpy(), w, h, int(qp[f]), vbs[f].cpu().numpy(), hbs[f].cpu().numpy())
        if not np.array_equal(deb[f].cpu().numpy(), d):
            raise SystemExit("PARITY FAILURE: deblocked frame %d differs from the oracle" % f)
        if not np.array_equal(out[f].cpu().numpy(), orc.sao_frame(d, w, h, np_par[k], ctb)):
            raise SystemExit("PARITY FAILURE: sample adaptive offset of frame %d differs from the oracle" % f)
    peak, peak_src = measured_peak()
    mpix = n * w * h / 1e6
    alg_sao = (2 * fb + 6 * 3 * rows * cols) * n                # every sample read once and written once + the parameters
    alg_deb = alg_bytes_per_frame(w, h) * n
    line = {"metric": METRIC, "value": round(mpix / (ms_pair / 1e3), 1), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_pair, 4), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args.workload, 1, "cuda"),
            "timing": "CUDA events on the launching stream",
            "roofline": {"bound": "hbm", "achieved": round(alg_sao / (ms_sao / 1e3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(alg_sao / (ms_sao / 1e3) / 1e9 / peak, 4), "traffic": None, "traffic_source": "no ncu capture of this workload",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_sao, "kernel": "SaoKernel", "mean_launch_ms": round(ms_sao, 4)},
            "deblock_alone": {"ms": round(ms_deb, 4), "mpixel_s": round(mpix / (ms_deb / 1e3), 1), "achieved_gbs": round(alg_deb / (ms_deb / 1e3) / 1e9, 1)},
            "sao_alone": {"ms": round(ms_sao, 4), "mpixel_s": round(mpix / (ms_sao / 1e3), 1)},
            "deblock_plus_sao": {"ms": round(ms_pair, 4), "mpixel_s": round(mpix / (ms_pair / 1e3), 1), "ratio_to_deblock_alone": round(ms_pair / ms_deb, 3),
                                 "api": "hdb_deblock_sao_batch_device (two kernels on one stream, scratch of the call's own between them)"},
            "gpu_launches": launches, "clocks": clocks,
            "parity": {"bit_exact": True, "frames_checked": len(idx), "changed_fraction_by_sao": round(float((out[idx] != deb[idx]).float().mean()), 4),
                       "checker": "port (oracle/deblock_oracle.c: hdbo_deblock_frame pinned to the reference; hdbo_sao_frame H.265 8.7.3 restated, unpinned: the reference has no SAO)"}}
    if not args.no_sub:
        # the same stage on ten-bit samples: half as many frames in 16-bit containers, the same bytes
        n10, depth = n // 2, 10
        low = torch.randint(0, 4, (n10, frames.shape[1]), device="cuda", generator=g, dtype=torch.int16)
        f10 = (frames[:n10].to(torch.int16) << 2) | low
        deb10, out10, pair10 = torch.empty_like(f10), torch.empty_like(f10), torch.empty_like(f10)
        flags10 = hdb.flag_depth(depth)
        par10 = params[:n10].contiguous()
        ms_deb10 = timed(lambda: dev.deblock_batch_device_ex(f10, deb10, n10, w, h, qp[:n10], vbs[:n10], hbs[:n10], flags=flags10, stream=stream), args.steps)
        ms_sao10 = timed(lambda: dev.sao_batch_device(deb10, out10, n10, w, h, par10, ctb, stream=stream, flags=flags10), args.steps)
        ms_pair10 = timed(lambda: dev.deblock_sao_batch_device(f10, pair10, n10, w, h, qp[:n10], vbs[:n10], hbs[:n10], par10, ctb, flags=flags10, stream=stream), args.steps)
        if not torch.equal(out10, pair10):
            raise SystemExit("PARITY FAILURE: the ten-bit chained pair differs from the two stages run one after the other")
        idx10 = [0, n10 - 1]
        for k, f in enumerate(idx10):
            p10 = par10[f].cpu().numpy().reshape(3, rows, cols, 6).copy().view(orc.SAO_PARAMS).reshape(3, rows, cols)
            d = orc.deblock_frame16(f10[f].cpu().numpy().view(np.uint16), w, h, int(qp[f]), vbs[f].cpu().numpy(), hbs[f].cpu().numpy(), bit_depth=depth)
            if not np.array_equal(out10[f].cpu().numpy().view(np.uint16), orc.sao_frame16(d, w, h, p10, ctb, depth)):
                raise SystemExit("PARITY FAILURE: ten-bit frame %d differs from the oracle" % f)
        alg10 = (2 * 2 * fb + 6 * 3 * rows * cols) * n10
        line["ten_bit"] = {"frames": n10, "kernel": "Sao16Kernel", "sao_alone_ms": round(ms_sao10, 4),
                           "achieved_gbs": round(alg10 / (ms_sao10 / 1e3) / 1e9, 1), "frac": round(alg10 / (ms_sao10 / 1e3) / 1e9 / peak, 4),
                           "deblock_alone_ms": round(ms_deb10, 4), "deblock_plus_sao_ms": round(ms_pair10, 4),
                           "ratio_to_deblock_alone": round(ms_pair10 / ms_deb10, 3),
                           "parity": {"bit_exact": True, "frames_checked": len(idx10),
                                      "checker": "port (hdbo_deblock_frame16 + hdbo_sao_frame16: depth > 8 is defined by the oracle, unpinned)"}}
    emit(line)
    dev.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference_gpu"])
    ap.add_argument("--workload", default="4k_batch1024", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=0, help="frames in the CPU sample (default 256)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-records (other_configs at N = 1, strips_8k at N > 1)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if args.impl == "reference_gpu":
        run_reference_gpu_arm(args, rank)
        return
    if args.workload == "8k_strips":
        if rank == 0:
            run_strips(args)
        return
    if args.workload == "4k_edges512":
        if rank == 0:
            run_edges(args)
        return
    if args.workload == "4k_sao256":
        if rank == 0:
            run_sao(args)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("--gpus %d needs torchrun: python -m torch.distributed.run --nnodes=1 --nproc-per-node %d "
                         "--master-addr 127.0.0.1 --master-port 29500 bench.py --gpus %d" % (args.gpus, args.gpus, args.gpus))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
