#!/usr/bin/env python
"""bench.py — reassigned frames/s at n_fft=4096 (BASELINE.json metric) on 1..8 B200.

Workload (config.workload): BASELINE.json configs[2] — offline single stream, 1 h mono
48 kHz, n_fft=4096, hop=128, fused STFT+reassignment -> points (dt, dk, energy), one stream
per GPU (clip-sharded, weak scaling, NCCL only for the final gather of per-rank summaries).

  value        : frames/s of ems_process_points with PCM resident in HBM (device timed).
  roofline     : the fused STFT+reassignment kernel, B_points = 4*hop + 12*(n_fft/2+1)
                 algorithmic bytes per frame, CUDA events inside the library.
  e2e          : frames/s of ems_process_host — pinned HOST PCM in, u8 colour-index image
                 out to pinned HOST memory (the whole a1-a5 path, copies inside the region).
  cpu_baseline : the float64 C stand-in oracle (whole a1-a5 path, oracle/reassign_oracle.c) on every host
                 thread; the NumPy restatement is timed beside it (cpu_baseline.numpy_oracle).

  batch        : configs[3] — 4096 clips x 60 s, n_fft=4096 hop=256, clip-sharded over the ranks
                 (strong scaling: the same 4096 clips at every N), clips fed to the engine as planar
                 channels in groups, per-clip checksums all_gathered at the end; the u8 images can
                 also be gathered to rank 0 over NCCL (timed separately).
  value_dense / pipeline_u8_dense / *_broadband : the same calls with every bin kept (gate -200 dB)
                 and on a broadband signal (pink noise + harmonic stacks) — the worst cases of the
                 epilogue and of the scatter.

--impl reference times the oracle port (the only CPU implementation that exists: EM-Spec
ships no source, /root/reference/README.md:73) on rank 0 with all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "em-spec_b200"))

SR = 48000
N_FFT = 4096
HOP = 128
METRIC = "reassigned frames/sec @ n_fft=4096"
UNIT = "frames/s"


def b_points(n_fft, hop):
    return 4 * hop + 12 * (n_fft // 2 + 1)


def frame_count(S, n_fft, hop):
    return 0 if S < n_fft else 1 + (S - n_fft) // hop


# ------------------------------------------------------------------ CPU arm (oracle port)
def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _cpu_slice(args):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reassign_oracle as orc
    x, n_fft, hop = args
    prm = orc.Params(n_fft=n_fft, hop=hop)
    grid, idx = orc.process(x, prm, workers=1)
    return idx.shape[0], int(idx.sum())


def _slices(F: int, per: int):
    """Frame ranges of at most `per` frames; each slice carries its n_fft - hop halo of samples."""
    return [(f0, min(F, f0 + per)) for f0 in range(0, F, per)]


def cpu_frames_per_s(seconds_audio: float, cores: int, repeats: int = 1, seed: int = 0, impl: str = "c"):
    """Oracle a1-a5 on `seconds_audio` of the workload signal with `cores` host threads.
    impl "c": oracle/reassign_oracle.c (float64, pthreads), the stream cut into slices of <= 10 s so
    that the float64 intermediates stay near 0.4 GB and are reused from slice to slice; impl "numpy": oracle/reassign_oracle.py, frames
    split over one process per core.  -> (frames/s, frames, wall)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reassign_oracle as orc
    S = int(seconds_audio * SR)
    x = orc.synth_signal(S, SR, seed=seed)
    F = frame_count(S, N_FFT, HOP)
    best = float("inf")
    if impl == "c":
        import c_oracle
        prm = orc.Params(n_fft=N_FFT, hop=HOP)
        parts = [x[f0 * HOP:(f1 - 1) * HOP + N_FFT] for f0, f1 in _slices(F, frame_count(10 * SR, N_FFT, HOP))]
        import numpy as np
        fmax = frame_count(len(parts[0]), N_FFT, HOP)
        out = (np.empty((fmax, prm.n_rows), np.float64), np.empty((fmax, prm.n_rows), np.uint8))
        c_oracle.process(parts[0], prm, threads=cores, out=out)      # load the library, touch the buffers
        for _ in range(repeats):
            t0 = time.perf_counter()
            frames = 0
            for part in parts:
                _, idx = c_oracle.process(part, prm, threads=cores, out=out)
                frames += idx.shape[0]
            best = min(best, time.perf_counter() - t0)
        return frames / best, frames, best
    import multiprocessing as mp
    jobs = [(x[f0 * HOP:(f1 - 1) * HOP + N_FFT], N_FFT, HOP) for f0, f1 in _slices(F, (F + cores - 1) // max(1, cores))]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_slice, jobs[:cores])            # warm the workers (imports, FFT plans)
        for _ in range(repeats):
            t0 = time.perf_counter()
            res = pool.map(_cpu_slice, jobs)
            best = min(best, time.perf_counter() - t0)
    frames = sum(r[0] for r in res)
    return frames / best, frames, best


def cpu_impl_available(impl: str) -> str:
    """The C restatement needs its shared library (prebuilt in oracle/_build/, else gcc): fall back to the NumPy one
    rather than lose the CPU arm."""
    if impl != "c":
        return impl
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import c_oracle
        c_oracle.load()
        return "c"
    except Exception as e:                                   # noqa: BLE001
        print(f"bench: C oracle unavailable ({e}); timing the NumPy oracle instead", file=sys.stderr)
        return "numpy"


CPU_WHAT = {"c": "float64 C stand-in oracle a1-a5 (oracle/reassign_oracle.c, pthreads, one thread per core)",
            "numpy": "float64 NumPy/SciPy stand-in oracle a1-a5, one process per core"}


# ------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0=None, t1=None):
        """Summary of the samples that arrived inside [t0, t1] (all samples if too few did)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for (t, ln) in self.lines if t0 is None or (t0 <= t <= t1 + 0.05)]
        scope = "timed region"
        if len(inside) < 3:
            inside, scope = [ln for (_, ln) in self.lines], "warm-up + timed region (region too short for 3 samples)"
        for ln in inside:
            p = [v.strip() for v in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "scope": scope, "reasons": sorted(reasons)}


# ------------------------------------------------------------------ synthetic stream on device
def synth_device(S: int, seed: int, device):
    """Workload signal of SURVEY.md §8d generated on the device (fp64 phase -> fp32 samples):
    0.5*logchirp(20->20k over the stream) + 0.25*sin(2pi 440 t) + 0.125*sin(2pi 3000.5 t)
    + 1e-3*N(0,1) (device RNG, seed = stream index).  Generated in blocks to bound memory."""
    import math
    import torch
    x = torch.empty(S, dtype=torch.float32, device=device)
    g = torch.Generator(device=device).manual_seed(seed)
    T = S / SR
    r = math.log(20000.0 / 20.0)
    fa = 440.0 + 7.0 * (seed % 64)
    fb = 3000.5 + 11.0 * (seed % 128)
    blk = 1 << 24
    for s0 in range(0, S, blk):
        s1 = min(S, s0 + blk)
        t = torch.arange(s0, s1, device=device, dtype=torch.float64) / SR
        ph = 2 * math.pi * 20.0 * T / r * torch.expm1(r * t / T)
        v = 0.5 * torch.sin(ph) + 0.25 * torch.sin(2 * math.pi * fa * t) + 0.125 * torch.sin(2 * math.pi * fb * t)
        x[s0:s1] = v.float() + 1e-3 * torch.randn(s1 - s0, device=device, generator=g)
    return x


def synth_broadband_device(S: int, seed: int, device):
    """Dense, music-like signal generated on the device (the worst case of the epilogue and the
    scatter): pink noise at -26 dBFS rms (spectral shaping per 2^22-sample block) + three notes
    with 12 harmonics each (1/h roll-off, peaks near -20 dBFS, 5 Hz vibrato).  Almost every bin
    of every frame clears the default -65 dB gate."""
    import math
    import torch
    x = torch.empty(S, dtype=torch.float32, device=device)
    g = torch.Generator(device=device).manual_seed(1000 + seed)
    blk = 1 << 22
    f = torch.fft.rfftfreq(blk, 1.0 / SR).to(device)
    shape = torch.zeros_like(f)
    shape[1:] = f[1:].rsqrt()
    for s0 in range(0, S, blk):
        s1 = min(S, s0 + blk)
        pink = torch.fft.irfft(torch.fft.rfft(torch.randn(blk, device=device, generator=g)) * shape, blk)
        pink = 0.05 * pink / pink.square().mean().sqrt()
        t = torch.arange(s0, s1, device=device, dtype=torch.float64) / SR
        v = torch.zeros(s1 - s0, dtype=torch.float64, device=device)
        for i, f0 in enumerate((110.0, 196.0, 329.63)):
            ph = f0 * t + 0.005 * f0 / 5.0 * torch.sin(2 * math.pi * 5.0 * t + i)
            for hn in range(1, 13):
                v += (0.1 / hn) * torch.sin(2 * math.pi * hn * ph + 0.3 * hn)
        x[s0:s1] = v.float() + pink[: s1 - s0]
    return x


def gpu_numa_cpus(local: int):
    """CPUs local to GPU `local` (sysfs local_cpulist of its PCI function), or None."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        txt = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        node = open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return sorted(cpus), int(node), bdf
    except Exception:
        return None


def numa_bind(local: int, world: int):
    """Pins this rank to its share of the CPUs local to its GPU, so that the pinned host buffers it
    allocates afterwards are first-touched on that NUMA node.  Returns what was done (for the line)."""
    info = gpu_numa_cpus(local)
    if not info:
        return {"bound": False, "why": "no sysfs topology for this GPU"}
    cpus, node, bdf = info
    allowed = sorted(os.sched_getaffinity(0))
    cpus = [c for c in cpus if c in allowed] or allowed
    # ranks whose GPUs share a node split its CPUs
    per = max(1, len(cpus) // max(1, world))
    mine = cpus[(local % max(1, len(cpus) // per)) * per:][:per] or cpus
    try:
        os.sched_setaffinity(0, mine)
    except OSError as e:
        return {"bound": False, "why": str(e), "numa_node": node}
    return {"bound": True, "numa_node": node, "pci": bdf, "cpus": f"{mine[0]}-{mine[-1]}", "n_cpus": len(mine)}


def pcie_ceiling(dev, world, barrier, nbytes_d2h: int, nbytes_h2d: int, dist=None):
    """Bare copies, all ranks at once: D2H of nbytes_d2h into pinned memory while H2D of nbytes_h2d runs
    on a second stream (what one e2e step moves, with no kernels).  -> aggregate GB/s over all ranks
    (bytes of all ranks / slowest rank's time): the ceiling the e2e number is compared with."""
    import torch
    d = torch.empty(nbytes_d2h, dtype=torch.uint8, device=dev)
    h = torch.empty(nbytes_d2h, dtype=torch.uint8, pin_memory=True)
    xd = torch.empty(nbytes_h2d, dtype=torch.uint8, device=dev)
    xh = torch.empty(nbytes_h2d, dtype=torch.uint8, pin_memory=True)
    h.zero_(); xh.zero_()                      # first touch on this rank's NUMA node
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def once():
        with torch.cuda.stream(s1):
            h.copy_(d, non_blocking=True)
        with torch.cuda.stream(s2):
            xd.copy_(xh, non_blocking=True)
    once(); torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    del d, h, xd, xh
    torch.cuda.empty_cache()
    return {"d2h_GBs": world * nbytes_d2h / dt.item() / 1e9, "h2d_GBs": world * nbytes_h2d / dt.item() / 1e9,
            "step_ms": 1e3 * dt.item(),
            "what": "bare concurrent cudaMemcpyAsync D2H + H2D of one e2e step's bytes per rank, all ranks at once, no kernels"}


def kernel_profile():
    """The committed ncu capture of the dominant kernel (profiles/r02_kernel_profile.json, written by
    tools/ncu_summary.py --json): DRAM bytes and executed flops per frame, with the hash of the kernel
    kernel sources (common.cuh, stft_generic.cuh, stft_r16.cuh) it was taken from.  bench.py recomputes the hash: a stale capture is reported as such."""
    import hashlib
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r02_kernel_profile.json")))
    except Exception:
        return None
    hsh = hashlib.sha256()
    csrc = os.path.join(ROOT, "em-spec_b200", "csrc")
    for fn in ("common.cuh", "stft_generic.cuh", "stft_r16.cuh"):      # the sources of the profiled kernel
        hsh.update(open(os.path.join(csrc, fn), "rb").read())
    prof["stale"] = prof.get("csrc_sha16") != hsh.hexdigest()[:16]
    return prof


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------ arms
# stdout carries exactly one JSON line: file descriptor 1 is pointed at stderr for the whole run
# (NCCL's version banner and any other library chatter go there) and the line is written to the
# saved descriptor at the end.
_REAL_STDOUT = None


def claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_json(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    args.cpu_impl = cpu_impl_available(args.cpu_impl)
    sample_s = args.cpu_seconds
    vals = []
    for i in range(args.warmup + args.steps):
        fps, frames, wall = cpu_frames_per_s(sample_s, cores, repeats=1, impl=args.cpu_impl)
        if i >= args.warmup:
            vals.append((fps, wall))
    fps = statistics.mean(v[0] for v in vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * statistics.mean(v[1] for v in vals),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "configs[2] offline single stream 48 kHz mono n_fft=4096 hop=128 "
                               "(bounded sample of the 1 h stream)", "n_fft": N_FFT, "hop": HOP},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"first {sample_s:g} s of the stream "
                                   f"({frame_count(int(sample_s * SR), N_FFT, HOP)} frames) per step, "
                                   + CPU_WHAT[args.cpu_impl]},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "stand-in: EM-Spec ships no runnable source; this is the oracle port",
    }
    emit_json(line)
    return 0


def time_calls(fn, steps, warmup, barrier=None):
    """CUDA-event time of `steps` calls of fn on the current stream after `warmup` calls -> ms per call."""
    import torch
    for _ in range(warmup):
        fn()
    if barrier:
        barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run_batch(args, emspec, rank, world, dev, dist, barrier):
    """configs[3]: `batch_clips` clips x 60 s at 48 kHz, n_fft=4096 hop=256, clip-sharded (strong
    scaling: the same clips at every N).  A rank holds its clips' PCM in HBM and feeds them to the engine
    as planar channels, `batch_group` clips per call; each group's u8 image is reduced to two per-clip
    checksums; one all_gather of the checksums ends the step (the job's only collective).  With
    --batch-gather (default when N > 1) one more step also ships every image to rank 0 over NCCL
    (isend / irecv per group, overlapped with the next group's kernels), timed separately."""
    import hashlib
    import torch
    from emspec.batch import shard_clips
    C, hop = args.batch_clips, 256
    S = 60 * SR
    F, B = frame_count(S, N_FFT, hop), N_FFT // 2 + 1
    lo, hi = shard_clips(C, rank, world)
    n = hi - lo
    G = min(args.batch_group, n)
    while n % G:
        G -= 1
    pcm_all = torch.empty((n, S), dtype=torch.float32, device=dev)
    for c in range(n):
        pcm_all[c] = synth_device(S, lo + c, dev)
    eng = emspec.Engine(n_fft=N_FFT, hop=hop, channels=G)
    eng.use_torch_stream()
    idx = [torch.empty((G, F, B), dtype=torch.uint8, device=dev) for _ in range(2)]
    sums = torch.zeros((n, 2), dtype=torch.int64, device=dev)
    all_sums = [torch.empty_like(sums) for _ in range(world)] if world > 1 else None
    images = None

    def step(gather=False):
        pend = [None, None]
        reqs = []
        for gi, c0 in enumerate(range(0, n, G)):
            buf = idx[gi & 1]
            if pend[gi & 1] is not None:
                pend[gi & 1].wait()                      # the image this buffer held has been sent
            eng.process_grid(pcm_all[c0:c0 + G], out=(None, buf))
            eng.image_summary(buf, out=sums[c0:c0 + G])
            if gather:
                if rank == 0:
                    images[c0:c0 + G].copy_(buf)
                    ops = [dist.P2POp(dist.irecv, images[shard_clips(C, p_, world)[0] + c0:][:G], p_) for p_ in range(1, world)]
                    reqs += dist.batch_isend_irecv(ops)
                else:
                    pend[gi & 1] = dist.batch_isend_irecv([dist.P2POp(dist.isend, buf, 0)])[0]
        for r in reqs + [q for q in pend if q is not None]:
            r.wait()
        if world > 1:
            dist.all_gather(all_sums, sums)

    for _ in range(args.batch_warmup):
        step()
    barrier()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.batch_steps):
        step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = eng.launch_count() - l0
    every = torch.cat(all_sums) if world > 1 else sums
    digest = hashlib.sha256(every.cpu().numpy().tobytes()).hexdigest()[:16]
    res = {"workload": f"configs[3] batch: {C} clips x 60 s 48 kHz, n_fft=4096 hop=256, clip-sharded x{world} "
                       f"({n} clips per rank, {G} per call as planar channels), PCM resident in HBM -> u8 image -> per-clip checksums",
           "value": C * F * args.batch_steps / (ms.item() * 1e-3), "unit": UNIT, "scaling": "strong",
           "steps": args.batch_steps, "warmup": args.batch_warmup, "ms_per_step": ms.item() / args.batch_steps,
           "clips": C, "frames_per_clip": F, "clip_range_rank0": [lo, hi], "clips_per_call": G,
           "gpu_launches": int(launches), "clip_checksums_sha16": digest,
           "checksum_of": "ems_image_summary per clip (sum of the index bytes, position-weighted sum), all clips in clip order: identical at every N",
           "pcm_bytes_per_rank": int(pcm_all.numel() * 4), "scratch_bytes_per_rank": int(eng.scratch_bytes())}
    if world > 1 and not args.no_batch_gather:
        if rank == 0:
            images = torch.empty((C, F, B), dtype=torch.uint8, device=dev)
        step(gather=True)        # warm-up: NCCL sets its peer-to-peer channels up on first use
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        step(gather=True)
        g1.record()
        barrier()
        gms = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        ok = None
        if rank == 0:      # the gathered images carry the checksums their owners computed
            every = torch.cat(all_sums)
            probe = [shard_clips(C, p_, world)[0] for p_ in range(world)] + [C - 1]
            ok = all(int(images[c].sum(dtype=torch.int64)) == int(every[c, 0]) for c in probe)
        recv = (C - n) * F * B
        res["image_gather"] = {"ms_step_with_gather": gms.item(), "ms_step_without": ms.item() / args.batch_steps,
                               "bytes_to_rank0": recv, "GBs_into_rank0": recv / gms.item() / 1e6,
                               "verified": ok,
                               "what": "one more step in which every group's u8 image also goes to rank 0 (NCCL isend / batched irecv, "
                                       "double-buffered, overlapped with the next group's kernels)"}
        images = None
    eng.close()
    del pcm_all, idx
    torch.cuda.empty_cache()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    import emspec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bind = {"bound": False, "why": "--no-bind"} if args.no_bind else numa_bind(local, world)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    S = int(args.seconds * SR)
    F = frame_count(S, N_FFT, HOP)
    B = N_FFT // 2 + 1
    pcm = synth_device(S, seed=rank, device=dev)
    eng = emspec.Engine(n_fft=N_FFT, hop=HOP, noise_gate_db=args.gate_db)
    eng.use_torch_stream()
    out = tuple(torch.empty((1, F, B), dtype=torch.float32, device=dev) for _ in range(3))

    # ---- device-resident: ems_process_points
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        eng.process_points(pcm, out=out)
    barrier()
    t_region0 = time.perf_counter()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.process_points(pcm, out=out)      # one kernel launch per step, nothing else on the stream
    e1.record()
    barrier()
    launches = eng.launch_count() - l0
    my_ms = e0.elapsed_time(e1)
    kern_last_ms = eng.stage_ms(emspec.STAGE_POINTS)   # library's own events, last step (cross-check)
    total_ms = allmax(my_ms)
    clocks = sampler.stop(t_region0, time.perf_counter()) if rank == 0 else None
    frames_all = F * world
    value = frames_all * args.steps / (total_ms * 1e-3)

    # the timed region on this stream is exactly `steps` launches of the fused kernel, so its
    # average launch duration is the region's CUDA-event time / steps (rank 0's own region)
    kern_ms = my_ms / args.steps
    peak, peak_src = peaks()
    achieved = b_points(N_FFT, HOP) * F / (kern_ms * 1e-3) / 1e9
    prof = kernel_profile()

    # ---- the worst cases of the same call: every bin kept, and a broadband signal
    worst = {}
    pcm_bb = None
    if not args.no_dense:
        dense = emspec.Engine(n_fft=N_FFT, hop=HOP, noise_gate_db=-200.0)
        dense.use_torch_stream()
        ms = allmax(time_calls(lambda: dense.process_points(pcm, out=out), max(3, args.steps // 4), 2, barrier))
        worst["value_dense"] = {"value": frames_all / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                                "what": "ems_process_points, same stream, noise gate -200 dB: every bin runs the "
                                        "reassignment arithmetic and is stored"}
        dense.close()
        pcm_bb = synth_broadband_device(S, seed=rank, device=dev)
        ms = allmax(time_calls(lambda: eng.process_points(pcm_bb, out=out), max(3, args.steps // 4), 2, barrier))
        worst["value_broadband"] = {"value": frames_all / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                                    "kept_fraction": float(torch.count_nonzero(out[2])) / out[2].numel(),
                                    "what": "ems_process_points at the default -65 dB gate on a broadband signal: pink noise "
                                            "(-26 dBFS rms) + three 12-harmonic notes with vibrato (peaks near -20 dBFS)"}

    # ---- whole pipeline device-resident (extra, not the headline): ems_process_grid -> u8
    pipe = None
    if not args.no_pipeline:
        del out
        torch.cuda.empty_cache()
        idx = torch.empty((1, F, B), dtype=torch.uint8, device=dev)
        ms = allmax(time_calls(lambda: eng.process_grid(pcm, out=(None, idx)), args.steps, 2, barrier))
        pipe = {"value": frames_all / (ms * 1e-3), "unit": UNIT,
                "what": "ems_process_grid: PCM in HBM -> u8 colour-index image in HBM (a1-a5, fused deposit)"}
        if not args.no_dense:
            dense = emspec.Engine(n_fft=N_FFT, hop=HOP, noise_gate_db=-200.0)
            dense.use_torch_stream()
            ms = allmax(time_calls(lambda: dense.process_grid(pcm, out=(None, idx)), max(3, args.steps // 4), 2, barrier))
            worst["pipeline_u8_dense"] = {"value": frames_all / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                                          "what": "ems_process_grid with the gate at -200 dB: 2,049 deposits per frame, every "
                                                  "64-row block of the image dirty"}
            dense.close()
            dense = emspec.Engine(n_fft=N_FFT, hop=HOP, noise_gate_db=-200.0, flags=emspec.FLAG_REASSIGN)
            dense.use_torch_stream()
            ms = allmax(time_calls(lambda: dense.process_grid(pcm, out=(None, idx)), max(3, args.steps // 4), 2, barrier))
            worst["pipeline_u8_dense_fast"] = {"value": frames_all / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                                               "what": "the same with the fp32 fast mode (red.global.add.f32 instead of the "
                                                       "64-bit fixed-point reductions; not bit-exact across runs)"}
            dense.close()
            ms = allmax(time_calls(lambda: eng.process_grid(pcm_bb, out=(None, idx)), max(3, args.steps // 4), 2, barrier))
            worst["pipeline_u8_broadband"] = {"value": frames_all / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                                              "what": "ems_process_grid at the default gate on the broadband signal"}
        del idx
    pcm_bb = None
    out = None
    torch.cuda.empty_cache()

    # ---- a4 from stored points in its three modes (north_star names sort-by-bin + segmented reduce as the
    # deterministic scatter; the engine's default is 64-bit fixed-point reductions): ems_scatter_points on the
    # points of the first 300 s of the stream, sparse (default gate) and with every bin kept
    scat = None
    if rank == 0 and not args.no_dense:
        scat = {"what": "ems_scatter_points (points in HBM -> u8 image in HBM) on the first 300 s of the stream: "
                        "frames/s of the whole call; scatter_ms = the library's own events around the deposit stage",
                "sample_s": 300}
        S3 = min(S, 300 * SR)
        F3 = frame_count(S3, N_FFT, HOP)
        idx3 = torch.empty((1, F3, B), dtype=torch.uint8, device=dev)
        for sig, gate in (("sparse", args.gate_db), ("dense", -200.0)):
            src = emspec.Engine(n_fft=N_FFT, hop=HOP, noise_gate_db=gate)
            src.use_torch_stream()
            pts = src.process_points(pcm[:S3])
            src.close()
            for mode, fl in (("u64_reds", emspec.FLAG_DETERMINISTIC), ("f32_reds", 0),
                             ("sorted", emspec.FLAG_DETERMINISTIC | emspec.FLAG_SORTED_SCATTER)):
                e3 = emspec.Engine(n_fft=N_FFT, hop=HOP, noise_gate_db=gate, flags=emspec.FLAG_REASSIGN | fl)
                e3.use_torch_stream()
                fn = lambda: e3.lib.ems_scatter_points(e3.h, pts[0].data_ptr(), pts[1].data_ptr(), pts[2].data_ptr(), F3,
                                                       None, idx3.data_ptr())
                ms = time_calls(fn, 3, 1)
                scat[f"{sig}_{mode}"] = {"value": F3 / (ms * 1e-3), "unit": UNIT, "ms_per_call": ms,
                                         "scatter_ms": e3.stage_ms(emspec.STAGE_SCATTER),
                                         "scratch_bytes": int(e3.scratch_bytes())}
                e3.close()
            del pts
        del idx3
        torch.cuda.empty_cache()

    # ---- end to end: pinned host PCM -> pinned host u8 image through ems_process_host
    e2e = None
    if not args.no_e2e:
        ceiling = pcie_ceiling(dev, world, barrier, F * B, S * 4, dist)
        pcm_host = torch.empty((1, S), dtype=torch.float32, pin_memory=True)
        pcm_host.copy_(pcm[None, :])
        idx_host = torch.empty((1, F, B), dtype=torch.uint8, pin_memory=True)
        idx_host.zero_()
        heng = emspec.Engine(n_fft=N_FFT, hop=HOP, noise_gate_db=args.gate_db)   # on its own stream, as an application would
        for _ in range(max(1, min(args.warmup, 2))):
            heng.process_host(pcm_host, index_out=idx_host)
        barrier()
        t0 = time.perf_counter()
        step_s = []
        for _ in range(args.steps):
            ts = time.perf_counter()
            heng.process_host(pcm_host, index_out=idx_host)      # returns with the image in host memory
            step_s.append(time.perf_counter() - ts)
        torch.cuda.synchronize()
        wall_s = time.perf_counter() - t0
        scratch = heng.scratch_bytes()
        heng.close()
        wall = allmax(wall_s)
        e2e = {"value": frames_all * args.steps / wall, "unit": UNIT,
               "h2d_bytes_per_step": S * 4 * world, "d2h_bytes_per_step": F * B * world,
               "step_ms": {"min": 1e3 * min(step_s), "median": 1e3 * sorted(step_s)[len(step_s) // 2],
                           "max": 1e3 * max(step_s), "first": 1e3 * step_s[0]},
               "achieved_d2h_gbs": F * B * world * args.steps / wall / 1e9,
               "pcie_ceiling_gbs": ceiling["d2h_GBs"], "pcie_ceiling": ceiling,
               "frac_of_pcie_ceiling": (F * B * world * args.steps / wall / 1e9) / ceiling["d2h_GBs"],
               "engine_scratch_bytes": int(scratch), "numa": bind,
               "what": "ems_process_host: pinned host fp32 PCM -> pinned host u8 colour-index "
                       "image [F][B] (a1-a5), chunked copies overlapped with compute, O(chunk) device memory; "
                       "bound by the D2H of the image (2,049 bytes per frame)"}
        del pcm_host, idx_host

    # ---- streaming latency (configs[1]): wall time of ems_stream_push, host hop in -> final
    # column in pinned host memory, one frame per launch
    stream = None
    if rank == 0 and not args.no_stream:
        import numpy as np
        s_nfft, s_hop, s_ch = 8192, 256, 2
        seng = emspec.Engine(n_fft=s_nfft, hop=s_hop, channels=s_ch)
        colbuf = torch.empty((s_ch, s_nfft // 2 + 1), dtype=torch.uint8, pin_memory=True)
        hopbuf = (0.1 * torch.randn(64, s_hop * s_ch)).contiguous()
        lat = []
        for i in range(200 + args.stream_pushes):
            hb = hopbuf[i % 64]
            t0 = time.perf_counter()
            seng.stream_push(hb, colbuf)
            if i >= 200:
                lat.append(time.perf_counter() - t0)
        lat = np.array(lat) * 1e6
        R = -(-(s_nfft // 2) // s_hop)
        stream = {"workload": "configs[1] streaming 48 kHz stereo n_fft=8192 hop=256, one frame per push",
                  "pushes": int(lat.size), "p50_us": float(np.percentile(lat, 50)),
                  "p99_us": float(np.percentile(lat, 99)), "mean_us": float(lat.mean()),
                  "algorithmic_delay_ms": 1e3 * R * s_hop / SR,
                  "what": "wall time of ems_stream_push incl. H2D of the hop, CUDA-graph launch, D2H of the column"}
        seng.close()

    # ---- extra: configs[4] n_fft sweep (hop = n_fft/4, display controls on), device-resident
    # ems_process_points per n_fft on a 600 s stream, CUDA events on the launching stream
    sweep = None
    if rank == 0 and world == 1 and not args.no_sweep:
        sweep = []
        S4 = min(S, 600 * SR)
        for n4 in (256, 512, 1024, 2048, 4096, 8192, 16384, 32768):
            h4 = n4 // 4
            e4 = emspec.Engine(n_fft=n4, hop=h4, low_end_boost=3.9, smoothing=0.5, noise_gate_db=-65.0)
            e4.use_torch_stream()
            F4 = frame_count(S4, n4, h4)
            o4 = tuple(torch.empty((1, F4, n4 // 2 + 1), dtype=torch.float32, device=dev) for _ in range(3))
            ms4 = time_calls(lambda: e4.process_points(pcm[:S4], out=o4), 5, 3)
            gbs = b_points(n4, h4) * F4 / (ms4 * 1e-3) / 1e9
            sweep.append({"n_fft": n4, "hop": h4, "frames": F4, "frames_per_s": F4 / (ms4 * 1e-3),
                          "algorithmic_GBs": gbs, "frac_of_peak": gbs / peak})
            e4.close()
            del o4
        torch.cuda.empty_cache()

    # ---- extra (not the headline): capture-format int16 PCM in,
    # 546 display rows of the warped frequency axis out — 6 x fewer PCIe bytes per frame
    e2e_display = None
    if not args.no_e2e and not args.no_display:
        rows = 546
        deng = emspec.Engine(n_fft=N_FFT, hop=HOP, noise_gate_db=args.gate_db, display_rows=rows, freq_scale=1.0)
        q_host = torch.empty((S, 1), dtype=torch.int16, pin_memory=True)
        q_host.copy_((pcm[:, None] * 32768.0).round().clamp_(-32768, 32767).to(torch.int16))
        didx = torch.empty((1, F, rows), dtype=torch.uint8, pin_memory=True)
        for _ in range(2):
            deng.process_host_i16(q_host, index_out=didx)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            deng.process_host_i16(q_host, index_out=didx)
        torch.cuda.synchronize()
        wall = allmax(time.perf_counter() - t0)
        e2e_display = {"value": frames_all * args.steps / wall, "unit": UNIT,
                       "h2d_bytes_per_step": S * 2 * world, "d2h_bytes_per_step": F * rows * world,
                       "what": "ems_process_host_i16: pinned host int16 PCM -> pinned host u8 image "
                               "[F][546] on the warped frequency axis (display_rows=546, freq_scale=1)"}
        deng.close()
        del q_host, didx
    del pcm
    torch.cuda.empty_cache()

    # ---- configs[3]: the batch, clip-sharded (the multi-GPU workload north_star names)
    batch = None
    if not args.no_batch:
        batch = run_batch(args, emspec, rank, world, dev, dist, barrier)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        args.cpu_impl = cpu_impl_available(args.cpu_impl)
        fps, frames, wall = cpu_frames_per_s(args.cpu_seconds, cores, repeats=2, impl=args.cpu_impl)
        cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {args.cpu_seconds:g} s of the stream ({frames} frames), "
                         + CPU_WHAT[args.cpu_impl] + ", best of 2"}
        if args.cpu_impl == "c":                       # the NumPy restatement beside it, on a tenth of the sample
            nfps, nframes, _ = cpu_frames_per_s(max(2.0, args.cpu_seconds / 10), cores, repeats=1, impl="numpy")
            cpu["numpy_oracle"] = {"value": nfps, "frames": nframes, "what": CPU_WHAT["numpy"]}

    if rank == 0:
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None,
                "algorithmic_bytes_per_launch": b_points(N_FFT, HOP) * F,
                "peak_source": peak_src,
                "kernel": "stft_reassign (fused frame gather + 3-window STFT + reassignment)",
                "kernel_ms": kern_ms, "kernel_ms_last_step_lib_events": kern_last_ms,
                "bytes_per_frame": b_points(N_FFT, HOP),
                "frac_of_nominal_8TBs": achieved / 8000.0}
        if prof:
            fps = F / (kern_ms * 1e-3)
            if not prof["stale"] and args.gate_db == -65.0:
                roof["traffic"] = prof["dram_bytes_per_frame"] * F
            roof["traffic_source"] = {"file": "profiles/r02_kernel_profile.json", "kernel": prof.get("kernel"),
                                      "frames_in_capture": prof.get("frames"), "csrc_sha16": prof.get("csrc_sha16"),
                                      "stale": prof["stale"],
                                      "unit": "bytes per launch = ncu dram bytes per frame of the capture x frames of this launch"}
            roof["fp32_executed"] = {"flops_per_frame": prof.get("flops_per_frame"),
                                     "frac_of_74.45TF": fps * prof.get("flops_per_frame", 0) / 74.45e12,
                                     "how": "2 x FFMA + FADD + FMUL thread-level counts (packed x2) of the SASS mix of the ncu capture / frames",
                                     "survey_count_483378_frac": fps * 483378 / 74.45e12}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[2] offline single stream {args.seconds:g} s mono 48 kHz "
                                   "n_fft=4096 hop=128 fused STFT+reassignment -> points, one stream per GPU",
                       "n_fft": N_FFT, "hop": HOP, "frames_per_gpu": F, "samples_per_gpu": S,
                       "l2_policy": "inputs (0.69 GB) and outputs (33 GB) larger than L2, no flush needed",
                       "parallelism": f"one stream per GPU x{world} (replicas, no data-path collective); the clip-sharded "
                                      "multi-GPU workload is the `batch` key"},
            "roofline": roof,
            "cpu_baseline": cpu, "e2e": e2e, "pipeline_u8": pipe, "worst_case": worst or None, "scatter_modes": scat, "batch": batch,
            "e2e_display_rows": e2e_display, "stream_latency": stream, "nfft_sweep": sweep, "gpu_launches": int(launches),
            "clocks": clocks,
        }
        emit_json(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seconds", type=float, default=3600.0, help="stream length per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=600.0, help="audio seconds of the CPU sample")
    ap.add_argument("--cpu-impl", choices=("c", "numpy"), default="c",
                    help="which restatement of the oracle the CPU arm times (default: the C one, all host threads)")
    ap.add_argument("--gate-db", type=float, default=-65.0,
                    help="noise gate; -200 keeps every bin (worst case for the epilogue)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true")
    ap.add_argument("--no-stream", action="store_true")
    ap.add_argument("--no-display", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--stream-pushes", type=int, default=5000)
    ap.add_argument("--no-dense", action="store_true", help="skip the worst-case (dense / broadband) measurements")
    ap.add_argument("--no-batch", action="store_true", help="skip configs[3]")
    ap.add_argument("--no-batch-gather", action="store_true", help="skip the NCCL image gather of the batch (N > 1)")
    ap.add_argument("--no-bind", action="store_true", help="do not pin the rank to its GPU's NUMA node")
    ap.add_argument("--batch-clips", type=int, default=4096)
    ap.add_argument("--batch-group", type=int, default=64, help="clips per engine call (planar channels)")
    ap.add_argument("--batch-steps", type=int, default=3)
    ap.add_argument("--batch-warmup", type=int, default=1)
    args = ap.parse_args()
    claim_stdout()
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
