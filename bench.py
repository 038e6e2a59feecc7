#!/usr/bin/env python
"""bench.py — benchmark of the STFT / log-mel / iSTFT hot path (BASELINE.json metric: audio-hours/s, % of roofline).

    python bench.py --gpus N --steps K --warmup W            # our arm, headline = BASELINE configs[1]
    python bench.py --impl reference --gpus N ...            # reference arm (CPU, oracle port), same workload
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU
    python bench.py --workload parakeet_1h                   # any other named shape as the headline line

Headline (`--workload whisper128_30s`, the default): one step = one pass of the fused Whisper large-v3 front-end
(16 kHz, n_fft=400, hop=160, 128 mels, log10 / per-clip max-8 clamp / (x+4)/4) over BASELINE's batch of 4096 x 30 s
synthetic clips resident in HBM, CLIP-SHARDED across the N ranks (`--scaling strong`, BASELINE's wording: 4096 / N
clips per rank, no data-path collective).  `--scaling weak` keeps 4096 clips per rank; the default line carries the
weak figure next to the strong one under "weak".  The other named shapes of BASELINE.json ride in the same JSON line
under "workloads" (C3 Parakeet 1-hour file, FRAME-RANGE sharded with hop halos and the NCCL all-reduce of the
per-feature sums inside the timed step; C3 64-file batch; C4 Kokoro iSTFT; C5 Vocos mel / iSTFT head), each with its
roofline (HBM or FP32, whichever binds by the algorithmic counts of SURVEY 8d), a bounded CPU baseline and an
end-to-end figure from host buffers.  "variants" times the headline kernel on clips with 30 % digital silence (the
clamp fix-up's data-dependent cost) and as whisper.py calls it (padding = N_SAMPLES); "e2e_i16_f16" is the end-to-end
figure with int16 PCM in / float16 features out; "e2e_api" times the reference-shaped Python call.
Printed: ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR, N_FFT, HOP, CLIP_S = 16000, 400, 160, 30
CLIP_LEN = SR * CLIP_S
HOUR_LEN = 57_600_000  # C3: one hour at 16 kHz
VOC_LEN, VOC_SR, VOC_T = 120_000, 24000, 468  # C5: 5 s at 24 kHz, 468 frames kept
KOK_T = 24001  # C4: frames of 5 s at 24 kHz with hop 5

# name: what one UNIT is, seconds of audio per unit, algorithmic bytes and flops per unit, units in BASELINE's batch
WORKLOADS = {
    "whisper128_30s": dict(unit="30 s clip", sec=30.0, bytes=CLIP_LEN * 4 + 3000 * 128 * 4, flops=10947.0 * 3001, units=4096, cpu="whisper128"),
    "whisper80_30s": dict(unit="30 s clip", sec=30.0, bytes=CLIP_LEN * 4 + 3000 * 80 * 4, flops=10749.0 * 3001, units=4096, cpu="whisper80"),
    "parakeet_1h": dict(unit="1-hour file (frame-range sharded)", sec=3600.0, bytes=HOUR_LEN * 4 + 360001 * 80 * 4, flops=14300.0 * 360001, units=1, cpu="parakeet"),
    "parakeet_64x1h": dict(unit="1-hour file", sec=3600.0, bytes=HOUR_LEN * 4 + 360001 * 80 * 4, flops=14300.0 * 360001, units=64, cpu="parakeet"),
    "kokoro_istft": dict(unit="5 s item, (11, 24001) magnitude + phase", sec=5.0, bytes=108 * KOK_T, flops=261.0 * KOK_T, units=1024, cpu="kokoro"),
    "vocos_mel": dict(unit="5 s clip", sec=5.0, bytes=1424 * VOC_T, flops=30800.0 * VOC_T, units=8192, cpu="vocos_mel"),
    "vocos_istft": dict(unit="5 s item, (513, 468) complex64", sec=119552 / 24000.0, bytes=5128 * VOC_T, flops=27900.0 * VOC_T, units=1024, cpu="vocos_istft"),
}
# kernels of ours per step when the library's counter is unavailable (bench.py reads b2a_launch_count() around every timed loop;
# profiles/r02_launches_summary.csv lists the kernels by name).  The clamping Whisper forward is ONE cooperative launch.
LAUNCHES = {"whisper128_30s": 1, "whisper80_30s": 1, "parakeet_1h": 3, "parakeet_64x1h": 3, "kokoro_istft": 1, "vocos_mel": 1, "vocos_istft": 1}


def synth_clip_np(i: int, n: int = CLIP_LEN, sr: int = SR) -> np.ndarray:
    """§8(d) synthetic clip i: 0.1*N(0,1) + 0.2*(sin 440 + sin 3k), scaled by 0.5 + (i mod 7)/7."""
    rng = np.random.default_rng(1234 + 1 + i)
    t = np.arange(n, dtype=np.float64) / sr
    x = 0.1 * rng.standard_normal(n) + 0.2 * (np.sin(2 * np.pi * 440 * t) + np.sin(2 * np.pi * 3000 * t))
    return (x * (0.5 + (i % 7) / 7)).astype(np.float32)


# ---- CPU arm: the oracle (NumPy port of the reference; MLX is not installable here) -------------------
def _cpu_case(kind: str, i: int):
    """(callable, seconds of audio per call) of one CPU-baseline unit; the reference's own per-item call of each shape."""
    from oracle import dsp_oracle as D
    from oracle import wrappers_oracle as W

    rng = np.random.default_rng(77 + i)
    if kind in ("whisper128", "whisper80"):
        x, m = synth_clip_np(i), int(kind[7:])
        return (lambda: W.whisper_log_mel(x, m)), 30.0
    if kind == "parakeet":  # 10-minute slice of the hour per call: same per-frame work, bounded sample
        x = synth_clip_np(i, 600 * SR)
        pa = W.PreprocessArgs(16000, "per_feature", 0.025, 0.01, "hann", 80, 512, 1e-5)
        return (lambda: W.parakeet_log_mel(x, pa)), 600.0
    if kind == "kokoro":
        mag = np.minimum(np.exp(0.5 * rng.standard_normal((1, 11, KOK_T))), 1e2).astype(np.float32)
        ph = np.sin(rng.standard_normal((1, 11, KOK_T))).astype(np.float32)
        return (lambda: W.kokoro_inverse(mag, ph)), 5.0
    if kind == "vocos_mel":
        x = synth_clip_np(i, VOC_LEN, VOC_SR)
        return (lambda: W.vocos_log_mel(x)), 5.0
    if kind == "vocos_istft":
        xl = np.concatenate([0.5 * rng.standard_normal((1, VOC_T, 513)), rng.standard_normal((1, VOC_T, 513))], axis=2).astype(np.float32)
        return (lambda: W.vocos_istft_head(xl, 1024, 256)), 119552 / 24000.0
    raise KeyError(kind)


def _cpu_init(barrier):
    # one process per core: pin each worker's BLAS pool to ONE thread.  The environment variable alone does nothing
    # here — the pool was sized when the parent imported numpy, before the fork.
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:  # noqa: BLE001
        pass
    fn, _ = _cpu_case("whisper80", 0)
    fn()
    barrier.wait()


_CASES = {}


def _cpu_worker(args):
    kind, i0, count = args
    if kind not in _CASES:  # one case per worker process and workload, built (and run once: tables, FFT plans) outside the timed loop
        _CASES.clear()
        _CASES[kind] = _cpu_case(kind, i0)
        _CASES[kind][0]()
    fn, _ = _CASES[kind]
    t0 = time.perf_counter()
    for _ in range(count):
        fn()
    return time.perf_counter() - t0


_CPU_POOL = {}


def _cpu_pool(procs: int):
    """One pool per run; every worker imports the oracle and transforms one clip before the parent goes on (barrier).
    The workers' imports (scipy, the oracle) and their window / filterbank construction take seconds and are NOT the
    reference's per-clip cost — timed with them, a bounded sample understates the reference several-fold (measured
    here: 128 clips on 8 cores, 6.2 s with the imports vs 0.7 s of transform work)."""
    import multiprocessing as mp

    if procs not in _CPU_POOL:
        ctx = mp.get_context("fork")
        barrier = ctx.Barrier(procs + 1)
        _CPU_POOL[procs] = ctx.Pool(procs, initializer=_cpu_init, initargs=(barrier,))
        barrier.wait(timeout=300)
        import atexit
        atexit.register(_CPU_POOL[procs].terminate)
    return _CPU_POOL[procs]


def cpu_units_per_second(kind: str, per_proc: int, procs: int):
    """Times `per_proc` calls of the workload's reference function on each of `procs` warmed host processes (the
    reference batches with a Python loop over items, dsp.py:131 is 1-D only).  Wall clock around the whole map.
    Returns (audio seconds per wall second, calls, wall seconds)."""
    pool = _cpu_pool(procs)
    pool.map(_cpu_worker, [(kind, i, 0) for i in range(procs)], chunksize=1)  # builds + warms each worker's case
    t0 = time.perf_counter()
    pool.map(_cpu_worker, [(kind, i, per_proc) for i in range(procs)], chunksize=1)
    dt = time.perf_counter() - t0
    _, sec = _CPU_SEC[kind]
    return per_proc * procs * sec / dt, per_proc * procs, dt


# (calls per process for ~3-6 s of CPU work at ~40 ms per 30 s clip, seconds of audio per call)
_CPU_SEC = {"whisper128": (128, 30.0), "whisper80": (128, 30.0), "parakeet": (5, 600.0), "kokoro": (256, 5.0),
            "vocos_mel": (512, 5.0), "vocos_istft": (384, 119552 / 24000.0)}


def cpu_baseline(kind: str, scale: float = 1.0):
    cores = host_cores()
    per = max(1, int(_CPU_SEC[kind][0] * scale))
    sps, n, dt = cpu_units_per_second(kind, per, cores)
    return {"value": sps / 3600.0, "unit": "audio-hours/s", "cores": cores, "kind": "port",
            "sample": f"{n} calls x {_CPU_SEC[kind][1]:.1f} s of audio in {dt:.1f} s, one process per core, Python loop over items "
                      "(oracle = NumPy restatement of mlx_audio.dsp + the model wrapper; MLX itself is not installable here)"}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---- clocks sampler -----------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def reference_arm(a):
    """--impl reference: the reference's CPU implementation of the path on the box's host cores.  MLX is not
    installable in this image, so this is the oracle port (kind="port") on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[a.workload]
    kind = w["cpu"]
    cores = host_cores()
    per = max(1, int(_CPU_SEC[kind][0] * a.cpu_scale))
    for _ in range(a.warmup):
        cpu_units_per_second(kind, max(1, per // 8), cores)
    t0 = time.perf_counter()
    secs = 0.0
    calls = 0
    for _ in range(a.steps):
        sps, n, dt = cpu_units_per_second(kind, per, cores)
        secs += sps * dt
        calls += n
    dt = time.perf_counter() - t0
    ah = secs / 3600.0 / dt
    line = {
        "impl": "reference", "metric": "log-mel audio-hours/sec", "value": ah, "unit": "audio-hours/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3,
        "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": a.workload, "calls_per_step": calls // a.steps, "audio_seconds_per_call": _CPU_SEC[kind][1],
                   "note": "reference restated in NumPy (oracle port); MLX unavailable in this image"},
        "cpu_baseline": {"value": ah, "unit": "audio-hours/s", "cores": cores, "kind": "port",
                         "sample": f"{calls // a.steps} calls x {_CPU_SEC[kind][1]:.1f} s of audio per step, one process per core, Python loop over items"},
        "e2e": {"value": ah, "unit": "audio-hours/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_near_gpu(torch, dev_index):
    """Host memory of the end-to-end path (pinned input / output) should live on the NUMA node the GPU hangs off.  Prefers
    that node for this process's allocations (set_mempolicy, MPOL_PREFERRED) and, when the cgroup allows, runs on its cores.
    Best effort: returns the node or None.  (This pool's boxes are VMs with ONE NUMA node and no GPU affinity — a no-op there.)"""
    try:
        pr = torch.cuda.get_device_properties(dev_index)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        try:
            libc = C.CDLL(None, use_errno=True)
            mask = (C.c_ulong * 16)()
            mask[node // 64] = 1 << (node % 64)
            libc.syscall(C.c_long(238), C.c_int(1), mask, C.c_ulong(16 * 64))  # SYS_set_mempolicy, MPOL_PREFERRED (x86-64)
        except Exception:
            pass
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


# ---- GPU arm --------------------------------------------------------------------------------------------
class Ctx:
    """Per-process plumbing: rank / device / stream, barrier, device-timed loops with the max over ranks."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.a = torch, dist, a
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.numa_node = bind_near_gpu(torch, self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.current_stream(self.dev)
        self.sp = C.c_void_p(self.stream.cuda_stream)
        self.peak, self.peak_src = measured_peaks()
        self._fp32 = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, warmup=3, collective=True):
        """ms per step: `warmup` untimed steps, then exactly `steps` steps between two CUDA events on the launch stream,
        bracketed by barrier + synchronize on both sides; max over ranks.  The library's own launch counter is read on both
        sides of the timed loop: `self.last_launches` = kernels of ours launched inside the timed region (this rank)."""
        from mlx_audio_plus_b200 import _lib as L

        for _ in range(warmup):
            fn()
        self.barrier() if collective else self.torch.cuda.synchronize()
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        n0 = int(L.lib.b2a_launch_count())
        e0.record(self.stream)
        for _ in range(steps):
            fn()
        e1.record(self.stream)
        self.last_launches = int(L.lib.b2a_launch_count()) - n0
        self.barrier() if collective else self.torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return self.max_over_ranks(ms) if collective else ms

    def wall(self, fn, steps, warmup=1):
        """End-to-end (host buffers): wall clock around synchronous calls, max over ranks; seconds per step."""
        for _ in range(warmup):
            fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.torch.cuda.synchronize()
        return self.max_over_ranks((time.perf_counter() - t0) / steps)

    def fp32_peak(self):
        if self._fp32 is None:
            from mlx_audio_plus_b200 import _lib as L
            tf = C.c_double(0.0)
            L.check(L.lib.b2a_measure_fp32_tflops(C.byref(tf), self.sp))
            self._fp32 = tf.value
        return self._fp32

    def synth(self, B, n, sr, seed, silence_frac=0.0):
        """§8(d) synthetic batch generated on the device (torch RNG): noise + two tones, clip i scaled by 0.5 + (i mod 7)/7;
        `silence_frac` zeroes that share of every clip (digital silence in the middle third and at the end)."""
        torch = self.torch
        g = torch.Generator(device=self.dev)
        g.manual_seed(seed + self.rank)
        x = torch.empty((B, n), dtype=torch.float32, device=self.dev)
        t = torch.arange(n, device=self.dev, dtype=torch.float64) / sr
        tone = (0.2 * (torch.sin(2 * np.pi * 440 * t) + torch.sin(2 * np.pi * 3000 * t))).float()
        per = max(1, (1 << 27) // n)
        for c0 in range(0, B, per):
            c1 = min(B, c0 + per)
            scale = (0.5 + (torch.arange(c0, c1, device=self.dev) % 7).float() / 7)[:, None]
            x[c0:c1] = (0.1 * torch.randn((c1 - c0, n), generator=g, device=self.dev) + tone[None]) * scale
        if silence_frac > 0:
            k = int(n * silence_frac / 2)
            x[:, n // 3 : n // 3 + k] = 0
            x[:, n - k :] = 0
        return x

    def roofline(self, w, units, kms, traffic=None):
        """Both sides of the roofline for `units` units in `kms` ms of the dominant kernel: algorithmic bytes / measured HBM
        peak and algorithmic flops / FFMA microbenchmark; `bound` = whichever needs more time at its peak."""
        by, fl = float(w["bytes"]) * units, float(w["flops"]) * units
        t = kms * 1e-3
        fpk = self.fp32_peak()
        t_hbm, t_fp = by / (self.peak * 1e9), fl / (fpk * 1e12)
        hb = by / t / 1e9
        r = {"bound": "hbm" if t_hbm >= t_fp else "fp32", "achieved": hb, "peak": self.peak, "unit": "GB/s", "frac": hb / self.peak,
             "traffic": traffic, "peak_source": self.peak_src, "kernel_ms": kms, "algorithmic_bytes_per_launch": by,
             "fp32": {"achieved_tflops": fl / t / 1e12, "peak_tflops_measured": fpk, "frac": fl / t / 1e12 / fpk,
                      "algorithmic_flops_per_launch": fl},
             "binding_frac": max(t_hbm, t_fp) / t}
        return r


def whisper_plan(n_mels, out_dtype="float32"):
    from mlx_audio_plus_b200 import _lib as L
    from mlx_audio_plus_b200.dsp import hanning, mel_filters
    from mlx_audio_plus_b200.frontend import FrontendPlan

    return FrontendPlan(
        n_fft=N_FFT, hop=HOP, window=np.asarray(hanning(N_FFT)), center=True, pad_mode="reflect", drop_last=True,
        spec_kind=L.SPEC_POWER, filterbank=np.asarray(mel_filters(SR, N_FFT, n_mels, norm="slaney", mel_scale=None)),
        log_kind=L.LOG_LOG10, guard_kind=L.GUARD_MAX, guard_eps=1e-10, clamp_kind=L.CLAMP_CLIP_MAX, clamp_value=8.0,
        affine_add=4.0, affine_div=4.0, out_dtype=out_dtype)


def parakeet_plan():
    from mlx_audio_plus_b200 import _lib as L
    from mlx_audio_plus_b200.dsp import hanning, mel_filters
    from mlx_audio_plus_b200.frontend import FrontendPlan

    return FrontendPlan(n_fft=512, hop=160, window=np.asarray(hanning(400)), preemph=0.97, spec_kind=L.SPEC_POWER,
                        filterbank=np.asarray(mel_filters(16000, 512, 80, norm="per_feature", mel_scale=None)),
                        log_kind=L.LOG_LN, guard_kind=L.GUARD_ADD, guard_eps=1e-5, norm_kind=L.NORM_PER_FEATURE,
                        norm_ddof=0, norm_eps=1e-5)


def traffic_for(kernel_name, units):
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        per = json.load(open(tp)).get(kernel_name, {}).get("dram_bytes_per_clip")
        return per * units if per else None  # ncu --set full capture, scaled per launch
    except Exception:
        return None


def run_whisper(ctx, name, B, steps, warmup, *, e2e_steps=0, silence_frac=0.0, padding=0, kernel_alone=True):
    """Whisper front-end over B device-resident clips on this rank.  Returns the measurements of one workload."""
    from mlx_audio_plus_b200 import _lib as L

    torch = ctx.torch
    w = WORKLOADS[name]
    n_mels = int(w["cpu"][7:])
    x = ctx.synth(B, CLIP_LEN, SR, 1234 + 1, silence_frac)
    plan = whisper_plan(n_mels)
    length = CLIP_LEN + padding
    T = plan.out_frames(length)
    out = torch.empty((B, T, n_mels), dtype=torch.float32, device=ctx.dev)
    args = plan._args(x.data_ptr(), CLIP_LEN, length, CLIP_LEN, B, out.data_ptr())

    def step():
        L.check(L.lib.b2a_frontend_forward(plan._h, C.byref(args), ctx.sp))

    def step_partial():
        L.check(L.lib.b2a_frontend_partial(plan._h, C.byref(args), ctx.sp))

    ms = ctx.timed(step, steps, warmup)
    res = {"ms_per_step": ms, "units_per_rank": B, "kernel": plan.kernel_name, "frames_per_clip": T, "gpu_launches": ctx.last_launches}
    if kernel_alone:  # dominant kernel alone (init_stats + fused kernel, no finalize), same stream, CUDA events
        res["kernel_ms"] = ctx.timed(step_partial, steps, 1, collective=False)
        step()  # leave finalized features in `out` (partial() alone skips the clamp)
    if e2e_steps:
        hx = torch.empty((B, CLIP_LEN), dtype=torch.float32, pin_memory=True)
        hx.copy_(x)
        hy = torch.empty((B, T, n_mels), dtype=torch.float32, pin_memory=True)
        hargs = plan._args(hx.data_ptr(), CLIP_LEN, length, CLIP_LEN, B, hy.data_ptr())
        sec = ctx.wall(lambda: L.check(L.lib.b2a_frontend_forward_host(plan._h, C.byref(hargs))), e2e_steps)
        res["e2e"] = {"value": ctx.world * B * CLIP_S / 3600.0 / sec, "unit": "audio-hours/s",
                      "h2d_bytes_per_step": int(B * CLIP_LEN * 4), "d2h_bytes_per_step": int(B * T * n_mels * 4),
                      "ms_per_step": sec * 1e3, "steps": e2e_steps,
                      "api": "b2a_frontend_forward_host (pinned host float32 in / float32 out, chunked H2D/compute/D2H on 2 streams)",
                      "host_numa_node": ctx.numa_node}
        chk = float((hy[:4] - out[:4].cpu()).abs().max())
        assert chk == 0.0, f"host path and device path disagree: {chk}"
        del hy
        # second end-to-end line: int16 PCM in (what a decoder delivers, audio_io.py:258-262), float16 features out (what
        # whisper.py:990-996 feeds the encoder) — half the bytes each way through the same pipelined host entry
        plan16 = whisper_plan(n_mels, "float16")
        pcm = torch.empty((B, CLIP_LEN), dtype=torch.int16, pin_memory=True)
        pcm.copy_((x * (32767.0 / 1.2)).clamp(-32768, 32767).to(torch.int16))
        hy16 = torch.empty((B, T, n_mels), dtype=torch.float16, pin_memory=True)
        a16 = plan16._args(pcm.data_ptr(), CLIP_LEN, length, CLIP_LEN, B, hy16.data_ptr())
        a16.audio_kind = L.PCM_I16
        sec = ctx.wall(lambda: L.check(L.lib.b2a_frontend_forward_host(plan16._h, C.byref(a16))), e2e_steps)
        res["e2e_i16_f16"] = {"value": ctx.world * B * CLIP_S / 3600.0 / sec, "unit": "audio-hours/s",
                              "h2d_bytes_per_step": int(B * CLIP_LEN * 2), "d2h_bytes_per_step": int(B * T * n_mels * 2),
                              "ms_per_step": sec * 1e3, "steps": e2e_steps,
                              "api": "b2a_frontend_forward_host, audio_kind = B2A_PCM_I16, out_dtype = float16 (pinned host buffers)"}
        # what bounds it: this rank's pinned copy rates while every rank copies (measured, not read off the topology)
        res["e2e_i16_f16"]["pinned_copy_gbs_all_ranks_active"] = copy_rates(ctx, hx)
        assert bool(torch.isfinite(hy16[:2].float()).all())
        del hx, pcm, hy16, plan16
    del x, out
    torch.cuda.empty_cache()
    return res


def copy_rates(ctx, hbuf):
    """Pinned host <-> device copy rate of this rank with all ranks copying at once (the host-side limiter of the end-to-end
    figure at N > 1): 1 GiB each way, then both ways at once on two streams.  GB/s, min over ranks."""
    torch = ctx.torch
    n = min(hbuf.numel(), 1 << 28)
    h = hbuf.view(-1)[:n]
    d = torch.empty(n, dtype=hbuf.dtype, device=ctx.dev)
    d2 = torch.empty(n, dtype=hbuf.dtype, device=ctx.dev)
    h2 = torch.empty(n, dtype=hbuf.dtype, pin_memory=True)
    by = n * hbuf.element_size()
    s2 = torch.cuda.Stream(ctx.dev)

    def t(fn):
        sec = ctx.wall(fn, 2, 1)
        return by / sec / 1e9

    def both():
        d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()

    r = {"h2d": t(lambda: (d.copy_(h, non_blocking=True), torch.cuda.synchronize())),
         "d2h": t(lambda: (h2.copy_(d2, non_blocking=True), torch.cuda.synchronize())),
         "two_way_each": t(both), "ranks_active": ctx.world}
    del d, d2, h2
    return r


def run_parakeet_frames(ctx, steps, warmup, e2e_steps=0):
    """C3: ONE 1-hour file, frames sharded over the ranks (halo n_fft - hop + 1 pre-emphasis sample).  A step = partial on
    the rank's frame range, the all-reduce of the (2 * 80) float64 per-feature sums (NCCL), finalize with the global
    statistics — parallel.long_form_features, the all-reduce INSIDE the timed region."""
    from mlx_audio_plus_b200.parallel import frame_shards, long_form_features, num_frames

    torch = ctx.torch
    w = WORKLOADS["parakeet_1h"]
    plan = parakeet_plan()
    sh = frame_shards(HOUR_LEN, 512, 160, ctx.world, preemph=True)[ctx.rank]
    T = num_frames(HOUR_LEN, 512, 160)
    xs = ctx.synth(1, sh.sample_hi - sh.sample_lo, SR, 1236)[0]
    group = None

    obuf = torch.empty(plan.out_shape(1, sh.frame_count), dtype=torch.float32, device=ctx.dev)

    def step():
        return long_form_features(plan, xs, sh, length=HOUR_LEN, global_frames=T, group=group, out=obuf)

    y = step()
    assert tuple(y.shape) == (sh.frame_count, 80) and bool(torch.isfinite(y).all())
    ms = ctx.timed(step, steps, warmup)
    res = {"ms_per_step": ms, "value": 1.0 / (ms * 1e-3), "unit": "audio-hours/s", "kernel": plan.kernel_name, "gpu_launches": ctx.last_launches,
           "shard": "frames", "frames_per_rank": sh.frame_count, "halo_samples": 512 - 160 + 1, "global_frames": T,
           "collective": "one all_reduce(SUM, 160 float64) per step (NCCL)" if ctx.world > 1 else "none (1 rank)",
           "api": "parallel.long_form_features (partial -> reduce_stats -> finalize)",
           "roofline": ctx.roofline(w, 1.0 / ctx.world, ms) | {"note": "whole step (2 launches + normalise sweep + all-reduce); one file is latency-bound, see parakeet_64x1h"}}
    if e2e_steps and ctx.world == 1:
        from mlx_audio_plus_b200.stt.models.parakeet.audio import PreprocessArgs, log_mel_spectrogram
        pa = PreprocessArgs(sample_rate=16000, normalize="per_feature", window_size=0.025, window_stride=0.01, window="hann",
                            features=80, n_fft=512, dither=0.0)
        hx = xs.cpu().numpy()
        sec = ctx.wall(lambda: log_mel_spectrogram(hx, pa), e2e_steps)
        res["e2e"] = {"value": 1.0 / sec, "unit": "audio-hours/s", "h2d_bytes_per_step": int(hx.nbytes), "d2h_bytes_per_step": T * 80 * 4,
                      "ms_per_step": sec * 1e3, "api": "stt.models.parakeet.audio.log_mel_spectrogram(numpy 1-D float32) -> numpy (1, T, 80)"}
    del xs, y
    torch.cuda.empty_cache()
    return res


def run_parakeet_batch(ctx, B, steps, warmup):
    """C3 roofline number: a batch of 1-hour files, clip-sharded (B on this rank), forward = fused kernel + normalise sweep."""
    from mlx_audio_plus_b200 import _lib as L

    torch = ctx.torch
    w = WORKLOADS["parakeet_64x1h"]
    plan = parakeet_plan()
    x = ctx.synth(B, HOUR_LEN, SR, 1236)
    T = plan.out_frames(HOUR_LEN)
    out = torch.empty((B, T, 80), dtype=torch.float32, device=ctx.dev)
    args = plan._args(x.data_ptr(), HOUR_LEN, HOUR_LEN, HOUR_LEN, B, out.data_ptr())
    ms = ctx.timed(lambda: L.check(L.lib.b2a_frontend_forward(plan._h, C.byref(args), ctx.sp)), steps, warmup)
    nl = ctx.last_launches
    kms = ctx.timed(lambda: L.check(L.lib.b2a_frontend_partial(plan._h, C.byref(args), ctx.sp)), steps, 1, collective=False)
    res = {"ms_per_step": ms, "value": ctx.world * B / (ms * 1e-3), "unit": "audio-hours/s", "kernel": plan.kernel_name,
           "units_per_rank": B, "shard": "clips", "gpu_launches": nl,
           "roofline": ctx.roofline(w, B, kms) | {"step_binding_frac_incl_normalise_sweep": None}}
    r = res["roofline"]
    r["step_binding_frac_incl_normalise_sweep"] = r["binding_frac"] * kms / ms
    del x, out
    torch.cuda.empty_cache()
    return res


def run_vocos_mel(ctx, B, steps, warmup, e2e_steps=0):
    from mlx_audio_plus_b200 import _lib as L
    from mlx_audio_plus_b200.dsp import hanning, mel_filters
    from mlx_audio_plus_b200.frontend import FrontendPlan

    torch = ctx.torch
    w = WORKLOADS["vocos_mel"]
    plan = FrontendPlan(n_fft=1024, hop=256, window=np.asarray(hanning(1024)), drop_last=True, spec_kind=L.SPEC_MAGNITUDE,
                        filterbank=np.asarray(mel_filters(24000, 1024, 100, norm=None, mel_scale="htk")),
                        log_kind=L.LOG_LN, guard_kind=L.GUARD_MAX, guard_eps=1e-5)
    x = ctx.synth(B, VOC_LEN, VOC_SR, 1238)
    T = plan.out_frames(VOC_LEN)
    out = torch.empty((B, T, 100), dtype=torch.float32, device=ctx.dev)
    args = plan._args(x.data_ptr(), VOC_LEN, VOC_LEN, VOC_LEN, B, out.data_ptr())
    ms = ctx.timed(lambda: L.check(L.lib.b2a_frontend_forward(plan._h, C.byref(args), ctx.sp)), steps, warmup)
    res = {"ms_per_step": ms, "value": ctx.world * B * 5.0 / 3600.0 / (ms * 1e-3), "unit": "audio-hours/s", "kernel": plan.kernel_name,
           "units_per_rank": B, "shard": "clips", "gpu_launches": ctx.last_launches, "roofline": ctx.roofline(w, B, ms)}
    if e2e_steps:
        hx = torch.empty((B, VOC_LEN), dtype=torch.float32, pin_memory=True)
        hx.copy_(x)
        hy = torch.empty((B, T, 100), dtype=torch.float32, pin_memory=True)
        ha = plan._args(hx.data_ptr(), VOC_LEN, VOC_LEN, VOC_LEN, B, hy.data_ptr())
        sec = ctx.wall(lambda: L.check(L.lib.b2a_frontend_forward_host(plan._h, C.byref(ha))), e2e_steps)
        res["e2e"] = {"value": ctx.world * B * 5.0 / 3600.0 / sec, "unit": "audio-hours/s", "h2d_bytes_per_step": int(hx.numel() * 4),
                      "d2h_bytes_per_step": int(hy.numel() * 4), "ms_per_step": sec * 1e3, "api": "b2a_frontend_forward_host (pinned)"}
        del hx, hy
    del x, out
    torch.cuda.empty_cache()
    return res


def run_istft(ctx, name, B, steps, warmup, e2e_steps=0):
    """C4 (Kokoro 20/5, magnitude + phase planes: cos / sin formed inside the kernel, istftnet.py:505-519) and C5 inverse
    (Vocos head 1024/256, complex64 (513, 468) per item, symmetric Hann array window, sum-w normalisation)."""
    from mlx_audio_plus_b200 import _lib as L
    from mlx_audio_plus_b200.dsp import hanning
    from mlx_audio_plus_b200.frontend import IstftPlan

    torch = ctx.torch
    w = WORKLOADS[name]
    g = torch.Generator(device=ctx.dev)
    g.manual_seed(7 + ctx.rank)
    a = L.InverseArgs()
    if name == "kokoro_istft":
        F, T, polar = 11, KOK_T, True
        plan = IstftPlan(n_fft=20, hop=5, window=np.asarray(hanning(21)[:-1]), center=True, polar=True)
        mag = torch.exp(0.5 * torch.randn((B, F, T), generator=g, device=ctx.dev)).clamp(max=1e2)
        ph = torch.sin(torch.randn((B, F, T), generator=g, device=ctx.dev))
        a.spec, a.spec_imag = mag.data_ptr(), ph.data_ptr()
        hold = (mag, ph)
    else:
        F, T, polar = 513, VOC_T, False
        plan = IstftPlan(n_fft=1024, hop=256, window=np.asarray(hanning(1024)), center=True)
        mag = torch.exp(0.5 * torch.randn((B, F, T), generator=g, device=ctx.dev)).clamp(max=1e2)
        ph = torch.randn((B, F, T), generator=g, device=ctx.dev)
        spec = torch.complex(mag * torch.cos(ph), mag * torch.sin(ph)).contiguous()
        del mag, ph
        a.spec, a.spec_imag = spec.data_ptr(), None
        hold = (spec,)
    n_out = plan.out_len(T)
    out = torch.empty((B, n_out), dtype=torch.float32, device=ctx.dev)
    a.clip_stride, a.num_frames, a.batch, a.length, a.out_clip_stride, a.out = 0, T, B, -1, 0, out.data_ptr()
    ms = ctx.timed(lambda: L.check(L.lib.b2a_istft_inverse(plan._h, C.byref(a), ctx.sp)), steps, warmup)
    sec_unit = w["sec"]
    res = {"ms_per_step": ms, "value": ctx.world * B * sec_unit / 3600.0 / (ms * 1e-3), "unit": "audio-hours/s", "kernel": plan.kernel_name,
           "units_per_rank": B, "shard": "clips", "input_form": "magnitude / phase planes" if polar else "complex64",
           "gpu_launches": ctx.last_launches, "roofline": ctx.roofline(w, B, ms)}
    if name == "kokoro_istft":  # the same spectra as ONE complex64 tensor (round 1's C4 figure): what the cos / sin inside the kernel cost
        planc = IstftPlan(n_fft=20, hop=5, window=np.asarray(hanning(21)[:-1]), center=True)
        spec = torch.complex(hold[0] * torch.cos(hold[1]), hold[0] * torch.sin(hold[1])).contiguous()
        ac = L.InverseArgs()
        ac.spec, ac.spec_imag = spec.data_ptr(), None
        ac.clip_stride, ac.num_frames, ac.batch, ac.length, ac.out_clip_stride, ac.out = 0, T, B, -1, 0, out.data_ptr()
        msc = ctx.timed(lambda: L.check(L.lib.b2a_istft_inverse(planc._h, C.byref(ac), ctx.sp)), steps, warmup)
        res["complex_input"] = {"ms_per_step": msc, "hbm_frac": float(w["bytes"]) * B / (msc * 1e-3) / 1e9 / ctx.peak,
                                "note": "same items handed over as complex64 (mag * exp(i phase) formed beforehand)"}
        L.check(L.lib.b2a_istft_inverse(plan._h, C.byref(a), ctx.sp))  # `out` holds the polar result again (checked below)
        del spec, planc
    if e2e_steps:
        hs = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in hold]
        for h, t in zip(hs, hold):
            h.copy_(t)
        hy = torch.empty((B, n_out), dtype=torch.float32, pin_memory=True)
        ha = L.InverseArgs()
        ha.spec, ha.spec_imag = hs[0].data_ptr(), (hs[1].data_ptr() if polar else None)
        ha.clip_stride, ha.num_frames, ha.batch, ha.length, ha.out_clip_stride, ha.out = 0, T, B, -1, 0, hy.data_ptr()
        sec = ctx.wall(lambda: L.check(L.lib.b2a_istft_inverse_host(plan._h, C.byref(ha))), e2e_steps)
        res["e2e"] = {"value": ctx.world * B * sec_unit / 3600.0 / sec, "unit": "audio-hours/s",
                      "h2d_bytes_per_step": int(sum(h.numel() * h.element_size() for h in hs)), "d2h_bytes_per_step": int(hy.numel() * 4),
                      "ms_per_step": sec * 1e3, "api": "b2a_istft_inverse_host (pinned)"}
        chk = float((hy[:2] - out[:2].cpu()).abs().max())
        assert chk == 0.0, f"iSTFT host path and device path disagree: {chk}"
        del hs, hy
    del hold, out
    torch.cuda.empty_cache()
    return res


def run_api(ctx):
    """The reference-shaped call, timed as a user makes it (whisper/audio.py:44-85): wrapper overhead included (ingest,
    plan cache, allocation, per-call workspace).  Rank 0 only."""
    from mlx_audio_plus_b200.stt.models.whisper.audio import log_mel_spectrogram

    torch = ctx.torch
    res = {}
    x1 = synth_clip_np(0)
    xd = torch.from_numpy(x1).to(ctx.dev)

    def lat(fn, n=200):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts) * 1e6)

    # C1: one 30 s clip, 80 mels
    res["c1_latency_us_torch_cuda_in"] = lat(lambda: log_mel_spectrogram(xd, n_mels=80))
    res["c1_latency_us_numpy_in_numpy_out"] = lat(lambda: log_mel_spectrogram(x1, n_mels=80), 50)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(10):
        log_mel_spectrogram(xd, n_mels=80)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(200):
        log_mel_spectrogram(xd, n_mels=80)
    e1.record()
    torch.cuda.synchronize()
    res["c1_us_per_call_back_to_back_device"] = e0.elapsed_time(e1) / 200 * 1e3  # launch-rate bound: wrapper + 3 launches
    # the same clip with 128 mels: a grid this small takes the single cooperative launch (statistics, grid barrier and clamp
    # fix-up inside the fused kernel) instead of init_stats + kernel + fix-up
    from mlx_audio_plus_b200 import _lib as L
    n0 = int(L.lib.b2a_launch_count())
    res["c1_128mel_latency_us_torch_cuda_in"] = lat(lambda: log_mel_spectrogram(xd, n_mels=128))
    res["c1_128mel_launches_per_call"] = (int(L.lib.b2a_launch_count()) - n0) / 220.0
    res["c1_audio_hours_per_s_back_to_back"] = 30.0 / 3600.0 / (res["c1_us_per_call_back_to_back_device"] * 1e-6)
    # per-clip loop (how the reference batches: dsp.py:131 is 1-D only) vs one batched call, host NumPy arrays, 128 mels
    xs = np.stack([synth_clip_np(i) for i in range(64)])
    log_mel_spectrogram(xs[0], n_mels=128)  # plan + staging buffers of the host path exist before the clock starts
    t0 = time.perf_counter()
    for i in range(64):
        log_mel_spectrogram(xs[i], n_mels=128)
    res["numpy_loop_64_clips_audio_hours_per_s"] = 64 * 30 / 3600.0 / (time.perf_counter() - t0)
    log_mel_spectrogram(xs, n_mels=128)
    t0 = time.perf_counter()
    log_mel_spectrogram(xs, n_mels=128)
    res["numpy_batched_64_clips_audio_hours_per_s"] = 64 * 30 / 3600.0 / (time.perf_counter() - t0)
    xb = ctx.synth(1024, CLIP_LEN, SR, 99)
    ms = ctx.timed(lambda: log_mel_spectrogram(xb, n_mels=128), 10, 3, collective=False)
    res["torch_cuda_batched_1024_clips_ms"] = ms
    res["torch_cuda_batched_1024_clips_audio_hours_per_s"] = 1024 * 30 / 3600.0 / (ms * 1e-3)
    res["api"] = "stt.models.whisper.audio.log_mel_spectrogram"
    del xb, xd
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="whisper128_30s", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: BASELINE's batch sharded over the ranks (default); weak: BASELINE's batch per rank")
    ap.add_argument("--shard", default=None, choices=["clips", "frames"], help="informational: parakeet_1h shards by frames, the rest by clips")
    ap.add_argument("--clips", type=int, default=None, help="units in the batch (default: BASELINE's, e.g. 4096 clips)")
    ap.add_argument("--cpu-scale", type=float, default=1.0, help="scales the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline workload only (no other workloads / variants / api timings)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3  # timing rule: W >= 3
    if a.impl == "reference":
        return reference_arm(a)
    # stdout carries exactly ONE JSON line: anything a library prints there on the way (NCCL's "NCCL version ..." banner
    # when NCCL_DEBUG is set in the environment) is sent to stderr; the real stdout comes back for the final print.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    name = a.workload
    w = WORKLOADS[name]
    extras = not a.no_extras and name == "whisper128_30s" and a.clips is None
    want_cpu = rank == 0 and world == 1 and not a.no_cpu_baseline

    # CPU baselines first: before CUDA is initialised (fork-safe), rank 0 at N = 1 only, bounded samples
    cpu = {}
    if want_cpu:
        kinds = [w["cpu"]] + ([k for k in ("parakeet", "kokoro", "vocos_mel", "vocos_istft")] if extras else [])
        for k in kinds:
            cpu[k] = cpu_baseline(k, a.cpu_scale * (2.0 if k == w["cpu"] else 1.0))
        if _CPU_POOL:
            for p in _CPU_POOL.values():
                p.terminate()
            _CPU_POOL.clear()

    ctx = Ctx(a)
    total = a.clips if a.clips is not None else w["units"]
    e2e_steps = 0 if a.no_e2e else a.e2e_steps

    def per_rank(n, scaling):  # contiguous balanced shard (parallel.clip_shard)
        if scaling == "weak":
            return n
        from mlx_audio_plus_b200.parallel import clip_shard
        lo, hi = clip_shard(n, world, rank)
        return hi - lo

    sampler = ClockSampler(ctx.local_rank)
    if rank == 0 and not os.environ.get("B2A_BENCH_NO_SAMPLER"):  # (development: A/B of the sampler's own perturbation)
        sampler.start()
        time.sleep(0.25)

    # ---- headline ----------------------------------------------------------------------------------------
    B = per_rank(total, a.scaling)
    units_global = total if a.scaling == "strong" else total * world
    if name.startswith("whisper"):
        h = run_whisper(ctx, name, B, a.steps, a.warmup, e2e_steps=e2e_steps)
        h["roofline"] = ctx.roofline(w, B, h["kernel_ms"], traffic_for(h["kernel"], B))
    elif name == "parakeet_1h":
        h = run_parakeet_frames(ctx, a.steps, a.warmup, e2e_steps)
        units_global = 1
    elif name == "parakeet_64x1h":
        h = run_parakeet_batch(ctx, B, min(a.steps, 5), a.warmup)
    elif name == "vocos_mel":
        h = run_vocos_mel(ctx, B, a.steps, a.warmup, e2e_steps)
    else:
        h = run_istft(ctx, name, B, a.steps, a.warmup, e2e_steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = h["ms_per_step"]
    value = units_global * w["sec"] / 3600.0 / (ms_per_step * 1e-3)

    line = {
        "metric": "log-mel audio-hours/sec", "value": value, "unit": "audio-hours/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if name == "parakeet_1h" else a.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "unit": w["unit"], "units_per_gpu": h.get("units_per_rank", h.get("frames_per_rank")),
                   "global_units": units_global, "global_clips": units_global, "audio_seconds_per_unit": w["sec"],
                   "parallelism": ("frame-range shard" if name == "parakeet_1h" else "clip-shard") + f" x{world}",
                   "l2_policy": "inputs + outputs per step far exceed the 126 MB L2 at N <= 8 (headline: 14.2 GB / N per rank)",
                   "kernel": h["kernel"]},
        "roofline": h["roofline"], "cpu_baseline": cpu.get(w["cpu"]), "e2e": h.get("e2e"),
        "gpu_launches": h.get("gpu_launches", LAUNCHES[name] * a.steps), "clocks": clocks,
    }
    if "e2e_i16_f16" in h:
        line["e2e_i16_f16"] = h["e2e_i16_f16"]

    # ---- the rest of BASELINE's shapes, in the same line ------------------------------------------------------
    if extras:
        ex_steps = max(3, min(a.steps, 10))
        if world > 1:  # weak-scaling companion of the strong headline
            hw = run_whisper(ctx, name, total, a.steps, a.warmup, kernel_alone=False)
            line["weak"] = {"value": world * total * CLIP_S / 3600.0 / (hw["ms_per_step"] * 1e-3), "unit": "audio-hours/s",
                            "ms_per_step": hw["ms_per_step"], "clips_per_gpu": total, "global_clips": total * world}
        else:
            line["weak"] = {"value": value, "unit": "audio-hours/s", "ms_per_step": ms_per_step, "clips_per_gpu": total, "global_clips": total}
        wl = {}
        wl["parakeet_1h"] = run_parakeet_frames(ctx, ex_steps, 3, e2e_steps)
        wl["parakeet_64x1h"] = run_parakeet_batch(ctx, per_rank(64, "strong"), 3, 3)
        wl["kokoro_istft"] = run_istft(ctx, "kokoro_istft", per_rank(1024, "strong"), ex_steps, 3, e2e_steps)
        wl["vocos_mel"] = run_vocos_mel(ctx, per_rank(8192, "strong"), ex_steps, 3, e2e_steps)
        wl["vocos_istft"] = run_istft(ctx, "vocos_istft", per_rank(1024, "strong"), ex_steps, 3, e2e_steps)
        for k, r in wl.items():
            r["config"] = {"workload": k, "unit": WORKLOADS[k]["unit"], "global_units": WORKLOADS[k]["units"], "scaling": "strong", "n_gpus": world}
            r["cpu_baseline"] = cpu.get(WORKLOADS[k]["cpu"])
        line["workloads"] = wl
        # the clamp fix-up's data-dependent cost, and the call as whisper.py makes it (padding = N_SAMPLES)
        var = {}
        hv = run_whisper(ctx, name, B, ex_steps, 3, silence_frac=0.3, kernel_alone=False)
        var["silence_30pct"] = {"ms_per_step": hv["ms_per_step"], "vs_headline": hv["ms_per_step"] / ms_per_step,
                                "note": "30 % of every clip is digital silence: those tiles lie below max - 8 and are rewritten by clamp_fixup_kernel"}
        hv = run_whisper(ctx, name, B, ex_steps, 3, padding=CLIP_LEN, kernel_alone=False)
        var["padding_n_samples"] = {"ms_per_step": hv["ms_per_step"], "frames_per_clip": hv["frames_per_clip"],
                                    "note": "log_mel_spectrogram(audio, padding=N_SAMPLES) as whisper.py calls it: (B, 6000, 128) out, all-padding rows filled"}
        line["variants"] = var
        if rank == 0:
            line["e2e_api"] = run_api(ctx)
    if rank == 0:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        ctx.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
