#!/usr/bin/env python
"""bench.py -- audio-seconds/second of the madmom-style spectral front end on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): a batch of 64 synthetic 3-minute 44.1 kHz mono stems through the
multi-resolution front end madmom's RNNBeatProcessor uses (frames 1024/2048/4096, hop 441, 3/6/12
bands per octave, log10(1+x), positive spectral-flux difference, stacked) -> (18000, 314) per stem.
One step = one pass of that path over the whole batch.  With N > 1 every rank runs the same batch on
its own GPU (job sharding, no data-path collective): weak scaling, value = all ranks' audio seconds /
max-over-ranks device time.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput; `e2e` is the same metric
through FrontEnd.process_batch_pinned (pinned host buffers in, pinned host buffers out, copies
inside the timed region; `e2e.pcie_bound_value` = what the larger one-direction copy of the step alone
allows, `e2e.pcie_concurrent_value` = one H2D and one D2H copy of the step's bytes started together);
`roofline` is for the dominant kernel (frame 4096), CUDA events around every launch of the timed region; `cpu_baseline` times the numpy oracle (madmom restatement) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

# The CPU legs run ONE single-threaded worker process per host core (how the reference scales: one Celery
# worker per job).  The BLAS / OpenMP pools must be capped BEFORE numpy is first imported -- a forked
# worker inherits the parent's already-initialised OpenBLAS pool, and a later os.environ change does
# nothing (round 1 measured 16 workers x 16 BLAS threads: ~6x too slow a baseline).  The GPU arm does
# no host arithmetic, so the cap costs it nothing.
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
    os.environ[_v] = "1"

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SR = 44100
N_CLIPS, CLIP_SECONDS = 64, 180
METRIC = "audio-sec/sec log-filtered spectrogram (madmom multi-resolution front end, frames 1024/2048/4096 + flux)"
UNIT = "audio-s/s"
WORKLOAD = "64 x 180 s 44.1 kHz mono f32 stems, frames 1024/2048/4096 hop 441, 3/6/12 bpo, log10(1+x), positive diff, stacked -> (18000,314)/stem"


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle (numpy restatement of madmom 0.16.1) on the host cores
# ------------------------------------------------------------------------------------------------
_CPU_CLIPS = {}      # (seed, seconds) -> synthetic guitar clip; filled in the parent BEFORE the timed pool forks


def _blas_threads(_=None):
    """threads of the BLAS / OpenMP pools loaded in this process (1 each is what the baseline wants)"""
    try:
        from threadpoolctl import threadpool_info
        return sorted({int(i.get("num_threads", 0)) for i in threadpool_info()}) or [1]
    except Exception:
        return None


def _cpu_synth(args):
    seed, seconds = args
    from audio_tabs_b200.synth import synth_guitar
    return synth_guitar(seed, seconds)


def _cpu_worker(args):
    from oracle import madmom_ref as ref
    x = _CPU_CLIPS[args]
    out = ref.rnn_beat_preprocessor()(x)
    return out.shape[0]


def cpu_front_end_throughput(clip_seconds=30.0, clips_per_core=4, cores=None):
    """audio-s/s of the oracle (madmom restatement: per-frame scipy.fftpack.fft loop, float32 filterbank dot,
    log10, difference, hstack -- the three resolutions of RNNBeatProcessor) on synthetic guitar clips
    (SURVEY.md section 8d generator, seed = 2000 + i as for config 2), one single-threaded worker process per
    host core (how the reference scales: one Celery worker per job, /root/reference/docker-compose.yml:28).
    Audio-s/s does not depend on the clip length on this path, so the sample is `clips_per_core` clips of
    `clip_seconds` per core instead of the 64 x 180 s of the GPU arm.  Clip synthesis is outside the timed
    region (the clips are made first and inherited by the forked workers)."""
    import multiprocessing as mp
    cores = cores or len(os.sched_getaffinity(0)) or os.cpu_count() or 1
    jobs = [(2000 + i, float(clip_seconds)) for i in range(cores * clips_per_core)]
    ctx = mp.get_context("fork")
    todo = [j for j in jobs + [(0, 1.0)] if j not in _CPU_CLIPS]
    if todo:
        with ctx.Pool(cores) as pool:                    # untimed: synthesise the clips in parallel (kept across steps)
            for j, x in zip(todo, pool.map(_cpu_synth, todo, chunksize=1)):
                _CPU_CLIPS[j] = x
    with ctx.Pool(cores) as pool:                        # forked now: the workers see _CPU_CLIPS
        pool.map(_cpu_worker, [(0, 1.0)] * cores)        # spin up workers, import scipy, build the filterbanks
        blas = pool.map(_blas_threads, range(cores))[0]
        t0 = time.perf_counter()
        pool.map(_cpu_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    audio = clip_seconds * len(jobs)
    info = {"cores": cores, "blas_threads_per_worker": blas, "workers": cores,
            "sample": "%d synthetic guitar clips x %.0f s (same generator and front end as the GPU arm), %d single-threaded "
                      "worker processes" % (len(jobs), clip_seconds, cores)}
    return audio / dt, info, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    for _ in range(min(args.warmup, 1)):
        cpu_front_end_throughput(clip_seconds=5.0, clips_per_core=1)
    vals, dts = [], []
    for _ in range(args.steps):
        v, info, dt = cpu_front_end_throughput()
        vals.append(v)
        dts.append(dt)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(dts)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(N_CLIPS, CLIP_SECONDS, args.gpus),
        "reference_note": "CPU oracle port of madmom 0.16.1 (madmom itself is not installable offline); computes f64 FFT / "
                          "f32 spectrogram like madmom; each step is a bounded sample of the workload: " + info["sample"],
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": "port", "sample": info["sample"],
                         "blas_threads_per_worker": info["blas_threads_per_worker"],
                         "per_core": value / info["cores"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bench_config(n_clips, clip_seconds, world):
    """the `config` object both arms print (same keys, same values for the same workload)"""
    full = (n_clips, clip_seconds) == (N_CLIPS, CLIP_SECONDS)
    in_gb = n_clips * clip_seconds * SR * 4 / 1e9
    out_gb = n_clips * int(np.ceil(clip_seconds * SR / 441.0)) * 314 * 4 / 1e9
    return {"workload": WORKLOAD if full else "%d x %.0f s stems (reduced)" % (n_clips, clip_seconds),
            "clips_per_gpu": n_clips, "clip_seconds": clip_seconds, "parallelism": "job-sharded x%d" % world,
            "l2": "inputs (%.2f GB) + outputs (%.2f GB) per step exceed the 126 MB L2" % (in_gb, out_gb)}


# ------------------------------------------------------------------------------------------------
# clock sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.file = None

    def start(self):
        try:
            self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, settle=0.15):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(settle)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        sm, smmax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.file.read().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smmax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.file.name)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(smmax)) if smmax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of each k_front kernel at the full bench size,
    from the committed `ncu --set full` capture (profiles/; written by tools/ncu_traffic.py)."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        return json.loads(p.read_text())
    except Exception:
        return {}


def pcie_bandwidth(dev, in_bytes, out_bytes):
    """What the host <-> device link allows for one step (explains the e2e bound): GB/s of a pinned H2D copy of
    the step's input and a D2H copy of its output, each alone and both at once (the pipeline overlaps them,
    and the two directions slow each other down by ~10 %)."""
    import torch
    hi = torch.empty(in_bytes, dtype=torch.uint8, pin_memory=True)
    ho = torch.empty(out_bytes, dtype=torch.uint8, pin_memory=True)
    di = torch.empty(in_bytes, dtype=torch.uint8, device=dev)
    do = torch.empty(out_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def timed(fn):
        fn()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b)

    def both():
        cur = torch.cuda.current_stream(dev)
        e = torch.cuda.Event()
        e.record(cur)
        s1.wait_event(e)
        s2.wait_event(e)
        with torch.cuda.stream(s1):
            di.copy_(hi, non_blocking=True)
            e1 = torch.cuda.Event()
            e1.record(s1)
        with torch.cuda.stream(s2):
            ho.copy_(do, non_blocking=True)
            e2 = torch.cuda.Event()
            e2.record(s2)
        cur.wait_event(e1)
        cur.wait_event(e2)

    ms_h2d = min(timed(lambda: di.copy_(hi, non_blocking=True)) for _ in range(2))
    ms_d2h = min(timed(lambda: ho.copy_(do, non_blocking=True)) for _ in range(2))
    ms_both = min(timed(both) for _ in range(3))
    return {"h2d_gbs": in_bytes / ms_h2d / 1e6, "d2h_gbs": out_bytes / ms_d2h / 1e6, "h2d_ms": ms_h2d, "d2h_ms": ms_d2h,
            "both_directions_ms": ms_both}


def fabric_concurrent(dev, in_bytes, out_bytes, barrier, world, reps=3):
    """The host <-> device copy ceiling of THIS box for THIS step, measured with every rank copying at once:
    after a barrier each rank starts one pinned H2D copy of its step's input and one D2H copy of its output
    on two streams; the time is the max over ranks, best of `reps`.  At N > 1 the ranks share the host's
    PCIe root / memory fabric, so this -- not rank 0's solo bandwidth times N -- is what e2e can reach."""
    import torch
    import torch.distributed as dist
    hi = torch.empty(in_bytes, dtype=torch.uint8, pin_memory=True)
    ho = torch.empty(out_bytes, dtype=torch.uint8, pin_memory=True)
    di = torch.empty(in_bytes, dtype=torch.uint8, device=dev)
    do = torch.empty(out_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    best = None
    for _ in range(reps + 1):                      # first round is a warm-up
        barrier()
        cur = torch.cuda.current_stream(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(cur)
        s1.wait_event(a)
        s2.wait_event(a)
        with torch.cuda.stream(s1):
            di.copy_(hi, non_blocking=True)
            e1 = torch.cuda.Event()
            e1.record(s1)
        with torch.cuda.stream(s2):
            ho.copy_(do, non_blocking=True)
            e2 = torch.cuda.Event()
            e2.record(s2)
        cur.wait_event(e1)
        cur.wait_event(e2)
        b.record(cur)
        torch.cuda.synchronize(dev)
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        best = ms if best is None or _ == 0 else min(best, ms)
    del hi, ho, di, do
    return best


def time_e2e(fe, host_in, lens, host_out, steps, barrier, dev, world):
    """`steps` passes of FrontEnd.process_batch_pinned (pinned host in -> device -> pinned host out), copies inside
    the timed region; CUDA events on the current stream, max over ranks."""
    import torch
    import torch.distributed as dist
    for _ in range(2):
        fe.process_batch_pinned(host_in, lens, host_out)
    barrier()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fe.process_batch_pinned(host_in, lens, host_out)
    b.record()
    barrier()
    ms = a.elapsed_time(b)
    wall_ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / steps, wall_ms / steps


def timeit_events(fn, dev, warm=3, reps=5):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(dev)
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize(dev)
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def other_configs(dev, peak, n_clips=256, seconds=180):
    """The other BASELINE.json configs and SURVEY section 8(f) rows, device resident, measured in the same run
    after the timed region (3 warm-ups, median of 5, CUDA events; inputs of 2-8 GB per step exceed L2):
    ms, audio-s/s, algorithmic GB/s and its fraction of the measured HBM peak, FP32 TFLOP/s (algorithmic
    flops, SURVEY section 8d formula) and its fraction of the 74.5 TFLOP/s CUDA-core peak."""
    import torch
    from audio_tabs_b200.frontends import log_filt_spec
    from audio_tabs_b200.onsets import OnsetStrength
    from audio_tabs_b200.plan import FrontEnd, Packed
    from audio_tabs_b200.synth import synth_batch_device
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    res = {}

    def flops_per_frame(spec):
        F = spec.frame_size
        return 2.5 * F * np.log2(F) + F + 4 * F / 2 + 2 * len(spec.filterbank.banded()[3]) + 3 * spec.num_bands

    def run(name, spec, nc, secs, dtype="f32", proj=False, what=""):
        n = int(secs * SR)
        sig = synth_batch_device(nc, n, seed=7, device=dev, dtype=dtype)
        fe = FrontEnd([spec], device=dev.index, dtype=dtype)
        packed = Packed(sig, [n] * nc, spec.hop_size)
        if proj:
            out = torch.empty((packed.total_frames, 12), dtype=torch.float32, device=dev)
            fn = lambda: fe.run_packed(packed, out=False, proj=[out])  # noqa: E731
        else:
            out = fe.alloc_output(packed.total_frames)
            fn = lambda: fe.run_packed(packed, out)  # noqa: E731
        ms = timeit_events(fn, dev)
        alg = sig.numel() * sig.element_size() + out.numel() * 4
        tf = packed.total_frames * flops_per_frame(spec) / ms / 1e9
        res[name] = {"what": what, "clips": nc, "seconds": secs, "frames": packed.total_frames, "ms": ms,
                     "audio_s_per_s": nc * secs / ms * 1e3, "alg_gb": alg / 1e9, "alg_gbs": alg / ms / 1e6,
                     "hbm_frac": alg / ms / 1e6 / peak, "fp32_tflops": tf, "fp32_frac": tf / fp32_peak}
        del sig, out, packed, fe
        torch.cuda.empty_cache()

    run("1", log_filt_spec(2048, 441.0, 12), 1, 30, what="configs[0]: 1 x 30 s, frame 2048 / hop 441 / 12 bpo -> (3000, 81) (latency bound)")
    run("3a", log_filt_spec(4096, 441.0, 24, 65.0, 2100.0, fold=True), n_clips, seconds, proj=True,
        what="configs[2] at hop 441: frame 4096 -> 24 bpo 65-2100 Hz log filterbank (87 bands) -> 12-bin chroma")
    run("3b", log_filt_spec(4096, 4410.0, 24, 65.0, 2100.0, fold=True), n_clips, seconds, proj=True,
        what="configs[2] at hop 4410 (the app's chord rate, chords/extract.py:19)")
    run("N1", log_filt_spec(8192, 4410.0, 24, 65.0, 2100.0), n_clips, seconds,
        what="DeepChroma front end: frame 8192 @ fps 10 -> (1800, 105) (chords/extract.py:54)")
    run("key", log_filt_spec(8192, 8820.0, 24, 65.0, 2100.0, int16=True), n_clips, seconds, dtype="i16",
        what="CNN key front end: int16, frame 8192 @ fps 5 -> (900, 105) (theory/key.py:101)")
    # N3: librosa-style onset strength (strum.py:114): frame 2048 hop 512, 128 mel, dB, median envelope
    n = int(seconds * SR)
    sig = synth_batch_device(n_clips, n, seed=9, device=dev)
    eng = OnsetStrength(sr=SR, aggregate=np.median, device=dev.index)
    packed = Packed(sig, [n] * n_clips, 512.0, "extend")
    ms = timeit_events(lambda: eng.envelope(packed), dev)
    alg = sig.numel() * 4 + packed.total_frames * 4
    res["N3"] = {"what": "librosa onset_strength (2048 / 512, 128 mel, power_to_db, median)", "clips": n_clips, "seconds": seconds,
                 "frames": packed.total_frames, "ms": ms, "audio_s_per_s": n_clips * seconds / ms * 1e3, "alg_gb": alg / 1e9,
                 "alg_gbs": alg / ms / 1e6, "hbm_frac": alg / ms / 1e6 / peak}
    del sig, packed, eng
    torch.cuda.empty_cache()
    return res


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from audio_tabs_b200 import _ffi
    from audio_tabs_b200.frontends import beat_specs
    from audio_tabs_b200.plan import FrontEnd, Packed
    from audio_tabs_b200.synth import synth_batch_device

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a B200; there is no CPU fallback")
    from audio_tabs_b200.sharding import bind_to_gpu_numa
    numa_cpus = bind_to_gpu_numa(local_rank)      # pinned staging buffers land on the GPU's NUMA node
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_clips = args.clips
    n_samples = int(args.clip_seconds * SR)
    specs = beat_specs()
    fe = FrontEnd(specs, device=local_rank, dtype="f32", channels=1, concurrent_streams=args.concurrent_streams,
                  one_launch=(True if args.one_launch else False if args.three_launches else None))
    sig = synth_batch_device(n_clips, n_samples, seed=2000 + rank, device=dev)
    packed = Packed(sig, [n_samples] * n_clips, fe.hop_size)
    out = fe.alloc_output(packed.total_frames)
    audio_seconds = n_clips * n_samples / SR
    in_bytes = sig.numel() * sig.element_size()
    out_bytes = out.numel() * out.element_size()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing: W warm-up, exactly K timed steps ---------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()           # nvidia-smi needs ~0.1 s to come up: start it before the warm-up
    for _ in range(args.warmup):
        fe.run_packed(packed, out)
    barrier()
    launches0 = _ffi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launch_events = []      # (resolution, start, end) of every front-end launch inside the timed region
    barrier()
    e0.record()
    for _ in range(args.steps):
        fe.run_packed(packed, out, timing=launch_events)
    e1.record()
    barrier()
    launches = _ffi.launch_count() - launches0
    elapsed_ms = e0.elapsed_time(e1)
    # the timed region is ~0.1 s: keep the SAME load running (untimed) until the sampler has a handful of
    # readings, so that none of them is taken on an idle GPU
    t_keep = time.perf_counter()
    while time.perf_counter() - t_keep < 0.6:
        fe.run_packed(packed, out)
        torch.cuda.synchronize(dev)
    clocks = sampler.stop(settle=0.0) if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed steps + 0.6 s of the same steps (GPU never idle while sampled)"
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * audio_seconds / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel: average launch duration INSIDE the timed region ----
    # (CUDA events recorded on the launching stream around every b200spec_logfilt call; each interval
    #  covers the one-block task-table kernel (~3 us) plus the persistent front-end kernel)
    peak, peak_src = load_peaks()
    traffic = load_traffic()
    per_kernel = []
    full_size = (n_clips, args.clip_seconds) == (N_CLIPS, CLIP_SECONDS)

    def spec_flops(s):
        return packed.total_frames * (2.5 * s.frame_size * np.log2(s.frame_size) + s.frame_size
                                      + 4 * s.frame_size / 2 + 2 * len(s.filterbank.banded()[3]) + 3 * s.num_bands)
    if fe.one_launch:      # ONE kernel runs all resolutions: every input sample and every output element once
        ms = float(np.mean([a.elapsed_time(b) for rr, a, b in launch_events]))
        alg = in_bytes + out_bytes
        per_kernel.append({"frame_size": "1024+2048+4096", "ms": ms, "alg_bytes": alg, "gbs": alg / ms / 1e6,
                           "fp32_tflops": sum(spec_flops(s) for s in specs) / ms / 1e9, "share_of_step": ms / ms_per_step,
                           "traffic": traffic.get("multi") if full_size else None})
    else:
        for r, s in enumerate(specs):
            ms = float(np.mean([a.elapsed_time(b) for rr, a, b in launch_events if rr == r]))
            alg = in_bytes + packed.total_frames * s.out_width * 4
            per_kernel.append({"frame_size": s.frame_size, "ms": ms, "alg_bytes": alg, "gbs": alg / ms / 1e6,
                               "fp32_tflops": spec_flops(s) / ms / 1e9, "share_of_step": ms / ms_per_step,
                               "traffic": traffic.get(str(s.frame_size)) if full_size else None})
    dom = max(per_kernel, key=lambda k: k["ms"])
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    roofline = {"bound": "hbm", "achieved": dom["gbs"], "peak": peak, "unit": "GB/s", "frac": dom["gbs"] / peak,
                "traffic": dom["traffic"], "traffic_source": traffic.get("source"),
                "peak_source": peak_src,
                "kernel": ("k_front_multi (all three resolutions in one launch" if fe.one_launch else "k_front_pair<%d> (" % dom["frame_size"])
                + "fused frame+FFT+filterbank+log+diff, two frames per complex FFT)",
                "launches_per_step": "1 front-end kernel + 1 task-table kernel" if fe.one_launch else "3 front-end kernels + 3 task-table kernels + 3 seam kernels (first diff rows of every task; inside the per-kernel times)",
                "kernel_ms": dom["ms"], "alg_bytes_per_launch": dom["alg_bytes"],
                "fp32_tflops": dom["fp32_tflops"], "fp32_frac_of_74.5": dom["fp32_tflops"] / fp32_peak,
                "note": "hop 441 makes the path FP32-issue bound (31-78 flop/B vs ridge 11); see DESIGN.md",
                "per_kernel": per_kernel,
                "step_alg_gbs": (in_bytes + out_bytes) / ms_per_step / 1e6}

    # ---- end to end: pinned host in -> device -> pinned host out, copies inside the timed region
    e2e = e2e_i16 = None
    if not args.no_e2e:
        host_in = torch.empty(sig.shape, dtype=sig.dtype, pin_memory=True)
        host_in.copy_(sig)
        host_out = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        lens = [n_samples] * n_clips
        e2e_ms, wall_ms = time_e2e(fe, host_in, lens, host_out, args.steps, barrier, dev, world)
        e2e = {"value": world * audio_seconds / (e2e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": int(out_bytes),
               "ms_per_step": e2e_ms, "wall_ms_per_step": wall_ms,
               "api": "audio_tabs_b200.plan.FrontEnd.process_batch_pinned", "numa_bound_cpus": len(numa_cpus)}
        # the copy ceiling with EVERY rank copying at once (one H2D + one D2H of the step's bytes per rank)
        fab_ms = fabric_concurrent(dev, int(in_bytes), int(out_bytes), barrier, world)
        e2e["fabric_concurrent_ms"] = fab_ms
        e2e["fabric_concurrent_value"] = world * audio_seconds / (fab_ms * 1e-3)
        e2e["frac_of_fabric_concurrent"] = e2e["value"] / e2e["fabric_concurrent_value"]
        e2e["fabric_gbs_per_gpu"] = {"h2d": in_bytes / fab_ms / 1e6, "d2h": out_bytes / fab_ms / 1e6}
        # int16 PCM ingest (the reference's audio is PCM-16 on disk: services/audio.py:18-21, pipeline.py:1672,
        # theory/key.py:144): same clips quantised to int16, window / 32767 as madmom does -- half the H2D bytes
        if not args.no_i16:
            fe16 = FrontEnd(beat_specs(int16=True), device=local_rank, dtype="i16", channels=1)
            host_in16 = torch.empty(sig.shape, dtype=torch.int16, pin_memory=True)
            host_in16.copy_((sig * 32767.0).round().clamp_(-32768, 32767).to(torch.int16))
            ms16, wall16 = time_e2e(fe16, host_in16, lens, host_out, args.steps, barrier, dev, world)
            fab16 = fabric_concurrent(dev, int(in_bytes // 2), int(out_bytes), barrier, world)
            e2e_i16 = {"value": world * audio_seconds / (ms16 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(in_bytes // 2),
                       "d2h_bytes_per_step": int(out_bytes), "ms_per_step": ms16, "wall_ms_per_step": wall16,
                       "fabric_concurrent_ms": fab16, "fabric_concurrent_value": world * audio_seconds / (fab16 * 1e-3),
                       "note": "same workload with int16 PCM input (madmom semantics: window / 32767); not the headline"}
            e2e_i16["frac_of_fabric_concurrent"] = e2e_i16["value"] / e2e_i16["fabric_concurrent_value"]
            del host_in16, fe16
        del host_in, host_out
        if rank == 0:
            pc = pcie_bandwidth(dev, int(in_bytes), int(out_bytes))
            e2e["pcie_solo_rank0"] = pc          # rank 0 alone (no other rank copying): NOT the N-GPU ceiling
            e2e["pcie_bound_value_per_gpu"] = audio_seconds / (max(pc["h2d_ms"], pc["d2h_ms"]) * 1e-3)

    configs = None
    if rank == 0 and world == 1 and not args.no_configs:
        del sig, out, packed
        torch.cuda.empty_cache()
        configs = other_configs(dev, peak)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, info, dt = cpu_front_end_throughput()
        cpu = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": "port", "sample": info["sample"], "seconds": dt,
               "blas_threads_per_worker": info["blas_threads_per_worker"], "per_core": v / info["cores"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(n_clips, args.clip_seconds, world),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_i16": e2e_i16, "configs": configs,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _dist_setup(world, local_rank):
    import torch
    import torch.distributed as dist
    from audio_tabs_b200.sharding import bind_to_gpu_numa
    numa = bind_to_gpu_numa(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms
    return dev, barrier, max_over_ranks, numa


def run_config4(args, rank, world, local_rank):
    """BASELINE configs[3]: 4096 synthetic 3-min clips job-sharded over 2 / 4 / 8 GPUs, full beat front end,
    inputs generated on the device, no inter-GPU traffic.  Every rank holds 4096 / N clips resident (512 at
    N = 8: 16.3 GB in + 11.6 GB out, 1024 at N = 4); with fewer GPUs a slice runs (1024 clips per GPU at N = 2,
    512 on one GPU -- its 4096 clips would need 222 GB) and the line says so: audio-s/s does not depend on the
    clip count on this path."""
    import torch
    import torch.distributed as dist
    from audio_tabs_b200 import _ffi
    from audio_tabs_b200.frontends import beat_specs
    from audio_tabs_b200.plan import FrontEnd, Packed
    from audio_tabs_b200.synth import synth_batch_device
    dev, barrier, max_over_ranks, _ = _dist_setup(world, local_rank)
    total_clips = 4096
    per_gpu = min(total_clips // world, 1024) if world >= 2 else 512      # <= 32.5 GB in + 23.2 GB out resident per GPU
    n = CLIP_SECONDS * SR
    specs = beat_specs()
    fe = FrontEnd(specs, device=local_rank)
    sig = synth_batch_device(per_gpu, n, seed=4000 + rank, device=dev)
    packed = Packed(sig, [n] * per_gpu, fe.hop_size)
    out = fe.alloc_output(packed.total_frames)
    for _ in range(args.warmup):
        fe.run_packed(packed, out)
    barrier()
    n0 = _ffi.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        fe.run_packed(packed, out)
    b.record()
    barrier()
    ms = max_over_ranks(a.elapsed_time(b)) / args.steps
    launches = _ffi.launch_count() - n0
    peak, peak_src = load_peaks()
    alg = sig.numel() * 4 + out.numel() * 4
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": world * per_gpu * CLIP_SECONDS / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[3]: 4096 x 180 s clips job-sharded, full beat front end -> (18000,314)/clip",
                       "clips_per_gpu": per_gpu, "clips_total_this_run": per_gpu * world,
                       "slice": None if per_gpu * world == total_clips else
                       "%d of 4096 clips (%d per GPU resident: 16.3 GB in + 11.6 GB out); throughput is clip-count invariant" % (per_gpu * world, per_gpu),
                       "parallelism": "job-sharded x%d, no collective on the data path" % world,
                       "l2": "27.9 GB per GPU per step exceed the 126 MB L2"},
            "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": peak, "unit": "GB/s", "frac": alg / ms / 1e6 / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": "whole step (three k_front_pair launches), per GPU"},
            "peak_device_memory_gb": torch.cuda.max_memory_allocated(dev) / 1e9, "gpu_launches": int(launches)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_config5(args, rank, world, local_rank):
    """BASELINE configs[4]: long-form 10-min STEREO f32 Demucs-shaped stems, 4 stems per job, 512 jobs on 8 GPUs
    (64 jobs = 54 GB of input per GPU): streamed through FrontEnd.process_batch_pinned (two device slots, three
    streams, stereo down-mix fused into the kernels' loads), never resident at once.  Host memory is bounded
    too: one pinned block of 16 stems is streamed repeatedly (the audio is synthetic anyway) -- a PIPELINE and
    memory-sizing stress, not 433 GB of distinct host data."""
    import torch
    import torch.distributed as dist
    from audio_tabs_b200.frontends import beat_specs
    from audio_tabs_b200.plan import FrontEnd
    dev, barrier, max_over_ranks, _ = _dist_setup(world, local_rank)
    seconds, jobs_per_gpu, stems_per_pass = 600, args.jobs_per_gpu, 16
    n = seconds * SR
    passes = max(1, jobs_per_gpu * 4 // stems_per_pass)
    fe = FrontEnd(beat_specs(), device=local_rank, dtype="f32", channels=2)
    gen = torch.Generator(device=dev).manual_seed(5000 + rank)
    host_in = torch.empty((stems_per_pass * n, 2), dtype=torch.float32, pin_memory=True)
    for i in range(stems_per_pass):
        host_in[i * n:(i + 1) * n].copy_(torch.randn((n, 2), generator=gen, device=dev) * 0.1)
    lens = [n] * stems_per_pass
    frames = -(-n // 441) * stems_per_pass
    host_out = torch.empty((frames, fe.width), dtype=torch.float32, pin_memory=True)
    torch.cuda.synchronize(dev)
    torch.cuda.reset_peak_memory_stats(dev)
    fe.process_batch_pinned(host_in, lens, host_out, group_clips=1)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(passes):
        fe.process_batch_pinned(host_in, lens, host_out, group_clips=1)
    b.record()
    barrier()
    ms = max_over_ranks(a.elapsed_time(b))
    stems = stems_per_pass * passes
    if rank == 0:
        print(json.dumps({
            "metric": "stem-seconds/s, beat front end on streamed 10-min stereo stems (BASELINE configs[4])",
            "value": world * stems * seconds / ms * 1e3, "unit": "audio-s/s", "n_gpus": world, "steps": passes, "warmup": 1,
            "ms_per_step": ms / passes, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[4]: %d jobs/GPU x 4 stems x 600 s stereo f32, streamed pinned host -> device -> pinned host" % jobs_per_gpu,
                       "stems_per_gpu": stems, "parallelism": "job-sharded x%d" % world,
                       "host_data": "one pinned block of 16 stems (3.4 GB) re-streamed %d times per GPU" % passes},
            "e2e": {"value": world * stems * seconds / ms * 1e3, "unit": "audio-s/s",
                    "h2d_bytes_per_step": stems_per_pass * n * 8, "d2h_bytes_per_step": frames * fe.width * 4},
            "h2d_gbs_per_gpu": stems * n * 8 / ms / 1e6, "total_ms": ms,
            "peak_device_memory_gb": torch.cuda.max_memory_allocated(dev) / 1e9,
            "pinned_host_gb_per_gpu": (host_in.numel() + host_out.numel()) * 4 / 1e9}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=N_CLIPS)
    ap.add_argument("--clip-seconds", type=float, default=CLIP_SECONDS)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-i16", action="store_true", help="skip the int16 end-to-end arm")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs (measured after the timed region)")
    ap.add_argument("--concurrent-streams", action="store_true", help="launch the three resolutions on three streams")
    ap.add_argument("--one-launch", action="store_true", help="all three resolutions in one kernel launch (b200spec_logfilt_multi)")
    ap.add_argument("--three-launches", action="store_true", help="one kernel launch per resolution (b200spec_logfilt)")
    ap.add_argument("--jobs-per-gpu", type=int, default=64, help="--config 5: jobs (of 4 stems) streamed per GPU")
    ap.add_argument("--config", type=int, default=2, choices=[2, 4, 5],
                    help="2: BASELINE configs[1] (default, the contract line); 4: configs[3] 4096 clips job-sharded; "
                         "5: configs[4] 10-min stereo stems streamed")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29541"),
               str(Path(__file__).resolve())] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.config == 4:
        run_config4(args, rank, world, local_rank)
    elif args.config == 5:
        run_config5(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
