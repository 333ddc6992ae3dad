#!/usr/bin/env python
"""BASELINE configs[4] / configs[3] at (scaled) full size -- memory-sizing and streaming check.

    python tools/bench_config5.py [--jobs-per-gpu 64] [--seconds 600]          (1 GPU)
    python -m torch.distributed.run --nproc-per-node 8 ... tools/bench_config5.py   (8 GPUs, job-sharded)

Config 5: every job = 4 Demucs-shaped stems, 10 min, stereo float32 (212 MB per stem).  64 jobs per GPU
are 54 GB of input: they STREAM through FrontEnd.process_batch_pinned (two device slots, three streams,
stereo down-mix fused into the kernels' loads) and are never resident at once.  Host memory is bounded
too: one pinned block of 4 jobs is streamed repeatedly (the audio is synthetic anyway).  Prints one JSON
line per rank-0: stem-seconds per second, bytes moved, peak device memory.
"""
import argparse, json, os, sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audio_tabs_b200.frontends import beat_specs
from audio_tabs_b200.plan import FrontEnd
from audio_tabs_b200.sharding import bind_to_gpu_numa

SR = 44100


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs-per-gpu", type=int, default=64)
    ap.add_argument("--seconds", type=int, default=600)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n = args.seconds * SR
    stems_per_pass = 16                                  # 4 jobs x 4 stems held in pinned memory
    passes = max(1, args.jobs_per_gpu * 4 // stems_per_pass)
    fe = FrontEnd(beat_specs(), device=local, dtype="f32", channels=2)
    gen = torch.Generator(device=dev).manual_seed(5000 + rank)
    host_in = torch.empty((stems_per_pass * n, 2), dtype=torch.float32, pin_memory=True)
    for i in range(stems_per_pass):                      # fill stem by stem (bounded device memory)
        host_in[i * n:(i + 1) * n].copy_(torch.randn((n, 2), generator=gen, device=dev) * 0.1)
    lens = [n] * stems_per_pass
    frames = -(-n // 441) * stems_per_pass
    host_out = torch.empty((frames, fe.width), dtype=torch.float32, pin_memory=True)
    torch.cuda.synchronize(dev)
    torch.cuda.reset_peak_memory_stats(dev)
    fe.process_batch_pinned(host_in, lens, host_out, group_clips=1)        # warm-up pass
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(passes):
        fe.process_batch_pinned(host_in, lens, host_out, group_clips=1)
    b.record()
    torch.cuda.synchronize(dev)
    ms = a.elapsed_time(b)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    stems = stems_per_pass * passes
    if rank == 0:
        print(json.dumps({
            "config": "5: %d jobs/GPU x 4 stems x %d s stereo f32, streamed (pinned host -> device -> pinned host)" % (args.jobs_per_gpu, args.seconds),
            "n_gpus": world, "stems_per_gpu": stems, "ms": ms,
            "stem_seconds_per_s": world * stems * args.seconds / ms * 1e3,
            "h2d_gb_per_gpu": stems * n * 8 / 1e9, "d2h_gb_per_gpu": stems * (frames // stems_per_pass) * fe.width * 4 / 1e9,
            "h2d_gbs_per_gpu": stems * n * 8 / ms / 1e6,
            "peak_device_memory_gb": torch.cuda.max_memory_allocated(dev) / 1e9,
            "pinned_host_gb_per_gpu": (host_in.numel() + host_out.numel()) * 4 / 1e9}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
