#!/usr/bin/env python
"""Yardstick, not product: the headline step (config 2: 64 clips x 180 s, frames 1024 / 2048 / 4096 at hop 441 ->
(T, 314)) written with LIBRARY kernels -- torch ops over cuFFT (rfft) and cuBLAS (filterbank matmul) -- timed on the
same box and the same resident input as the fused kernels of `libb200spec.so`.

It answers two questions the roofline fraction cannot: (1) what does a straightforward GPU port of the madmom chain
cost, and (2) how long does cuFFT ALONE take on frames that are already framed and windowed in HBM (no framing, no
magnitude, no filterbank, no log, no difference) -- the part of the work a fused kernel cannot avoid.

Prints one JSON line per resolution and a summary; the library result is checked against the fused output
(max abs difference) so both arms provably compute the same thing.

    python tools/library_yardstick.py [--clips 64] [--seconds 180] [--chunk 8]
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audio_tabs_b200.frontends import beat_specs          # noqa: E402
from audio_tabs_b200.plan import FrontEnd, Packed          # noqa: E402
from audio_tabs_b200.synth import synth_batch_device       # noqa: E402

SR = 44100


def events(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=180.0)
    ap.add_argument("--chunk", type=int, default=8, help="clips per library pass (bounds the framed intermediates)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    n = int(args.seconds * SR)
    nc = args.clips
    sig = synth_batch_device(nc, n, seed=2000, device=dev)                 # nc clips of n float32 samples, resident
    sig2 = sig.view(nc, n)
    specs = beat_specs()
    hop = 441
    T = int(np.ceil(n / hop))                                               # madmom: ceil(len / hop) frames

    # ---- fused arm -------------------------------------------------------------------------------------------
    fe = FrontEnd(specs, device=0, dtype="f32")
    packed = Packed(sig, [n] * nc, fe.hop_size)
    assert packed.total_frames == nc * T
    out = fe.alloc_output(packed.total_frames)
    ms_fused = events(lambda: fe.run_packed(packed, out))
    ms_fused_res = []
    for s in specs:
        f1 = FrontEnd([s], device=0, dtype="f32")
        o1 = f1.alloc_output(packed.total_frames)
        ms_fused_res.append(events(lambda: f1.run_packed(packed, o1)))
        del f1, o1

    # ---- library arm -----------------------------------------------------------------------------------------
    tabs = []
    for s in specs:
        tabs.append((s.frame_size, torch.from_numpy(s.window32).to(dev),
                     torch.from_numpy(np.ascontiguousarray(np.asarray(s.filterbank), dtype=np.float32)).to(dev),
                     s.diff_frames))
    lib_out = torch.empty((nc * T, fe.width), dtype=torch.float32, device=dev)

    def frames_of(x, F):
        # madmom FramedSignal (origin 0): frame f is centred on sample f * hop, zeros outside the clip
        xp = torch.nn.functional.pad(x, (F // 2, F // 2 + hop))
        return xp.unfold(1, F, hop)[:, :T]                                  # (clips, T, F) strided view

    def library_res(x, k, dst):
        F, win, fb, kd = tabs[k]
        fr = frames_of(x, F) * win                                          # materialises (clips, T, F)
        mag = torch.fft.rfft(fr, dim=-1)[..., : F // 2].abs()
        lg = torch.log10(torch.matmul(mag, fb) + 1.0)                       # (clips, T, B)
        B = lg.shape[-1]
        d = torch.zeros_like(lg)
        d[:, kd:] = (lg[:, kd:] - lg[:, :-kd]).clamp_min_(0.0)
        dst[..., :B] = lg
        dst[..., B:2 * B] = d

    col = [0]
    for s in specs:
        col.append(col[-1] + 2 * s.num_bands)

    def library_step(which=(0, 1, 2)):
        v = lib_out.view(nc, T, fe.width)
        for c0 in range(0, nc, args.chunk):
            x = sig2[c0:c0 + args.chunk]
            for k in which:
                library_res(x, k, v[c0:c0 + args.chunk, :, col[k]:col[k + 1]])

    ms_lib = events(library_step, reps=3, warm=2)
    ms_lib_res = [events(lambda k=k: library_step((k,)), reps=3, warm=2) for k in range(3)]
    library_step()
    torch.cuda.synchronize()
    fe.run_packed(packed, out)
    torch.cuda.synchronize()
    diff = float((lib_out - out).abs().max())

    # ---- cuFFT alone on frames already framed + windowed in HBM ---------------------------------------------
    ms_fft = []
    for k, s in enumerate(specs):
        F = s.frame_size
        ncl = min(nc, args.chunk * 2)
        fr = (frames_of(sig2[:ncl], F) * tabs[k][1]).contiguous()
        ms = events(lambda: torch.fft.rfft(fr, dim=-1), reps=5, warm=3)
        ms_fft.append(ms * nc / ncl)                                        # scaled to the whole step
        del fr
        torch.cuda.empty_cache()

    for k, s in enumerate(specs):
        print(json.dumps({"frame_size": s.frame_size, "frames": nc * T, "fused_ms": round(ms_fused_res[k], 3),
                          "library_ms": round(ms_lib_res[k], 3), "cufft_rfft_alone_ms": round(ms_fft[k], 3),
                          "library_over_fused": round(ms_lib_res[k] / ms_fused_res[k], 2),
                          "cufft_alone_over_fused": round(ms_fft[k] / ms_fused_res[k], 2)}), flush=True)
    print(json.dumps({"step": "config 2", "clips": nc, "seconds": args.seconds, "fused_ms": round(ms_fused, 3),
                      "library_ms": round(ms_lib, 3), "library_over_fused": round(ms_lib / ms_fused, 2),
                      "cufft_rfft_alone_ms": round(sum(ms_fft), 3),
                      "cufft_alone_over_fused": round(sum(ms_fft) / ms_fused, 2),
                      "max_abs_diff_library_vs_fused": diff,
                      "note": "library arm = torch ops (cuFFT rfft, cuBLAS matmul, elementwise kernels), "
                              "%d clips per pass; cuFFT alone = rfft of frames already framed and windowed in HBM, "
                              "nothing else" % args.chunk}), flush=True)


if __name__ == "__main__":
    main()
