#!/usr/bin/env python
"""Timeline of one FrontEnd.process_batch_pinned pass: CUDA-event timestamps of every group's H2D copy,
kernels and D2H copy (ms since the start of the pass) -- where the pipeline's bubbles are.

    python tools/e2e_timeline.py [--i16] [--group 2] [--slots 2]
"""
import json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audio_tabs_b200.frontends import beat_specs
from audio_tabs_b200.plan import FrontEnd
from audio_tabs_b200.synth import synth_batch_device

SR, NC, SEC = 44100, 64, 180
DT = "i16" if "--i16" in sys.argv else "f32"
arg = lambda k, d: int(sys.argv[sys.argv.index(k) + 1]) if k in sys.argv else d   # noqa: E731
GROUP, SLOTS = arg("--group", 2), arg("--slots", 2)
dev = torch.device("cuda", 0)
n = SEC * SR
fe = FrontEnd(beat_specs(int16=(DT == "i16")), device=0, dtype=DT)
sig = synth_batch_device(NC, n, seed=2000, device=dev, dtype=DT)
host_in = torch.empty(sig.shape, dtype=sig.dtype, pin_memory=True); host_in.copy_(sig); del sig
host_out = torch.empty((18000 * NC, fe.width), dtype=torch.float32, pin_memory=True)
lens = [n] * NC
for _ in range(3):
    fe.process_batch_pinned(host_in, lens, host_out, group_clips=GROUP, n_slots=SLOTS)
torch.cuda.synchronize()

# the same schedule as process_batch_pinned, with timing events around every operation
pipe = fe._pipeline(lens, GROUP, SLOTS)
h2d, comp, d2h = pipe["h2d"], pipe["comp"], pipe["d2h"]
E = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731
cur = torch.cuda.current_stream(dev)
t0 = E(); t0.record(cur)
for s in (h2d, comp, d2h):
    s.wait_event(t0)
ev_comp, ev_d2h, marks = [None] * SLOTS, [None] * SLOTS, []
for gi, g in enumerate(pipe["groups"]):
    slot = gi % SLOTS
    buf = pipe["slots"][slot]
    m = {}
    if ev_comp[slot] is not None:
        h2d.wait_event(ev_comp[slot])
    with torch.cuda.stream(h2d):
        m["h0"] = E(); m["h0"].record(h2d)
        buf["sig"][:g["nsamp"]].copy_(host_in[g["samp0"]:g["samp0"] + g["nsamp"]], non_blocking=True)
        m["h1"] = E(); m["h1"].record(h2d)
    comp.wait_event(m["h1"])
    if ev_d2h[slot] is not None:
        comp.wait_event(ev_d2h[slot])
    with torch.cuda.stream(comp):
        m["c0"] = E(); m["c0"].record(comp)
        fe.run_packed(g["packed"][slot], buf["out"][:g["rows"]])
        m["c1"] = E(); m["c1"].record(comp)
        ev_comp[slot] = m["c1"]
    d2h.wait_event(m["c1"])
    with torch.cuda.stream(d2h):
        m["d0"] = E(); m["d0"].record(d2h)
        host_out[g["row0"]:g["row0"] + g["rows"]].copy_(buf["out"][:g["rows"]], non_blocking=True)
        m["d1"] = E(); m["d1"].record(d2h)
        ev_d2h[slot] = m["d1"]
    marks.append(m)
torch.cuda.synchronize()
rows = []
for gi, m in enumerate(marks):
    rows.append({k: round(t0.elapsed_time(v), 3) for k, v in m.items()})
tot = rows[-1]["d1"]
busy = lambda a, b: sum(r[b] - r[a] for r in rows)   # noqa: E731
print(json.dumps({"dtype": DT, "group": GROUP, "slots": SLOTS, "total_ms": tot, "h2d_busy_ms": round(busy("h0", "h1"), 2),
                  "comp_busy_ms": round(busy("c0", "c1"), 2), "d2h_busy_ms": round(busy("d0", "d1"), 2),
                  "mean_h2d_ms": round(busy("h0", "h1") / len(rows), 3), "mean_comp_ms": round(busy("c0", "c1") / len(rows), 3),
                  "mean_d2h_ms": round(busy("d0", "d1") / len(rows), 3)}))
for gi, r in enumerate(rows[:6] + rows[-4:]):
    print(gi if gi < 6 else len(rows) - 10 + gi, r)
