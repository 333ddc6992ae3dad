"""The task-size model of the fused launches (b200spec_task_plan: a pure host function, no GPU needed).
W workers pull equal tasks from one counter, so a launch runs ceil(tasks / W) rounds of (chunk + overhead) frames;
the plan is the cheapest candidate under that model, optionally with short tasks from the last clips that fill the
last round ("Tasks"; measured sweeps in profiles/r02_variants.txt)."""
import ctypes as C
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

CAND = [2, 4, 8, 12, 16, 20, 24, 28, 32, 40, 48, 56, 64, 72, 80, 96]


def plan(num_sms, per_sm, overhead, total_frames, n_clips, kd=0):
    from audio_tabs_b200 import _ffi
    out = (C.c_int32 * 3)()
    _ffi.check(_ffi.lib().b200spec_task_plan(num_sms, per_sm, overhead, total_frames, n_clips, kd, out))
    return tuple(out)


def one_size_cost(c, workers, overhead, total_frames, n_clips, kd=0):
    tasks = n_clips * math.ceil(total_frames / n_clips / c)
    return max(1, math.ceil(tasks / workers)) * (c + kd + overhead)


def test_headline_batch():
    """config 2 (64 x 18000 frames) on 148 SMs: 72-frame tasks = 6.76 rounds for the 2368 warps of the warp kernels;
    frame 4096 (444 groups): 80-frame tasks, the last two clips in 16-frame tasks that fill the last round."""
    assert plan(148, 16, 0.75, 64 * 18000, 64) == (72, 72, 0)
    assert plan(148, 3, 1.0, 64 * 18000, 64) == (80, 16, 2)
    # sizes that leave the last round nearly empty are what the model avoids: 48- and 96-frame tasks are 10.15 / 5.08 rounds
    w = 148 * 16
    assert one_size_cost(48, w, 0.75, 64 * 18000, 64) > 1.04 * one_size_cost(72, w, 0.75, 64 * 18000, 64)
    assert one_size_cost(96, w, 0.75, 64 * 18000, 64) > 1.12 * one_size_cost(72, w, 0.75, 64 * 18000, 64)


def test_small_batches_get_small_tasks():
    """less than one round of work: the smallest tasks, so that every SM has something to do"""
    assert plan(148, 16, 0.75, 3000, 1) == (2, 2, 0)          # configs[0]: one 30 s clip
    assert plan(148, 3, 1.0, 300, 1)[0] == 2
    assert plan(148, 3, 1.0, 0, 0)[2] == 0                    # nothing to do: any plan, no tail


def test_warm_up_rows_keep_whole_tail_batches():
    """a task transforms chunk + warm-up rows: a multiple of four from 16 frames on, whole pairs below"""
    for kd in (1, 2, 3, 5):
        for frames, clips in ((64 * 18000, 64), (3000, 1), (50000, 7)):
            c, cs, tail = plan(148, 3, 1.0, frames, clips, kd)
            assert (c + kd) % (4 if c >= 16 else 2) == 0 and (cs + kd) % (4 if cs >= 16 else 2) == 0


@settings(max_examples=300, deadline=None)
@given(st.integers(1, 4000), st.integers(1, 40000), st.sampled_from([(16, 0.75), (3, 1.0), (4, 1.0), (5, 1.0)]),
       st.sampled_from([132, 148, 160]))
def test_plan_is_the_cheapest_under_the_model(n_clips, per_clip, worker, num_sms):
    per_sm, overhead = worker
    total = n_clips * per_clip
    workers = num_sms * per_sm
    c, cs, tail = plan(num_sms, per_sm, overhead, total, n_clips)
    assert c in CAND and cs in CAND and c % 2 == 0 and cs % 2 == 0
    best_one = min(one_size_cost(k, workers, overhead, total, n_clips) for k in CAND)
    if tail == 0:
        assert cs == c
        assert one_size_cost(c, workers, overhead, total, n_clips) <= best_one * (1 + 1e-9)
    else:
        # two sizes: only with enough clips, the tail holds at least one round of long tasks and at most half the batch,
        # and the modelled cost beats every single size
        assert n_clips >= 8 and 1 <= tail <= n_clips // 2 and cs == 16 and c >= 48
        assert tail * per_clip >= workers * c - per_clip
        long_tasks = (n_clips - tail) * math.ceil(per_clip / c)
        short_tasks = tail * math.ceil(per_clip / cs)
        assert long_tasks >= 2 * workers
        cost = (long_tasks * (c + overhead) + short_tasks * (cs + overhead)) / workers + (cs + overhead)
        assert cost < best_one


def test_bad_arguments():
    from audio_tabs_b200 import _ffi
    out = (C.c_int32 * 3)()
    lib = _ffi.lib()
    assert lib.b200spec_task_plan(0, 16, 0.75, 100, 1, 0, out) != 0
    assert lib.b200spec_task_plan(148, 16, -1.0, 100, 1, 0, out) != 0
    assert lib.b200spec_task_plan(148, 16, 0.75, -5, 1, 0, out) != 0
    assert lib.b200spec_task_plan(148, 16, 0.75, 100, 1, 0, None) != 0
