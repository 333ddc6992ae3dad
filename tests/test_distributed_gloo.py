"""N > 1 host logic on CPU: world_size-2 gloo processes shard the job list (no data-path collective)
and reduce the three reporting scalars, exactly what bench.py does over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, lengths, q):
    import torch.distributed as dist
    from audio_tabs_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.local_shard(lengths, rank, world)
    audio = sum(lengths[i] for i in mine) / 44100.0
    tot_audio, max_elapsed, tot_bytes = sharding.reduce_stats(audio, 0.5 + rank, 4.0 * sum(lengths[i] for i in mine))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, mine, tot_audio, max_elapsed, tot_bytes))


@pytest.mark.timeout(120)
def test_two_rank_sharding_and_stats():
    lengths = [441000, 882000, 441000, 220500, 1323000, 441000]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (r0, s0, a0, e0, b0), (r1, s1, a1, e1, b1) = results
    assert sorted(s0 + s1) == list(range(len(lengths))) and not set(s0) & set(s1)
    assert a0 == a1 == pytest.approx(sum(lengths) / 44100.0)
    assert e0 == e1 == 1.5 and b0 == b1 == 4.0 * sum(lengths)
    loads = [sum(lengths[i] for i in s) for s in (s0, s1)]
    assert abs(loads[0] - loads[1]) <= max(lengths)
