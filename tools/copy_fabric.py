#!/usr/bin/env python
"""Host <-> device copy fabric of one box with ALL ranks copying at once, for different kinds of host memory.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/copy_fabric.py [--mb 1024]

Every rank (one per GPU, bound to the GPU's NUMA CPUs) times, after a barrier, one H2D copy, one D2H copy and
both together (two streams) of `--mb` MiB, max over ranks, best of 3, for host buffers that are
  pinned      torch pin_memory (cudaHostAlloc, default flags)
  wc          cudaHostAlloc(..., cudaHostAllocWriteCombined)   (H2D source only: CPU reads of WC memory are slow)
  registered  anonymous mmap + madvise(MADV_HUGEPAGE) + cudaHostRegister (transparent huge pages where the kernel grants them)
Rank 0 prints one JSON line per kind with per-GPU and aggregate GB/s.  Explains why the streamed (e2e) path
stops scaling beyond two GPUs and what, if anything, a different allocation buys.
"""
import ctypes, json, mmap, os, sys
from pathlib import Path
import torch
import torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audio_tabs_b200.sharding import bind_to_gpu_numa

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
MB = int(sys.argv[sys.argv.index("--mb") + 1]) if "--mb" in sys.argv else 1024
NB = MB << 20
cpus = bind_to_gpu_numa(local)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")


def as_tensor(ptr, nbytes):
    buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
    return torch.frombuffer(buf, dtype=torch.uint8)


def alloc(kind):
    """(tensor, release)"""
    if kind == "pinned":
        t = torch.empty(NB, dtype=torch.uint8, pin_memory=True)
        return t, lambda: None
    if kind == "wc":
        p = ctypes.c_void_p()
        rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(NB), ctypes.c_uint(0x04))   # cudaHostAllocWriteCombined
        if rc != 0:
            raise RuntimeError("cudaHostAlloc(WC) -> %d" % rc)
        return as_tensor(p.value, NB), lambda: rt.cudaFreeHost(p)
    if kind == "registered":
        m = mmap.mmap(-1, NB, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        try:
            m.madvise(mmap.MADV_HUGEPAGE)
        except Exception:
            pass
        t = torch.frombuffer(m, dtype=torch.uint8)
        t.zero_()                                                       # touch: pages land on this rank's NUMA node
        rc = rt.cudaHostRegister(ctypes.c_void_p(t.data_ptr()), ctypes.c_size_t(NB), ctypes.c_uint(0))
        if rc != 0:
            raise RuntimeError("cudaHostRegister -> %d" % rc)
        return t, lambda: rt.cudaHostUnregister(ctypes.c_void_p(t.data_ptr()))
    raise ValueError(kind)


def timed(fn):
    best = None
    for i in range(4):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize(dev)
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        if i > 0:
            best = ms if best is None else min(best, ms)
    return best


d_in = torch.empty(NB, dtype=torch.uint8, device=dev)
d_out = torch.empty(NB, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
for kind in ("pinned", "wc", "registered"):
    try:
        h_in, rel_in = alloc(kind)
        h_out, rel_out = alloc("pinned" if kind == "wc" else kind)
    except Exception as exc:   # noqa: BLE001
        if rank == 0:
            print(json.dumps({"kind": kind, "error": str(exc)}), flush=True)
        continue

    def both():
        cur = torch.cuda.current_stream(dev)
        e = torch.cuda.Event(); e.record(cur)
        s1.wait_event(e); s2.wait_event(e)
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True); e1 = torch.cuda.Event(); e1.record(s1)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True); e2 = torch.cuda.Event(); e2.record(s2)
        cur.wait_event(e1); cur.wait_event(e2)

    ms_h = timed(lambda: d_in.copy_(h_in, non_blocking=True))
    ms_d = timed(lambda: h_out.copy_(d_out, non_blocking=True))
    ms_b = timed(both)
    if rank == 0:
        g = NB / 1e6
        print(json.dumps({"kind": kind, "n_gpus": world, "mib": MB, "numa_cpus": len(cpus),
                          "h2d_gbs_per_gpu": round(g / ms_h, 1), "d2h_gbs_per_gpu": round(g / ms_d, 1),
                          "both_gbs_per_gpu_each_dir": round(g / ms_b, 1),
                          "aggregate_both_gbs": round(2 * g * world / ms_b, 1)}), flush=True)
    del h_in, h_out
    rel_in(); rel_out()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
