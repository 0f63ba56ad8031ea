"""Aggregate host-to-device ceiling of this box at N = 1, 2, 4, 8 concurrent ranks: plain cudaMemcpyAsync (through
torch's copy_) of the SAME bytes per rank and step as bench.py's e2e arm (64K x 256 matches: 338 MB of raw inputs),
from pinned memory, in the solver's chunking (one stream, one copy per 4096-problem chunk) and as one copy.

    python tools/h2d_probe.py                                        # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/h2d_probe.py                                           # N ranks at once

Each rank copies to its own GPU at the same time (barrier before the timed region, max over ranks after).  Prints one
JSON line: per-rank and aggregate GB/s.  bench.py's e2e `h2d_gbs_per_gpu` is compared with this in DESIGN.md section 6:
if the solve path reaches the plain-copy figure, e2e is bound by the platform's PCIe / host-memory path.
"""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

BYTES = 65536 * 256 * 20 + 65536 * 10 * 4      # points (12 B) + observations (8 B) per match, x0 per problem
n = BYTES // 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
h.normal_()
d = torch.empty(n, dtype=torch.float32, device="cuda")


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def timed(chunks: int, reps: int = 10) -> float:
    cuts = [n * i // chunks for i in range(chunks + 1)]
    s = torch.cuda.Stream()
    best = []
    for rep in range(reps + 2):
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s):
            for lo, hi in zip(cuts[:-1], cuts[1:]):
                d[lo:hi].copy_(h[lo:hi], non_blocking=True)
        s.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        if rep >= 2:
            best.append(dt)
    return sorted(best)[len(best) // 2]


out = {"n_ranks": world, "bytes_per_rank": BYTES}
for label, chunks in (("one_copy", 1), ("chunks_of_4096_problems", 16)):
    dt = timed(chunks)
    out[label] = {"ms": dt * 1e3, "gbs_per_rank": BYTES / dt / 1e9, "gbs_aggregate": world * BYTES / dt / 1e9}
try:
    import pynvml
    pynvml.nvmlInit()
    hnd = pynvml.nvmlDeviceGetHandleByIndex(local)
    out["pcie"] = {"gen": pynvml.nvmlDeviceGetCurrPcieLinkGeneration(hnd), "width": pynvml.nvmlDeviceGetCurrPcieLinkWidth(hnd)}
except Exception as e:  # noqa: BLE001
    out["pcie"] = f"unavailable: {type(e).__name__}"
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
