"""CPU, world_size 2 over gloo: the host-side multi-GPU logic — contiguous sharding of the problems and the
single all-gather of the solved records (davo_b200.distributed).  The CUDA kernel is stood in for by the C
oracle writing into the rank's slab; what is under test is the plumbing around it."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import davo_b200
from davo_b200.distributed import ResultSlab, shard_range
from oracle import c_oracle


def test_shard_range_covers_everything_once():
    for total in (0, 1, 7, 64, 65536, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0].lo == 0 and spans[-1].hi == total
            assert all(a.hi == b.lo for a, b in zip(spans, spans[1:]))
            sizes = [s.size for s in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_result_slab_views_are_disjoint_and_aligned():
    slab = ResultSlab(10, 10, torch.float32, 3, "cpu")  # shards of 4, 3, 3
    assert slab.rows == 4 and slab.slab_bytes % 256 == 0
    for r, rows in enumerate((4, 3, 3)):
        b = slab.buffers(r)
        assert b.x.shape == (rows, 10) and b.cost.shape == (rows,) and b.converged.dtype == torch.uint8
        b.x.fill_(r + 1.0); b.cost.fill_(r + 10.0); b.iterations.fill_(r + 20); b.evaluations.fill_(r + 30)
        b.reason.fill_(r); b.converged.fill_(1)
    g = slab.gathered()
    assert g.x.shape == (10, 10)
    assert torch.equal(g.x[:, 0], torch.tensor([1.0] * 4 + [2.0] * 3 + [3.0] * 3))
    assert torch.equal(g.iterations, torch.tensor([20] * 4 + [21] * 3 + [22] * 3, dtype=torch.int32))
    assert torch.equal(g.reason, torch.tensor([0] * 4 + [1] * 3 + [2] * 3, dtype=torch.int32))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch = davo_b200.synthetic.make_distort10(total, 32, seed=5, dtype=np.float64)
        span = shard_range(total, rank, world)
        mine = batch.slice(span.lo, span.hi)
        r = c_oracle.solve_batch(mine, threads=1, error_threshold=1e-12, iterations=100)
        slab = ResultSlab(total, 10, torch.float64, world, "cpu")
        b = slab.buffers(rank)
        b.x.copy_(torch.from_numpy(r["x"])); b.cost.copy_(torch.from_numpy(r["cost"]))
        b.iterations.copy_(torch.from_numpy(r["iters"])); b.evaluations.copy_(torch.from_numpy(r["fevals"]))
        b.reason.copy_(torch.from_numpy(r["reason"])); b.converged.copy_(torch.from_numpy(r["converged"].astype(np.uint8)))
        slab.all_gather(rank)  # the one collective of the path
        g = slab.gathered()
        torch.save({"x": g.x, "cost": g.cost, "iters": g.iterations, "reason": g.reason, "conv": g.converged},
                   f"{out_path}.{rank}")
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_equals_single_rank_solve(tmp_path):
    total, world = 37, 2  # odd: the shards differ in size, slabs do not
    out = str(tmp_path / "gathered")
    mp.start_processes(_worker, args=(world, _free_port(), total, out), nprocs=world, join=True, start_method="spawn")
    batch = davo_b200.synthetic.make_distort10(total, 32, seed=5, dtype=np.float64)
    ref = c_oracle.solve_batch(batch, threads=1, error_threshold=1e-12, iterations=100)
    for rank in range(world):  # every rank ends up with every record, in global order
        g = torch.load(f"{out}.{rank}")
        assert np.array_equal(g["x"].numpy(), ref["x"])
        assert np.array_equal(g["cost"].numpy(), ref["cost"])
        assert np.array_equal(g["iters"].numpy(), ref["iters"])
        assert np.array_equal(g["reason"].numpy(), ref["reason"])
        assert np.array_equal(g["conv"].numpy().astype(bool), ref["converged"])
