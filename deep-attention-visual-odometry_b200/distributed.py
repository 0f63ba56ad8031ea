"""Multi-GPU plumbing: problems are independent, so the batch shards by contiguous ranges of problems,
one process per GPU, and the ONLY collective is one all-gather of the solved records.

The reference has no distributed code at all; this is the B200-side equivalent of
running BFGSSolver on a batch larger than one device.  The solve kernel writes its outputs directly
into this rank's slot of the gather buffer (no pack kernel), and a single
``all_gather_into_tensor`` over NCCL / NVLink moves every rank's slab to every rank.
"""
from __future__ import annotations

from typing import NamedTuple

import torch
import torch.distributed as dist

from . import _lib
from .solvers import SolveBuffers


class Shard(NamedTuple):
    lo: int
    hi: int

    @property
    def size(self) -> int:
        return self.hi - self.lo


def shard_range(total: int, rank: int, world: int) -> Shard:
    """Contiguous, balanced: the first `total % world` ranks own one extra problem."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return Shard(lo, lo + base + (1 if rank < extra else 0))


def _align(x: int, a: int = 256) -> int:
    return (x + a - 1) // a * a


class ResultSlab:
    """One byte buffer [world, slab_bytes]; slab r holds rank r's x | cost | iterations | evaluations |
    reason | converged segments, each 256-byte aligned and sized for the LARGEST shard so every slab has
    the same length (all_gather_into_tensor needs equal contributions)."""

    def __init__(self, total: int, n: int, dtype: torch.dtype, world: int, device):
        self.total, self.n, self.dtype, self.world, self.device = int(total), int(n), dtype, int(world), device
        self.rows = shard_range(total, 0, world).size  # largest shard
        item = torch.empty((), dtype=dtype).element_size()
        sizes = [self.rows * n * item, self.rows * item, self.rows * 4, self.rows * 4, self.rows * 4, self.rows]
        self.offsets = []
        off = 0
        for s in sizes:
            self.offsets.append(off)
            off += _align(s)
        self.offsets.append(off)
        self.slab_bytes = off
        self.buffer = torch.zeros(self.world, self.slab_bytes, dtype=torch.uint8, device=device)
        self.workspace = torch.empty(_lib.WORKSPACE_BYTES, dtype=torch.uint8, device=device)

    def _segment(self, slab: torch.Tensor, i: int, dtype: torch.dtype, rows: int, cols: int = 0):
        item = torch.empty((), dtype=dtype).element_size()
        count = rows * (cols or 1)
        seg = slab[self.offsets[i]: self.offsets[i] + count * item].view(dtype)
        return seg.view(rows, cols) if cols else seg

    def buffers(self, rank: int) -> SolveBuffers:
        """Views of rank `rank`'s slab with exactly that rank's shard size (what the kernel writes into)."""
        rows = shard_range(self.total, rank, self.world).size
        slab = self.buffer[rank]
        return SolveBuffers(self._segment(slab, 0, self.dtype, rows, self.n), self._segment(slab, 1, self.dtype, rows),
                            self._segment(slab, 5, torch.uint8, rows), self._segment(slab, 2, torch.int32, rows),
                            self._segment(slab, 3, torch.int32, rows), self._segment(slab, 4, torch.int32, rows),
                            self.workspace)

    def all_gather(self, rank: int, group=None) -> None:
        """The single collective of the path.  In place: rank's own slab is the send buffer."""
        if self.world == 1:
            return
        dist.all_gather_into_tensor(self.buffer.view(-1), self.buffer[rank], group=group)

    def gathered(self) -> SolveBuffers:
        """All problems in global order (a small gather copy across slabs, after the collective)."""
        parts = [self.buffers(r) for r in range(self.world)]
        cat = lambda k: torch.cat([getattr(p, k) for p in parts], dim=0)
        return SolveBuffers(cat("x"), cat("cost"), cat("converged"), cat("iterations"), cat("evaluations"),
                            cat("reason"), self.workspace)
