"""Multi-GPU plumbing (one process per GPU, torch.distributed): host-side logic only.

Two modes:

* read-partitioned, replicated table — `read_range`: contiguous read ranges per
  rank, no collective on the data path;
* table-partitioned — what the reference does for `-d N`
  (src/CuClarkDB.cu:546-574: every device receives every read batch, :886-895, and
  the sparse per-read rows are merged pairwise to device 0, :953-974). Here every
  rank holds one shard of the device table, classifies ALL reads against it and the
  rows are exchanged with ONE all-to-all so that rank r ends up with the G partial
  rows of ITS reads, merged locally by `cuclark_merge_rows_device`
  (mergeKernel + resultKernel generalised to G inputs).

The functions work on torch tensors of any device, so the exchange logic is
tested on CPU with the gloo backend (tests/test_multigpu_gloo.py) and runs over
NCCL/NVLink on the GPUs (bench.py --mode table).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def read_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice of the reads owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_fixed_reads(cont_local: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """All-gather equally sized packed-read buffers: every rank gets every rank's containers
    (the reference copies each batch to every device, src/CuClarkDB.cu:886-895)."""
    if world == 1:
        return cont_local
    src = cont_local.contiguous().view(torch.uint8)          # bytes: every backend moves uint8
    out = torch.empty(world * src.numel(), dtype=torch.uint8, device=src.device)
    dist.all_gather_into_tensor(out, src, group=group)
    return out.view(cont_local.dtype)


def exchange_rows(rows_all: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """rows_all: [world * n, pitch] = this shard's partial rows for ALL reads, rank-major.
    Returns [world, n, pitch]: for the n reads this rank owns, the partial row of every shard."""
    n = rows_all.shape[0] // world
    if world == 1:
        return rows_all.view(1, n, -1)
    pitch = rows_all.shape[1]
    src = rows_all.contiguous().view(torch.uint8)            # bytes: every backend moves uint8
    if rows_all.device.type == "cpu":
        # gloo (CPU tests) has no all-to-all: gather everything and keep this rank's slice
        rank = dist.get_rank(group)
        everything = [torch.empty_like(src) for _ in range(world)]
        dist.all_gather(everything, src, group=group)
        recv = torch.stack([e.view(world, -1)[rank] for e in everything])
    else:
        recv = torch.empty_like(src)
        dist.all_to_all_single(recv, src, group=group)
    return recv.view(rows_all.dtype).view(world, n, pitch)
