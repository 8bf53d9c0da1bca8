"""Static sharding of the hot path across the GPUs of one box (one process per GPU).

Keys are independent and EvalAll subtrees below any node depend only on that node (dpf.cuh:170-214,
:291-301), so the path shards with NO data-path collective: rank r owns a contiguous key range
(point evaluation) or a contiguous leaf range of every key (full-domain evaluation, BASELINE
config 4 "subtrees sharded across 8 GPUs").  The only optional communication is a final gather of
the (small) point-evaluation outputs.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def key_shard(nkeys: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the keys rank ``rank`` evaluates: contiguous, sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(nkeys, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def leaf_shard(in_bits: int, granule: int, rank: int, world: int) -> Tuple[int, int]:
    """(leaf_begin, leaf_count) of rank ``rank``'s slice of the 2^in_bits leaves of every key.

    The slice is a whole number of EvalAll work units (``granule`` = Context.granule(), a power of
    two).  When the domain has fewer units than ranks, trailing ranks get an empty slice."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    n_leaves = 1 << in_bits
    if granule <= 0 or n_leaves % granule:
        raise ValueError("granule must divide the domain size")
    units = n_leaves // granule
    b, e = key_shard(units, rank, world)
    return b * granule, (e - b) * granule


def gather_point_outputs(ys_local: torch.Tensor, nkeys: int, group: Optional[dist.ProcessGroup] = None
                         ) -> Optional[torch.Tensor]:
    """Optional final gather of point-evaluation outputs ((n_local, 4) int32) to every rank, in key
    order.  16 B x 2^22 keys = 64 MiB in total: latency-, not bandwidth-bound on NVLink 5."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return ys_local
    world = dist.get_world_size(group)
    sizes = [e - b for b, e in (key_shard(nkeys, r, world) for r in range(world))]
    rows = max(sizes)  # shards differ by at most one row: pad to equal size for the collective
    mine = torch.zeros((rows, ys_local.shape[1]), dtype=ys_local.dtype, device=ys_local.device)
    mine[: ys_local.shape[0]] = ys_local
    bufs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(bufs, mine, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)
