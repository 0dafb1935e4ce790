"""Multi-GPU deployment of the hot path: shard-by-call (option A).

Reads shard naturally and the index is replicated in every GPU's HBM, so the unit of distribution is one
`createAlignments` call (= one batch = one insert-size domain, read ids counted from 0 -- exactly what the reference
does per `BwaMemAligner.alignSeqs` call, /root/reference/src/main/c/jnibwa.c:197-235).  Rank r of W (one process per
GPU) takes calls r, r+W, r+2W, ...; no collective touches the data path.  The only exchange is the host-side ordered
gather of the result streams, below -- results are byte-identical to a single-GPU run by construction because no
call is ever split."""
from __future__ import annotations

from typing import Dict, List, Optional


def calls_of_rank(n_calls: int, rank: int, world: int) -> List[int]:
    """Indices of the createAlignments calls that rank `rank` of `world` runs (round robin, call order preserved)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside the world")
    return list(range(rank, n_calls, world))


def ordered_gather(local: Dict[int, bytes], n_calls: int, dst: int = 0, group=None) -> Optional[List[bytes]]:
    """Host-side ordered gather: every rank contributes {call index: record stream}; rank `dst` receives the streams
    of all calls in call order (None elsewhere).  Works on any torch.distributed backend (gloo on CPU boxes, nccl on
    GPU boxes -- the payload travels as pickled host bytes, not through a data-path collective)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bucket = [None] * world if rank == dst else None
    dist.gather_object(local, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    merged: Dict[int, bytes] = {}
    for part in bucket:
        for k, v in part.items():
            if k in merged:
                raise RuntimeError(f"call {k} was produced by two ranks")
            merged[k] = v
    missing = [k for k in range(n_calls) if k not in merged]
    if missing:
        raise RuntimeError(f"calls {missing} were produced by no rank")
    return [merged[k] for k in range(n_calls)]
