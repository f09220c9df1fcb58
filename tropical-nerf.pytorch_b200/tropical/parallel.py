"""Multi-GPU plumbing for the extraction path: one process per GPU over torch.distributed.

The path shards by OBJECT: every rank extracts the networks assigned to it and no
collective touches the data path.  What needs agreement across ranks is only
  * which objects a rank owns            -> `shard`
  * the time of the slowest rank         -> `max_over_ranks`
  * the catalogue of the extracted meshes -> `gather_catalogue`
All functions work with any backend (NCCL on the B200 box, gloo in the CPU tests).
"""
from typing import Any, Callable, Dict, List, Sequence

import torch
import torch.distributed as dist


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard(n_objects: int, world_size: int = None, r: int = None) -> List[int]:
    """Object numbers owned by rank r: contiguous blocks whose sizes differ by at most one."""
    world_size = world() if world_size is None else world_size
    r = rank() if r is None else r
    base, extra = divmod(n_objects, world_size)
    start = r * base + min(r, extra)
    return list(range(start, start + base + (1 if r < extra else 0)))


def max_over_ranks(value: float, device: str = None) -> float:
    """Largest value over all ranks (device times are reported as the max over ranks)."""
    if world() == 1:
        return float(value)
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_catalogue(local: Dict[int, Dict[str, Any]]) -> Dict[int, Dict[str, Any]]:
    """Every rank contributes {object number: summary}; every rank gets the union.  Raises if
    two ranks claim the same object."""
    if world() == 1:
        return dict(local)
    parts: List[Dict[int, Dict[str, Any]]] = [None] * world()
    dist.all_gather_object(parts, local)
    merged: Dict[int, Dict[str, Any]] = {}
    for p in parts:
        for k, v in p.items():
            if k in merged:
                raise RuntimeError(f"object {k} extracted by two ranks")
            merged[k] = v
    return merged


def extract_many(nets: Sequence[Any], extract: Callable[[Any], Dict[str, Any]]) -> Dict[int, Dict[str, Any]]:
    """Extract every network of `nets` once across the ranks; `extract(net)` returns the
    summary to publish (e.g. mesh sizes).  Returns the full catalogue on every rank."""
    mine = {i: extract(nets[i]) for i in shard(len(nets))}
    return gather_catalogue(mine)
