"""Tensor helpers the reference installs as `torch.ext` (tropical/torch_ext.py).

Host-side utilities (not on the device hot path); vectorised instead of the reference's
per-element Python loops, same results.
"""
import torch
import torch.nn.functional as F
from torch import Tensor


def _first_last(t: Tensor, last: bool) -> Tensor:
    assert 2 == len(t.shape)
    nz = t != 0
    rows = nz.any(dim=1).nonzero()[:, 0]
    cols = torch.arange(t.shape[1], device=t.device).expand_as(nz)
    if last:
        col = torch.where(nz, cols, torch.full_like(cols, -1)).max(dim=1)[0]
    else:
        col = torch.where(nz, cols, torch.full_like(cols, t.shape[1])).min(dim=1)[0]
    return torch.stack([rows, col[rows]], dim=1).long()


def nonzero_last(t: Tensor) -> Tensor:
    """(row, column of the last nonzero) for every row that has one (torch_ext.py:18-29)."""
    return _first_last(t, True)


def nonzero_first(t: Tensor) -> Tensor:
    """(row, column of the first nonzero) for every row that has one (torch_ext.py:32-43)."""
    return _first_last(t, False)


def batched_index_select(t, dim, inds):
    """torch_ext.py:47-50."""
    dummy = inds.unsqueeze(2).expand(inds.size(0), inds.size(1), t.size(2))
    return t.gather(dim, dummy)


def batched_unique_consecutive(t, null_value=-1):
    """torch_ext.py:54-66."""
    rows = [torch.unique_consecutive(row) for row in t]
    width = max(len(row) for row in rows)
    return torch.stack([F.pad(row, (0, width - len(row)), value=null_value) for row in rows])
