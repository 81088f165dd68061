"""Collectives used by the search path — mirrors the helper names of reference src/dist_utils.py.

The reference needs 3 + 4*W small collectives per search (src/dist_utils.py:47-115 called from
src/index.py:128-142) and ships pickled passage text for all W*k candidates.  Here a search uses
one size exchange + one padded query all-gather (same contract as ``varsize_all_gather``) and ONE
all-gather of packed (score, id) candidates; passage text is resolved after the merge, for the k
winners only.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_rank() -> int:
    return dist.get_rank() if is_dist() else 0


def get_world_size() -> int:
    return dist.get_world_size() if is_dist() else 1


def barrier() -> None:
    if is_dist():
        dist.barrier()


@torch.no_grad()
def get_varsize(x: torch.Tensor, dim: int = 0) -> List[int]:
    """Per-rank sizes of ``x`` along ``dim`` (src/dist_utils.py:104-115), as a Python list."""
    if not is_dist():
        return [int(x.size(dim))]
    size = torch.tensor([x.size(dim)], device=x.device, dtype=torch.int64)
    allsizes = [torch.zeros_like(size) for _ in range(dist.get_world_size())]
    dist.all_gather(allsizes, size)
    return [int(s) for s in torch.cat(allsizes).tolist()]


@torch.no_grad()
def varsize_all_gather(x: torch.Tensor, sizes: Sequence[int] = None) -> torch.Tensor:
    """all_gather of tensors with different dim-0 sizes, concatenated (src/dist_utils.py:47-71).

    Padding rows are zero-filled (the reference leaves them uninitialised and discards them).
    """
    if not is_dist():
        return x
    if sizes is None:
        sizes = get_varsize(x)
    max_size = max(sizes) if sizes else 0
    if max_size == 0:
        return x.new_zeros((0,) + tuple(x.shape[1:]))
    if x.size(0) != max_size:
        pad = x.new_zeros((max_size - x.size(0),) + tuple(x.shape[1:]))
        x = torch.cat((x, pad), dim=0)
    x = x.contiguous()
    out = [torch.empty_like(x) for _ in sizes]
    dist.all_gather(out, x)
    return torch.cat([t[:n] for t, n in zip(out, sizes)], dim=0)


@torch.no_grad()
def all_gather_candidates(scores: torch.Tensor, ids: torch.Tensor):
    """ONE all-gather of the packed per-rank result: scores fp32 [B,k] + ids int64 [B,k] ->
    ([W,B,k] fp32, [W,B,k] int64).  12*B*k bytes per rank (77 KB at B=64, k=100)."""
    w = get_world_size()
    if w == 1:
        return scores.unsqueeze(0), ids.unsqueeze(0)
    b, k = scores.shape
    packed = torch.empty((3, b, k), dtype=torch.int32, device=scores.device)
    packed[0] = scores.contiguous().view(torch.int32)
    packed[1:] = ids.contiguous().view(torch.int32).view(b, k, 2).permute(2, 0, 1)
    out = [torch.empty_like(packed) for _ in range(w)]
    dist.all_gather(out, packed)
    allp = torch.stack(out, dim=0)  # [W, 3, B, k]
    g_scores = allp[:, 0].contiguous().view(torch.float32)
    g_ids = allp[:, 1:].permute(0, 2, 3, 1).contiguous().view(torch.int64).view(w, b, k)
    return g_scores, g_ids


def all_gather_object(obj):
    if not is_dist():
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out
