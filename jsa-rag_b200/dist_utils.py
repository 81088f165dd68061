"""Collectives used by the search path — mirrors the helper names of reference src/dist_utils.py.

The reference needs 3 + 4*W small collectives per search (src/dist_utils.py:47-115 called from
src/index.py:128-142) and ships pickled passage text for all W*k candidates.  Here a search uses
one size exchange + one padded query all-gather (same contract as ``varsize_all_gather``) and ONE
all-gather of packed (score, id) candidates; passage text is resolved after the merge, for the k
winners only.  On one node ``B200Index.search`` replaces both tensor all-gathers with NVLink peer
stores (``exchange.py``); the functions here remain the multi-node / gloo path.
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
    out = torch.empty((len(sizes),) + tuple(x.shape), dtype=x.dtype, device=x.device)
    _all_gather_into(out, x)
    if all(n == max_size for n in sizes):
        return out.view((len(sizes) * max_size,) + tuple(x.shape[1:]))
    return torch.cat([out[r, :n] for r, n in enumerate(sizes)], dim=0)


def _all_gather_into(out: torch.Tensor, x: torch.Tensor) -> None:
    """dist.all_gather_into_tensor where the backend has it (NCCL), list all_gather otherwise (gloo)."""
    try:
        dist.all_gather_into_tensor(out, x)
    except (RuntimeError, NotImplementedError):
        parts = list(out.view(get_world_size(), *x.shape).unbind(0))
        dist.all_gather(parts, x)


@torch.no_grad()
def all_gather_candidates(scores: torch.Tensor, ids: torch.Tensor):
    """The one exchange step of the search: every rank contributes its local top-k
    (fp32 scores [B,k], int64 global ids [B,k]; 12*B*k bytes per rank — 77 KB at B=64, k=100) and
    receives all of them as ([W,B,k] fp32, [W,B,k] int64), laid out for the device merge: two
    all_gather_into_tensor calls back to back (no packing kernels, no host sync)."""
    w = get_world_size()
    if w == 1:
        return scores.unsqueeze(0), ids.unsqueeze(0)
    b, k = scores.shape
    scores, ids = scores.contiguous(), ids.contiguous()
    g_scores = torch.empty((w, b, k), dtype=scores.dtype, device=scores.device)
    g_ids = torch.empty((w, b, k), dtype=ids.dtype, device=ids.device)
    _all_gather_into(g_scores, scores)
    _all_gather_into(g_ids, ids)
    return g_scores, g_ids


def all_gather_object(obj):
    if not is_dist():
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def all_to_all_objects(objs, device=None):
    """objs[d] is delivered to rank d; returns [object from rank 0, ..., object from rank W-1].

    NCCL: two all_to_all_single calls (sizes, then the pickled bytes) so that every rank receives only
    what is meant for it.  Other backends (gloo on CPU): one object all-gather, then pick."""
    if not is_dist():
        return list(objs)
    w, r = dist.get_world_size(), dist.get_rank()
    if dist.get_backend() != "nccl" or device is None or torch.device(device).type != "cuda":
        gathered = all_gather_object(list(objs))
        return [gathered[src][r] for src in range(w)]
    import pickle
    blobs = [pickle.dumps(o, protocol=pickle.HIGHEST_PROTOCOL) for o in objs]
    in_sizes = torch.tensor([len(b) for b in blobs], dtype=torch.int64, device=device)
    out_sizes = torch.empty_like(in_sizes)
    dist.all_to_all_single(out_sizes, in_sizes)
    out_list = out_sizes.tolist()
    send = torch.frombuffer(bytearray(b"".join(blobs)), dtype=torch.uint8).to(device)
    recv = torch.empty(int(sum(out_list)), dtype=torch.uint8, device=device)
    dist.all_to_all_single(recv, send, out_list, [len(b) for b in blobs])
    raw = recv.cpu().numpy().tobytes()
    out, at = [], 0
    for n in out_list:
        out.append(pickle.loads(raw[at:at + n]))
        at += n
    return out
