"""Tensor-level wrapper over the C ABI: torch owns every buffer, the extension borrows pointers.

``MipsEngine.search`` is the measured hot path (scan + per-shard merge, stream-ordered on torch's
current stream); ``B200Index`` (index.py) puts the reference's ``DistributedIndex`` API on top.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _native as N

_TORCH2MIPS = {torch.float16: N.MIPS_DTYPE_F16, torch.bfloat16: N.MIPS_DTYPE_BF16, torch.float32: N.MIPS_DTYPE_F32}


def _stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class MipsEngine:
    """Exact top-k inner-product search over one shard resident on one B200."""

    def __init__(self, dim: int, dtype: torch.dtype = torch.float16, device: Optional[torch.device] = None):
        if dtype not in (torch.float16, torch.bfloat16):
            raise ValueError(f"index dtype must be float16 or bfloat16, got {dtype}")
        self._lib = N.load()  # raises if the CUDA extension is missing: no fallback
        if not torch.cuda.is_available():
            raise RuntimeError("MipsEngine needs a CUDA (sm_100) device; there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dim = int(dim)
        self.dtype = dtype
        h = ctypes.c_void_p()
        N.check(self._lib.mips_create(ctypes.byref(h), self.device.index or 0, self.dim, _TORCH2MIPS[dtype]), None,
                "mips_create")
        self._h = h
        self._store = None
        self.max_k = self._lib.mips_max_k()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mips_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ index binding
    def bind(self, store: torch.Tensor, id_base: int = 0, id_stride: int = 1) -> None:
        """``store``: [n_local, dim] CUDA tensor.  Rows contiguous (K-major, row stride may exceed dim) or
        the transposed view of a reference-layout [dim, n_local] tensor (columns contiguous): both are
        consumed in place, the latter as an MN-major tensor-core operand."""
        if store.dim() != 2 or store.shape[1] != self.dim:
            raise ValueError(f"store must be [n, {self.dim}], got {tuple(store.shape)}")
        if store.dtype != self.dtype or store.device != self.device:
            raise ValueError(f"store must be {self.dtype} on {self.device}")
        n = int(store.shape[0])
        if store.stride(1) == 1 or n <= 1 and store.is_contiguous():
            layout, ld = 1, (store.stride(0) if n > 1 else self.dim)
        elif store.stride(0) == 1:
            layout, ld = 0, store.stride(1)           # [dim, n] storage seen through .t()
        else:
            raise ValueError(f"store needs unit stride along rows or columns, got strides {store.stride()}")
        N.check(self._lib.mips_bind_index_layout(self._h, ctypes.c_void_p(store.data_ptr()), n, ld, layout,
                                                 int(id_base), int(id_stride)), self._h, "mips_bind_index")
        self._store = store  # keep alive: the extension only borrows the pointer

    def pin_workspace(self, delta: int) -> None:
        """+1 while a captured CUDA graph holds pointers into the internal workspace, -1 when it is released: a
        pinned workspace that a larger search outgrows is kept alive (retired) instead of freed."""
        N.check(self._lib.mips_workspace_pin(self._h, int(delta)), self._h, "mips_workspace_pin")

    @property
    def n_local(self) -> int:
        return 0 if self._store is None else int(self._store.shape[0])

    # ------------------------------------------------------------------ search
    def search(self, queries: torch.Tensor, k: int, normalize: bool = False,
               out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """queries [B, dim] (fp32/fp16/bf16, CUDA) -> (scores fp32 [B, k] desc, global ids int64 [B, k])."""
        if queries.dim() != 2 or queries.shape[1] != self.dim:
            raise ValueError(f"queries must be [B, {self.dim}], got {tuple(queries.shape)}")
        if queries.dtype not in _TORCH2MIPS:
            queries = queries.float()
        if queries.device != self.device:
            queries = queries.to(self.device)
        if queries.stride(1) != 1:
            queries = queries.contiguous()
        b = int(queries.shape[0])
        if out is None:
            scores = torch.empty((b, k), dtype=torch.float32, device=self.device)
            ids = torch.empty((b, k), dtype=torch.int64, device=self.device)
        else:
            scores, ids = out
        q_ld = queries.stride(0) if b > 1 else self.dim
        rc = self._lib.mips_search_local(self._h, ctypes.c_void_p(queries.data_ptr()), _TORCH2MIPS[queries.dtype], q_ld,
                                         b, int(k), int(bool(normalize)), ctypes.c_void_p(scores.data_ptr()),
                                         ctypes.c_void_p(ids.data_ptr()), None, 0, _stream_ptr(self.device))
        N.check(rc, self._h, "mips_search_local")
        return scores, ids

    def search_host(self, host_queries: torch.Tensor, k: int, normalize: bool = False,
                    out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, wait: bool = True
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
        """End to end with HOST buffers (fp32 CPU tensor in, CPU tensors out): H2D + search + D2H + sync.
        ``wait=False`` only enqueues (mips_search_host_async): the outputs are valid after the current stream has
        completed and ``host_queries`` (pinned, contiguous fp32) must stay untouched until then."""
        if host_queries.device.type != "cpu" or host_queries.dtype != torch.float32 or not host_queries.is_contiguous():
            host_queries = host_queries.detach().to("cpu", torch.float32).contiguous()
        b = int(host_queries.shape[0])
        if out is None:
            scores = torch.empty((b, k), dtype=torch.float32).pin_memory()
            ids = torch.empty((b, k), dtype=torch.int64).pin_memory()
        else:
            scores, ids = out
        fn = self._lib.mips_search_host if wait else self._lib.mips_search_host_async
        self._async_keep = None if wait else host_queries      # a converted copy must outlive the enqueued H2D
        rc = fn(self._h, ctypes.cast(host_queries.data_ptr(), ctypes.POINTER(ctypes.c_float)), b,
                                        int(k), int(bool(normalize)),
                                        ctypes.cast(scores.data_ptr(), ctypes.POINTER(ctypes.c_float)),
                                        ctypes.cast(ids.data_ptr(), ctypes.POINTER(ctypes.c_int64)),
                                        _stream_ptr(self.device))
        N.check(rc, self._h, "mips_search_host")
        return scores, ids

    # ------------------------------------------------------------------ diagnostics
    STAT_NAMES = ["prod_wait", "mma_wait_full", "mma_wait_tmem", "epi_wait_tmem", "epi_select", "epi_compact",
                  "n_compact", "n_append", "total_cycles", "epi_ld"]

    def debug_config(self, flags: int = 0, collect_stats: bool = False):
        """flags: 1 = skip select, 2 = skip MMAs (results meaningless).  Returns the stats tensor
        [num_sms, 9] (uint64 as int64) the next scan launches accumulate into, or None."""
        stats = None
        if collect_stats:
            stats = torch.zeros((self._lib.mips_num_sms(self._h), self._lib.mips_debug_num_stats()), dtype=torch.int64,
                                device=self.device)
        self._dbg_stats = stats
        N.check(self._lib.mips_debug_config(self._h, int(flags),
                                            ctypes.c_void_p(stats.data_ptr()) if stats is not None else None),
                self._h, "mips_debug_config")
        return stats

    def scan_times_ms(self):
        """Durations (ms) of the full-shard scan launches recorded since the last call (needs flag 8)."""
        buf = (ctypes.c_float * 256)()
        n = ctypes.c_int(0)
        N.check(self._lib.mips_scan_times_ms(self._h, buf, 256, ctypes.byref(n)), self._h, "mips_scan_times_ms")
        return [float(buf[i]) for i in range(n.value)]

    def last_launch_count(self) -> int:
        return int(self._lib.mips_last_launch_count(self._h))

    # ------------------------------------------------------------------ merge / gather
    def merge(self, scores: torch.Tensor, ids: torch.Tensor, k_out: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """[L, B, k_in] candidate lists (each sorted desc) -> merged top k_out per query."""
        return merge_topk(scores, ids, k_out)

    def gather_rows(self, local_rows: torch.Tensor) -> torch.Tensor:
        rows = local_rows.reshape(-1).to(self.device, torch.int64).contiguous()
        out = torch.empty((rows.numel(), self.dim), dtype=self.dtype, device=self.device)
        N.check(self._lib.mips_gather_rows(self._h, ctypes.c_void_p(rows.data_ptr()), rows.numel(),
                                           ctypes.c_void_p(out.data_ptr()), _stream_ptr(self.device)), self._h,
                "mips_gather_rows")
        return out


def packed_result_buffer(batch: int, k: int, device, lists: int = 1):
    """One byte buffer per list laid out [fp32 scores [B,k] | pad to 8 B | int64 ids [B,k]] plus the two
    views into it.  A rank's search writes through the views; all ranks' buffers are exchanged with ONE
    all_gather and merged in place by ``merge_packed`` (no packing kernels)."""
    s_bytes = (batch * k * 4 + 7) // 8 * 8
    block = s_bytes + batch * k * 8
    buf = torch.empty((lists, block), dtype=torch.uint8, device=device)
    scores = buf[:, :batch * k * 4].view(torch.float32).view(lists, batch, k)
    ids = buf[:, s_bytes:].view(torch.int64).view(lists, batch, k)
    return buf, scores, ids


def merge_packed(buf: torch.Tensor, batch: int, k_in: int, k_out: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merges the [W, block] buffer produced by all-gathering ``packed_result_buffer`` blocks."""
    lib = N.load()
    if not buf.is_cuda:
        raise RuntimeError("merge_packed needs CUDA tensors; there is no CPU fallback")
    lists, block = buf.shape
    s_bytes = (batch * k_in * 4 + 7) // 8 * 8
    out_s = torch.empty((batch, k_out), dtype=torch.float32, device=buf.device)
    out_i = torch.empty((batch, k_out), dtype=torch.int64, device=buf.device)
    rc = lib.mips_merge_topk_strided(buf.device.index or 0, ctypes.c_void_p(buf.data_ptr()),
                                     ctypes.c_void_p(buf.data_ptr() + s_bytes), lists, block // 4, block // 8, batch, k_in,
                                     int(k_out), ctypes.c_void_p(out_s.data_ptr()), ctypes.c_void_p(out_i.data_ptr()),
                                     _stream_ptr(buf.device))
    N.check(rc, None, "mips_merge_topk_strided")
    return out_s, out_i


def merge_topk(scores: torch.Tensor, ids: torch.Tensor, k_out: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Device-side L-way merge of sorted (score, id) lists: replaces src/index.py:135-157."""
    lib = N.load()
    if scores.dim() != 3 or scores.shape != ids.shape:
        raise ValueError("scores/ids must both be [L, B, k_in]")
    if not scores.is_cuda:
        raise RuntimeError("merge_topk needs CUDA tensors; there is no CPU fallback")
    scores = scores.contiguous().float()
    ids = ids.contiguous().to(torch.int64)
    L, b, k_in = scores.shape
    out_s = torch.empty((b, k_out), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((b, k_out), dtype=torch.int64, device=scores.device)
    rc = lib.mips_merge_topk(scores.device.index or 0, ctypes.c_void_p(scores.data_ptr()), ctypes.c_void_p(ids.data_ptr()),
                             L, b, k_in, int(k_out), ctypes.c_void_p(out_s.data_ptr()), ctypes.c_void_p(out_i.data_ptr()),
                             _stream_ptr(scores.device))
    N.check(rc, None, "mips_merge_topk")
    return out_s, out_i


def rerank_topk(query_emb: torch.Tensor, cand_emb: torch.Tensor, k: int, want_rank: bool = False,
                want_emb: bool = True):
    """Fused re-rank of [B, L, D] candidates against [B, D] queries (one launch of mips_rerank).

    Returns (scores [B,k] fp32, positions [B,k] int64, ranks [B,L] int64 or None, emb [B,k,D] or None):
    the einsum + sort + slice + gather of reference src/rag.py:228-233."""
    lib = N.load()
    if query_emb.dim() != 2 or cand_emb.dim() != 3 or cand_emb.shape[0] != query_emb.shape[0] or \
            cand_emb.shape[2] != query_emb.shape[1]:
        raise ValueError("rerank_topk expects query_emb [B, D] and cand_emb [B, L, D]")
    if not (query_emb.is_cuda and cand_emb.is_cuda):
        raise RuntimeError("rerank_topk needs CUDA tensors; there is no CPU fallback")
    dt = cand_emb.dtype if cand_emb.dtype in _TORCH2MIPS else torch.float32
    cand = cand_emb.to(dt).contiguous()
    q = query_emb.to(dt)
    if q.stride(1) != 1:
        q = q.contiguous()
    b, L, d = cand.shape
    dev = cand.device
    out_s = torch.empty((b, k), dtype=torch.float32, device=dev)
    out_p = torch.empty((b, k), dtype=torch.int64, device=dev)
    out_r = torch.empty((b, L), dtype=torch.int64, device=dev) if want_rank else None
    out_e = torch.empty((b, k, d), dtype=dt, device=dev) if want_emb else None
    ptr = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None and t.numel() else 0)  # noqa: E731
    rc = lib.mips_rerank(dev.index or 0, ptr(q), q.stride(0) if b else d, ptr(cand), _TORCH2MIPS[dt], b, L, d, int(k),
                         ptr(out_s), ptr(out_p), ptr(out_r), ptr(out_e), _stream_ptr(dev))
    N.check(rc, None, "mips_rerank")
    return out_s, out_p, out_r, out_e
