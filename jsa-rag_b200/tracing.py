"""Tracing hooks of the search path: NVTX ranges and CUDA-synchronised timers.

The reference's only instrumentation is a wall-clock delta around the search stuffed into ``iter_stats``
(``iter_stats["runtime/search"] = (time.time() - search_start, 1)``, src/rag.py:156,170) — not CUDA-synchronised
except by accident (``.tolist()``).  Here:

* ``nvtx_range(name)`` — context manager; ranges show up in Nsight timelines (``mips.query_gather``,
  ``mips.local_search``, ``mips.exchange_merge``, ``mips.resolve_passages``; the C library adds
  ``mips.prep`` / ``mips.scan`` / ``mips.select`` per launch when ``JSA_MIPS_NVTX=1``).  A no-op without CUDA.
* ``SearchTimer`` — CUDA events on the search's stream around the device part plus a wall clock around the whole
  call; ``B200Index.search_knn`` fills ``index.last_search_stats`` and, when ``index.iter_stats`` is a dict, the
  reference's ``runtime/search`` key (same ``(seconds, count)`` tuples as src/util.py's WeightedAvgStats consumes).
"""
from __future__ import annotations

import contextlib
import os
import time

import torch

_NVTX = os.environ.get("JSA_MIPS_NVTX", "1") != "0"


@contextlib.contextmanager
def nvtx_range(name: str):
    on = _NVTX and torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


class SearchTimer:
    """with SearchTimer(device) as t: ... t.device_done() ... ; t.stats() after the block."""

    def __init__(self, device=None):
        self.cuda = device is not None and torch.device(device).type == "cuda"
        self.device = device
        self._ev0 = self._ev1 = None
        self.t0 = self.t1 = self.t_host = None

    def __enter__(self):
        self.t0 = time.perf_counter()
        if self.cuda:
            self._ev0 = torch.cuda.Event(enable_timing=True)
            self._ev1 = torch.cuda.Event(enable_timing=True)
            self._ev0.record(torch.cuda.current_stream(self.device))
        return self

    def device_done(self):
        """Call right after the last kernel of the search has been enqueued."""
        if self.cuda:
            self._ev1.record(torch.cuda.current_stream(self.device))
        self.t_host = time.perf_counter()

    def __exit__(self, *exc):
        if self.cuda:
            torch.cuda.current_stream(self.device).synchronize()    # a CUDA-synchronised runtime/search
        self.t1 = time.perf_counter()
        return False

    def stats(self) -> dict:
        total = self.t1 - self.t0
        out = {"runtime/search": total}
        if self.cuda and self._ev1 is not None and self.t_host is not None:
            dev = self._ev0.elapsed_time(self._ev1) * 1e-3
            out["runtime/search_device"] = dev
            out["runtime/search_host_tail"] = max(0.0, total - dev)
        return out
