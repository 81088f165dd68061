"""Node-shared passage store: global id -> passage dict, resolved locally on every rank.

The reference moves passage TEXT through NCCL on every search: each rank pickles the dicts of all its W*k candidates
and 2*W gathers ship them (src/index.py:34-41,137-142).  Here the exchange carries (score, global id) only, and the
winners' text is looked up after the merge.  A rank holds its own shard's dicts in memory (``doc_map``, as in the
reference); winners that live on OTHER ranks are read from this store: every rank serialises its shard once into
``passages.<rank>.bin`` + an int64 offset table under a node-local directory (/dev/shm when available), all ranks map
all W files read-only, and the files are unlinked as soon as they are mapped (the page cache keeps one copy per
node; nothing is left behind, also after a crash).  A lookup is ``pickle.loads`` of one record: no collective, no
device sync, no dependence on the other ranks being in the same call.
"""
from __future__ import annotations

import mmap
import os
import pickle
import shutil
import socket
import tempfile
from typing import List, Optional, Sequence

import numpy as np

from . import dist_utils


def _shared_dir_root() -> str:
    for cand in (os.environ.get("JSA_MIPS_PASSAGE_DIR"), "/dev/shm", tempfile.gettempdir()):
        if cand and os.path.isdir(cand) and os.access(cand, os.W_OK):
            return cand
    return tempfile.gettempdir()


class PassageStore:
    """W memory-mapped shards of pickled passage records; ``get(owner, local_row)`` -> dict."""

    def __init__(self, maps: List[Optional[mmap.mmap]], offsets: List[np.ndarray]):
        self._maps = maps
        self._views = [memoryview(m) if m is not None else None for m in maps]
        self._offsets = offsets
        # decoded records of recently returned remote passages (popular passages recur from search to search):
        # bounded, cleared wholesale when full
        self._cache = {}
        self._cache_cap = int(os.environ.get("JSA_MIPS_PASSAGE_CACHE", 1 << 18))

    # ------------------------------------------------------------------ construction
    @staticmethod
    def serialise(passages: Sequence, n: Optional[int] = None):
        """(blob, offsets[n + 1]) of the records ``pickle.dumps(passages[i])``, i in [0, n)."""
        n = len(passages) if n is None else n
        offs = np.zeros(n + 1, dtype=np.int64)
        parts = []
        at = 0
        dumps = pickle.dumps
        for i in range(n):
            rec = dumps(passages[i], protocol=pickle.HIGHEST_PROTOCOL)
            parts.append(rec)
            at += len(rec)
            offs[i + 1] = at
        return b"".join(parts), offs

    @classmethod
    def build_shared(cls, local_passages: Sequence, n_local: int) -> Optional["PassageStore"]:
        """Collective.  Every rank contributes its shard; returns None when the ranks do not share a host (the caller
        then keeps exchanging the winners' text through the process group)."""
        w, r = dist_utils.get_world_size(), dist_utils.get_rank()
        hosts = dist_utils.all_gather_object(socket.gethostname())
        if len(set(hosts)) != 1:
            return None
        root = None
        if r == 0:
            root = tempfile.mkdtemp(prefix="jsa_mips_passages_", dir=_shared_dir_root())
        root = dist_utils.all_gather_object(root)[0]
        ok = True
        try:
            # cheap feasibility check before serialising millions of records: the node-local directory must hold
            # every rank's shard (sizes estimated from the first records of this shard)
            probe = min(n_local, 512)
            if probe:
                avg = sum(len(pickle.dumps(local_passages[i], protocol=pickle.HIGHEST_PROTOCOL)) for i in range(probe)) / probe
                need = int((avg + 8) * n_local * 1.1) * w
                if shutil.disk_usage(root).free < need:
                    raise OSError(f"{root}: {need} bytes needed for the shared passage store")
            blob, offs = cls.serialise(local_passages, n_local)
            with open(os.path.join(root, f"passages.{r}.bin"), "wb") as f:
                f.write(blob if blob else b"\0")        # an empty file cannot be mapped
            np.save(os.path.join(root, f"offsets.{r}.npy"), offs)
            del blob
        except Exception:  # noqa: BLE001 - e.g. /dev/shm too small: fall back, do not fail the search
            ok = False
        ok = all(dist_utils.all_gather_object(ok))      # also the barrier: every shard is on disk
        store = None
        if ok:
            maps, offsets = [], []
            for src in range(w):
                with open(os.path.join(root, f"passages.{src}.bin"), "rb") as f:
                    maps.append(mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ))
                offsets.append(np.load(os.path.join(root, f"offsets.{src}.npy")))
            store = cls(maps, offsets)
        dist_utils.barrier()                             # everyone has mapped: the names can go
        if r == 0:
            shutil.rmtree(root, ignore_errors=True)
        return store

    @classmethod
    def from_shards(cls, shards: Sequence[Sequence]) -> "PassageStore":
        """Single-process construction (tests, tools): shard s holds the passages of rank s."""
        maps, offsets = [], []
        for sh in shards:
            blob, offs = cls.serialise(sh)
            m = mmap.mmap(-1, max(1, len(blob)))
            m.write(blob)
            maps.append(m)
            offsets.append(offs)
        return cls(maps, offsets)

    # ------------------------------------------------------------------ lookups
    def shard_len(self, owner: int) -> int:
        return int(self._offsets[owner].shape[0]) - 1

    def get(self, owner: int, local_row: int) -> dict:
        offs = self._offsets[owner]
        return pickle.loads(self._views[owner][offs[local_row]:offs[local_row + 1]])

    def get_many(self, owners: np.ndarray, local_rows: np.ndarray) -> list:
        """Flat arrays -> list of passage dicts (same order)."""
        out = []
        loads, cache, cap = pickle.loads, self._cache, self._cache_cap
        if len(cache) >= cap:
            cache.clear()
        for o, l in zip(owners.tolist(), local_rows.tolist()):
            key = (o << 40) | l
            doc = cache.get(key)
            if doc is None:
                offs = self._offsets[o]
                doc = loads(self._views[o][offs[l]:offs[l + 1]])
                if cap:
                    cache[key] = doc
            out.append(doc)
        return out

    def close(self) -> None:
        for v in self._views:
            if v is not None:
                v.release()
        for m in self._maps:
            if m is not None:
                m.close()
        self._maps, self._views, self._offsets = [], [], []
        self._cache = {}
