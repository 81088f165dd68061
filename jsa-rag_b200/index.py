"""B200Index — drop-in for the reference's ``DistributedIndex`` (src/index.py:44-161).

Same public surface (``init_embeddings``, ``.embeddings``, ``.doc_map``, ``search_knn``,
``save_index``, ``load_index``, ``is_index_trained``), same argument meaning, return order
``(docs, scores)`` and error behaviour.  What is different underneath:

* storage is K-major ``[n_local, dim]`` (one contiguous 1536-byte row per passage — what TMA and
  tcgen05 want); ``.embeddings`` is the transposed *view* ``[dim, n_local]``, so the reference's
  write site ``index.embeddings[:, a:b] = emb.T`` (src/rag.py:120) writes straight through and
  ``torch.save(self.embeddings[:, a:b].clone())`` still produces the reference's shard bytes;
* ``torch.matmul`` + ``torch.topk`` (src/index.py:118-119) is one fused sm_100a kernel + a merge;
* the cross-rank merge (src/index.py:135-157: 2*W gathers of scores and pickled passage text) is
  ONE all-gather of (score, global id) + a device merge; text is resolved for the k winners only.

There is no CPU search path: ``search_knn`` raises if the CUDA extension is missing.
"""
from __future__ import annotations

import math
import os
import pickle
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import dist_utils
from .tracing import SearchTimer, nvtx_range

EMBEDDINGS_DIM: int = 768  # reference src/retrievers.py:14


_LOAD_CHUNK = 1 << 20


def _load_shard(path: str) -> torch.Tensor:
    """A saved [D, n_s] shard (src/index.py:80 writes it with torch.save) as a host tensor; memory-mapped when the
    file is in the zip format, so that load_index never holds a whole shard set in host memory."""
    try:
        return torch.load(path, map_location="cpu", mmap=True)
    except Exception:
        return torch.load(path, map_location="cpu")


class B200Index(object):
    def __init__(self, dtype: torch.dtype = torch.float16, device: Optional[str] = None, layout: str = "nd"):
        """``layout``: physical storage of the matrix.  "nd" (default) = one contiguous row per passage
        (K-major; streams at full HBM bandwidth).  "dn" = the reference's own contiguous [dim, n_local]
        layout (src/index.py:52), searched in place as an MN-major tensor-core operand (~0.6x the
        bandwidth: a passage tile then touches dim separate DRAM pages); n_local should be a multiple of 8."""
        if layout not in ("nd", "dn"):
            raise ValueError("layout must be 'nd' or 'dn'")
        self.layout = layout
        self._store: Optional[torch.Tensor] = None  # [n_local, dim] (a transposed view when layout == "dn")
        self.doc_map = dict()
        self.is_in_gpu = True  # reference attribute (src/index.py:48); False keeps storage on the host (I/O only)
        self.dtype = dtype
        self._device = device
        self._engine = None
        self._bound_key = None
        self._any_rank_has_queries = False
        self._last_all = None
        self._all_counts = [0]
        # global id of local row r = id_base + r * id_stride  (see _set_sharding)
        self._id_base, self._id_stride = 0, 1
        self._sharding = "round_robin"
        self._sharding_epoch = 0    # bumped by every (collective) _set_sharding: invalidates the shared passage store
        self.last_passage_path = None
        self.round_scores_to_index_dtype = True  # reference returns fp16-rounded scores (src/index.py:118,153)
        # True: every rank always passes the same number of queries (fixed per-GPU batch, as in the
        # reference's training loop; evaluate.py:49-54 pads iterators to keep ranks in step) -> the size
        # exchange (a collective + a host sync, src/index.py:129-130) is skipped
        self.equal_batch = False
        # search_knn leaves its CUDA-synchronised timings here ("runtime/search", ".../search_device",
        # ".../search_host_tail", seconds); set ``iter_stats`` to the trainer's dict and the reference's
        # ``runtime/search`` entry (src/rag.py:170) is filled with the same (value, count) tuples
        self.last_search_stats = {}
        self.iter_stats = None

    # ------------------------------------------------------------------ storage
    def _storage_device(self) -> torch.device:
        if self._device is not None:
            return torch.device(self._device)
        if self.is_in_gpu and torch.cuda.is_available():
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    @property
    def embeddings(self) -> Optional[torch.Tensor]:
        """[dim, n_local] view of the K-major storage (reference layout, src/index.py:52)."""
        return None if self._store is None else self._store.t()

    @embeddings.setter
    def embeddings(self, value: Optional[torch.Tensor]) -> None:
        if value is None:
            self._store = None
            return
        if value.dim() != 2:
            raise ValueError("embeddings must be [dim, n]")
        value = value.to(device=self._storage_device(), dtype=self.dtype)
        self._store = value.contiguous().t() if self.layout == "dn" else value.t().contiguous()
        # shard sizes and the global-id mapping follow the new matrix.  With more than one rank this is collective
        # (every rank assigns its own shard, like init_embeddings / load_index).
        self._set_sharding(self._sharding)

    def init_embeddings(self, passages, dim: Optional[int] = EMBEDDINGS_DIM):
        """src/index.py:50-54 — allocates zeroed storage; passages were round-robin sharded by
        load_passages (src/index_io.py:41), hence global id = local * W + rank."""
        self.doc_map = {i: doc for i, doc in enumerate(passages)}
        self._store = self._alloc(len(passages), dim, zero=True)
        self._set_sharding("round_robin")

    def _alloc(self, n: int, dim: int, zero: bool = False) -> torch.Tensor:
        make = torch.zeros if zero else torch.empty
        if self.layout == "dn":
            return make(dim, n, dtype=self.dtype, device=self._storage_device()).t()
        return make(n, dim, dtype=self.dtype, device=self._storage_device())

    def _set_sharding(self, mode: str) -> None:
        self._sharding = mode
        self._sharding_epoch += 1
        w, r = dist_utils.get_world_size(), dist_utils.get_rank()
        n = 0 if self._store is None else int(self._store.shape[0])
        counts = dist_utils.all_gather_object(n) if w > 1 else [n]     # every rank knows every shard size
        self._all_counts = counts
        if mode == "round_robin":
            self._id_base, self._id_stride = r, w
        else:  # contiguous: rank r owns rows [offset_r, offset_r + n_r)
            self._id_base, self._id_stride = int(sum(counts[:r])), 1
        self._bound_key = None

    def is_index_trained(self) -> bool:
        return True  # src/index.py:160-161

    def train_index_bychunks(self) -> None:  # never reached for a flat index (src/rag.py:127-130)
        return None

    # ------------------------------------------------------------------ shard I/O (src/index.py:56-112)
    def _get_saved_embedding_path(self, save_dir: str, shard: int) -> str:
        return os.path.join(save_dir, f"embeddings.{shard}.pt")

    def _get_saved_passages_path(self, save_dir: str, shard: int) -> str:
        return os.path.join(save_dir, f"passages.{shard}.pt")

    def save_index(self, path: str, total_saved_shards: int, overwrite_saved_passages: bool = False) -> None:
        """Writes the reference's shard files: ``embeddings.{s}.pt`` = torch.save of a contiguous
        fp16 [dim, n_s] tensor, ``passages.{s}.pt`` = raw pickle of the passage list (src/index.py:62-88)."""
        assert self._store is not None
        rank = dist_utils.get_rank()
        ws = dist_utils.get_world_size()
        assert total_saved_shards % ws == 0, f"N workers must be a multiple of shards to save"
        shards_per_worker = total_saved_shards // ws
        n_embeddings = self._store.shape[0]
        embeddings_per_shard = math.ceil(n_embeddings / shards_per_worker)
        assert n_embeddings == len(self.doc_map), len(self.doc_map)
        for shard_ind, shard_start in enumerate(range(0, n_embeddings, embeddings_per_shard)):
            shard_end = min(shard_start + embeddings_per_shard, n_embeddings)
            shard_id = shard_ind + rank * shards_per_worker
            passage_shard_path = self._get_saved_passages_path(path, shard_id)
            if not os.path.exists(passage_shard_path) or overwrite_saved_passages:
                passage_shard = [self.doc_map[i] for i in range(shard_start, shard_end)]
                with open(passage_shard_path, "wb") as fobj:
                    pickle.dump(passage_shard, fobj, protocol=pickle.HIGHEST_PROTOCOL)
            # [dim, n_s] contiguous, exactly what `self.embeddings[:, a:b].clone()` holds in the reference
            embeddings_shard = self._store[shard_start:shard_end].t().contiguous()
            torch.save(embeddings_shard, self._get_saved_embedding_path(path, shard_id))

    def load_index(self, path: str, total_saved_shards: int):
        """Loads the shard files of this rank (src/index.py:90-112).  Shards are copied one by one
        into preallocated K-major storage (no 2x concat peak)."""
        rank = dist_utils.get_rank()
        ws = dist_utils.get_world_size()
        assert total_saved_shards % ws == 0, f"N workers must be a multiple of shards to save"
        shards_per_worker = total_saved_shards // ws
        passages, shards = [], []
        for shard_id in range(rank * shards_per_worker, (rank + 1) * shards_per_worker):
            with open(self._get_saved_passages_path(path, shard_id), "rb") as fobj:
                passages.append(pickle.load(fobj))
            shards.append(_load_shard(self._get_saved_embedding_path(path, shard_id)))
        self.doc_map = {}
        n_passages = 0
        for chunk in passages:
            for p in chunk:
                self.doc_map[n_passages] = p
                n_passages += 1
        dim = shards[0].shape[0] if shards else EMBEDDINGS_DIM
        n = sum(int(s.shape[1]) for s in shards)
        self._store = self._alloc(n, dim)
        at = 0
        for s in shards:
            n_s = int(s.shape[1])
            if self._store.is_cuda:
                # stream: <= 1M columns at a time host -> device as stored ([D, cols]), transposed into the K-major
                # rows on the device; host peak = one chunk (the shard itself is memory-mapped), device peak = + one chunk
                for c0 in range(0, n_s, _LOAD_CHUNK):
                    c1 = min(n_s, c0 + _LOAD_CHUNK)
                    blk = s[:, c0:c1].to(self._store.device)
                    self._store[at + c0:at + c1].copy_(blk.t())
            else:
                self._store[at:at + n_s].copy_(s.t())
            at += n_s
        self._set_sharding("contiguous")

    # ------------------------------------------------------------------ native search
    def _get_engine(self):
        if self._store is None:
            raise RuntimeError("index has no embeddings: call init_embeddings or load_index first")
        if not self._store.is_cuda:
            raise RuntimeError("B200Index.search_knn needs the index on a B200 (is_in_gpu=True and a CUDA device); "
                               "there is no CPU fallback")
        from .engine import MipsEngine
        if self._engine is None or self._engine.dim != self._store.shape[1] or self._engine.dtype != self._store.dtype \
                or self._engine.device != self._store.device:
            self._engine = MipsEngine(int(self._store.shape[1]), self._store.dtype, self._store.device)
            self._bound_key = None
        key = (self._store.data_ptr(), tuple(self._store.shape), self._store.stride(0), self._id_base, self._id_stride)
        if key != self._bound_key:
            self._engine.bind(self._store, self._id_base, self._id_stride)
            self._bound_key = key
        return self._engine

    def _local_search(self, allqueries: torch.Tensor, topk: int, normalize: bool = False
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Fused score + select over this rank's shard -> (fp32 scores [B,k], global ids [B,k])."""
        n_local = 0 if self._store is None else int(self._store.shape[0])
        if topk > n_local:
            raise RuntimeError("selected index k out of range")  # torch.topk's message, src/index.py:119
        return self._get_engine().search(allqueries, topk, normalize=normalize)

    def _merge_lists(self, scores: torch.Tensor, ids: torch.Tensor, topk: int):
        from .engine import merge_topk
        return merge_topk(scores, ids, topk)

    @torch.no_grad()
    def search(self, queries: torch.Tensor, topk: int, normalize: bool = False, replicated: bool = False
               ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Tensor-native distributed search: this rank's queries -> (scores fp32 [b,k], global ids [b,k]).

        All ranks must call it together (like the reference's search_knn).  Steps: all-gather
        queries -> local fused search -> exchange of candidates -> device merge -> own rows.  On one node
        both exchanges are NVLink peer stores (exchange.py); otherwise NCCL all-gathers.

        ``replicated=True``: every rank passes the SAME [B, dim] queries (a front end that broadcasts a request to
        the shards, SURVEY §8e "in the server/benchmark the same [B, D] is broadcast"): there is no query exchange
        and every rank receives the merged result of all B queries.
        """
        w, r = dist_utils.get_world_size(), dist_utils.get_rank()
        self._any_rank_has_queries = False
        if self._store is not None and self._store.is_cuda and queries.device != self._store.device:
            queries = queries.to(self._store.device)
        if w == 1:
            if queries.shape[0] == 0:
                dev = queries.device
                return torch.empty(0, topk, device=dev), torch.empty(0, topk, dtype=torch.int64, device=dev)
            return self._local_search(queries, topk, normalize)
        if topk > min(self._all_counts):
            # the reference fails inside torch.topk on the rank whose shard is too small (and leaves the
            # others waiting in a collective); here every rank raises the same error before communicating
            raise RuntimeError("selected index k out of range")
        if replicated:
            sizes = [int(queries.shape[0])] + [0] * (w - 1)      # one logical owner; every rank returns all rows
            allqueries = queries
        else:
            sizes = [int(queries.shape[0])] * w if self.equal_batch else dist_utils.get_varsize(queries)   # src/index.py:129
            allqueries = None
        nbytes = queries.numel() * queries.element_size()
        if allqueries is None and queries.is_cuda and sizes[0] > 0 and len(set(sizes)) == 1 and nbytes % 8 == 0 and queries.dim() == 2 \
                and torch.distributed.get_backend() == "nccl":
            xq = self._peer_exchange(nbytes, queries.device, "_xchg_q")
            if xq is not None:                                                     # NVLink peer stores, no collective call
                allqueries = xq.gather(queries).view(w * sizes[0], queries.shape[1])
        if allqueries is None:
            with nvtx_range("mips.query_gather"):
                allqueries = dist_utils.varsize_all_gather(queries, sizes)         # src/index.py:128
        offs = np.cumsum([0] + sizes)
        if allqueries.shape[0] == 0:
            return (torch.empty(0, topk, device=queries.device),
                    torch.empty(0, topk, dtype=torch.int64, device=queries.device))
        self._any_rank_has_queries = True
        if allqueries.is_cuda and type(self)._local_search is B200Index._local_search:
            # device path: the local result is written straight into this rank's block of the exchange
            # buffer; ONE all-gather moves 12*B*k bytes per rank; the merge reads the gathered blocks
            from .engine import merge_packed, packed_result_buffer
            bt = int(allqueries.shape[0])
            buf, ls, li = packed_result_buffer(bt, topk, allqueries.device)
            n_local = 0 if self._store is None else int(self._store.shape[0])
            if topk > n_local:
                raise RuntimeError("selected index k out of range")
            with nvtx_range("mips.local_search"):
                self._get_engine().search(allqueries, topk, normalize=normalize, out=(ls[0], li[0]))   # src/index.py:132
            xchg = self._peer_exchange(int(buf.shape[1]), buf.device)
            with nvtx_range("mips.exchange_merge"):
                if xchg is not None:
                    # NVLink peer stores into every rank's slot + wait-and-merge kernel (replaces :135-157)
                    ms, mi = xchg.merge(buf[0], bt, topk, topk)
                else:
                    gathered = torch.empty((w, buf.shape[1]), dtype=torch.uint8, device=buf.device)
                    torch.distributed.all_gather_into_tensor(gathered, buf[0])                     # replaces :139-142
                    ms, mi = merge_packed(gathered, bt, topk, topk)                                  # replaces :143-157
        else:
            ls, li = self._local_search(allqueries, topk, normalize)               # src/index.py:132
            gs, gi = dist_utils.all_gather_candidates(ls, li)                      # replaces :139-142
            ms, mi = self._merge_lists(gs, gi, topk)                               # replaces :143-157
        self._last_all = (mi, offs)
        if replicated:
            return ms, mi
        sl = slice(int(offs[r]), int(offs[r + 1]))
        return ms[sl], mi[sl]

    def _peer_exchange(self, block_bytes: int, device, which: str = "_xchg"):
        """The exchange object for blocks of ``block_bytes`` — ``_xchg`` for candidates, ``_xchg_q`` for queries
        (created or grown collectively: block sizes derive from the global batch and k, which are the same on
        every rank).  None when the ranks cannot map each other's memory: the caller then uses NCCL."""
        x = getattr(self, which, None)
        if x is False or getattr(self, "_p2p_off", False):
            return None
        if x is not None and x.capacity >= block_bytes:
            return x
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("the peer exchange must be sized before graph capture (run one eager search first)")
        from .exchange import make_peer_exchange
        if x is not None:
            if getattr(self, "_live_graphs", 0) > 0:
                # a captured search still stores into / reads from this exchange on replay (on every rank): keep it
                # mapped until the last graph is released, and serve the larger request from a new one
                self._retired_xchg = getattr(self, "_retired_xchg", []) + [x]
            else:
                x.close()
        x = make_peer_exchange(device, max(block_bytes, getattr(self, "_xchg_min_bytes", 1 << 20)))
        if x is None:
            self._p2p_off = True
        setattr(self, which, x)
        return x

    def close_exchange(self):
        """Collective.  Releases the peer-mapped exchange buffers (call before destroy_process_group())."""
        for which in ("_xchg", "_xchg_q"):
            x = getattr(self, which, None)
            if x:
                x.close()
            setattr(self, which, None)
        if getattr(self, "_pstore", None) is not None:
            self._pstore.close()
            self._pstore, self._store_epoch = None, None

    def make_graphed_search(self, batch: int, topk: int, normalize: bool = False, query_dtype=torch.float32,
                            replicated: bool = False):
        """Captures one search for a fixed per-rank batch into a CUDA graph — on several ranks the whole
        distributed flow (query all-gather, fused scan + select, candidate all-gather, merge; NCCL
        collectives are captured too).  Returns ``run(queries) -> (scores, ids)`` that copies the queries
        into the captured input and replays the graph: a few microseconds of host time per search
        instead of ~40 us (1 rank) / ~150 us (N ranks), which matters when a search lasts ~1 ms.
        All ranks must call this together; ``equal_batch`` is implied.  Call ``run.release()`` (or drop
        every reference to ``run``) BEFORE ``destroy_process_group()``: tearing the NCCL communicator
        down while a graph that captured its kernels is alive hangs."""
        if self._store is None or not self._store.is_cuda:
            raise RuntimeError("make_graphed_search needs the index on a CUDA device; there is no CPU fallback")
        dev = self._store.device
        prev_equal = self.equal_batch
        self.equal_batch = True
        # The graph bakes in raw pointers into the engine's workspace and the exchange buffers.  From here until
        # release() neither may be freed: the engine retires (instead of freeing) a workspace it outgrows, and
        # _peer_exchange keeps an outgrown exchange mapped (a larger eager search in between is therefore safe).
        engine = self._get_engine()
        engine.pin_workspace(+1)
        self._live_graphs = getattr(self, "_live_graphs", 0) + 1
        static_q = torch.zeros(batch, int(self._store.shape[1]), dtype=query_dtype, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):                      # warm-up: workspace growth, NCCL channels, lazy inits
                self.search(static_q, topk, normalize, replicated=replicated)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out_s, out_i = self.search(static_q, topk, normalize, replicated=replicated)
        self.equal_batch = prev_equal

        state = {"graph": graph, "out": (out_s, out_i)}

        def run(queries: torch.Tensor):
            static_q.copy_(queries, non_blocking=True)
            state["graph"].replay()
            return state["out"]

        def release():
            """Collective when an exchange was outgrown meanwhile (its unmapping is a barrier)."""
            if state["graph"] is None:
                return
            state["graph"] = None
            state["out"] = None
            torch.cuda.synchronize(dev)
            engine.pin_workspace(-1)
            self._live_graphs -= 1
            if self._live_graphs == 0:
                for old in getattr(self, "_retired_xchg", []):
                    old.close()
                self._retired_xchg = []

        run.release = release
        return run

    # ------------------------------------------------------------------ passage resolution
    def _owner_and_local(self, gids: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        w = dist_utils.get_world_size()
        if self._sharding == "round_robin":
            return gids % w, gids // w
        starts = np.cumsum([0] + list(self._all_counts))
        owner = np.searchsorted(starts, gids, side="right") - 1
        return owner, gids - starts[owner]

    def _resolve_docs(self, my_ids: torch.Tensor) -> List[List[dict]]:
        """global ids [b,k] -> passage dicts.  Rows of this rank's own shard come out of the in-memory table (one
        vectorised gather instead of b*k dict lookups; the dict OBJECTS of ``doc_map``, like the reference).  Winners
        owned by other ranks are read from the node-shared passage store (passages.py): no collective, no text over
        NVLink.  Only when the ranks do not share a host is the winners' text exchanged through the process group
        (one all-to-all: k winners per query, not the W*k candidates the reference ships through 2*W gathers)."""
        w, r = dist_utils.get_world_size(), dist_utils.get_rank()
        ids_np = my_ids.cpu().numpy()
        if w == 1:
            loc = (ids_np - self._id_base) // self._id_stride
            self.last_passage_path = "local table"
            return self._doc_table()[loc].tolist()
        store = self._shared_passages()
        if store is not None:
            self.last_passage_path = "local table + node-shared passage store (no collective)"
            flat = ids_np.reshape(-1)
            owner, local = self._owner_and_local(flat)
            out = np.empty(flat.shape[0], dtype=object)
            mine = owner == r
            if mine.any():
                out[mine] = self._doc_table()[local[mine]]
            if (~mine).any():
                fetched = store.get_many(owner[~mine], local[~mine])
                tmp = np.empty(len(fetched), dtype=object)
                tmp[:] = fetched
                out[~mine] = tmp
            return out.reshape(ids_np.shape).tolist()
        self.last_passage_path = "all-to-all of the winners' pickled passages"
        all_ids, offs = self._last_all
        all_np = all_ids.cpu().numpy()
        owner, local = self._owner_and_local(all_np.reshape(-1))
        owner, local = owner.reshape(all_np.shape), local.reshape(all_np.shape)
        table = self._doc_table()
        # for every destination rank d: the passages I own among the winners of d's queries
        outgoing = []
        for d in range(w):
            rows = slice(int(offs[d]), int(offs[d + 1]))
            sel = owner[rows] == r
            outgoing.append(dict(zip(all_np[rows][sel].tolist(), table[local[rows][sel]].tolist())))
        merged = {}
        for part in dist_utils.all_to_all_objects(outgoing, device=my_ids.device):
            merged.update(part)
        return [[merged[int(g)] for g in row] for row in ids_np]

    def _shared_passages(self):
        """The node-shared store of every rank's passages, (re)built collectively after the shard layout changed
        (init_embeddings / load_index / assignment to .embeddings all pass through _set_sharding on every rank).
        None: ranks on different hosts, JSA_MIPS_PASSAGES=a2a, or the store could not be written."""
        if os.environ.get("JSA_MIPS_PASSAGES", "store").lower() == "a2a":
            return None
        if getattr(self, "_store_epoch", None) != self._sharding_epoch:
            old = getattr(self, "_pstore", None)
            if old is not None:
                old.close()
            from .passages import PassageStore
            self._pstore = PassageStore.build_shared(self.doc_map, len(self.doc_map))
            self._store_epoch = self._sharding_epoch
        return self._pstore

    def refresh_passages(self) -> None:
        """Collective.  Call after mutating ``doc_map`` in place on a multi-rank index: the other ranks read this
        rank's passages from the node-shared store, which is rebuilt at the next search_knn."""
        self._sharding_epoch += 1

    def _doc_table(self) -> np.ndarray:
        """doc_map ({local row -> passage dict}, the reference's public attribute) as an object array,
        rebuilt only when the dict object or its size changes."""
        key = (id(self.doc_map), len(self.doc_map))
        if getattr(self, "_doc_table_key", None) != key:
            if hasattr(self.doc_map, "as_object_array"):       # a mapping that can hand over its table at once
                tab = self.doc_map.as_object_array()
            else:
                tab = np.empty(len(self.doc_map), dtype=object)
                for i in range(len(self.doc_map)):
                    tab[i] = self.doc_map[i]
            self._doc_table_arr, self._doc_table_key = tab, key
        return self._doc_table_arr

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def search_knn(self, queries, topk, return_embeddings: bool = False):
        """Exhaustive inner-product k-NN (src/index.py:123-158).  Returns ``(docs, scores)`` — docs
        first — as nested Python lists, rows sorted by descending score; with
        ``return_embeddings=True`` also the passage embeddings ``[b, k, dim]`` like the
        build_server twin (build_server/index.py:217-261)."""
        dev = self._store.device if self._store is not None and self._store.is_cuda else None
        with SearchTimer(dev) as timer:
            scores, ids = self.search(queries, topk)
            timer.device_done()
            # (every rank takes part in the passage exchange, also one whose own batch is empty)
            with nvtx_range("mips.resolve_passages"):
                docs = self._resolve_docs(ids) if (scores.shape[0] > 0 or self._any_rank_has_queries) else []
        self.last_search_stats = timer.stats()
        if isinstance(self.iter_stats, dict):
            for key, val in self.last_search_stats.items():
                self.iter_stats[key] = (val, 1)                  # src/rag.py:170 convention
        if scores.shape[0] == 0:
            if return_embeddings:
                emb = self._gather_embeddings(ids) if self._any_rank_has_queries else None
                return [], [], (emb if emb is not None else torch.empty(0, topk, self._store.shape[1], dtype=self.dtype))
            return [], []
        if self.round_scores_to_index_dtype:
            scores = scores.to(self.dtype)
        scores_list = scores.float().tolist()
        if return_embeddings:
            return docs, scores_list, self._gather_embeddings(ids)
        return docs, scores_list

    def _gather_embeddings(self, ids: torch.Tensor) -> torch.Tensor:
        """[b,k] global ids -> [b,k,dim] passage embeddings (build_server/index.py:228-229,254-255).
        Multi rank: every owner fills the slots it owns, one all-reduce(sum) completes the tensor."""
        w, r = dist_utils.get_world_size(), dist_utils.get_rank()
        b, k = ids.shape
        eng = self._get_engine()
        if w == 1:
            loc = (ids - self._id_base) // self._id_stride
            return eng.gather_rows(loc).view(b, k, -1)
        all_ids, offs = self._last_all
        flat = all_ids.reshape(-1)
        owner_np, local_np = self._owner_and_local(flat.cpu().numpy())
        local = torch.from_numpy(np.where(owner_np == r, local_np, -1)).to(ids.device)
        emb = eng.gather_rows(local).view(all_ids.shape[0], k, -1)   # rows with -1 come back as zeros
        torch.distributed.all_reduce(emb)
        return emb[int(offs[r]):int(offs[r + 1])]


class B200IndexWithEmbeddings(B200Index):
    """Twin of build_server/index.py:135-264 — ``search_knn`` returns (docs, scores, embeddings)."""

    @torch.no_grad()
    def search_knn(self, queries, topk):
        return super().search_knn(queries, topk, return_embeddings=True)
