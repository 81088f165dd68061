"""Index server — same request/response contract as the reference's faiss server
(build_server/server_start.py), backed by the B200 engine.

    POST /retrieve {"query_embs": [bsz*dim floats], "bsz": int = 1, "topk": int = 10} -> [docs, scores]
    POST /rebuild  {"checkpoint_path": str, "response_url": str}                     -> swaps the index

Reference semantics kept (build_server/server_start.py:139-163): queries are L2-normalised
(faiss.normalize_L2; passages are NOT normalised), exact inner product over fp16-stored vectors,
ids = global insertion order across the embedding files (IndexShards(successive_ids=True), :45),
file i -> GPU i (:75-77,97).  Differences underneath: the pickle streams are read ONCE straight
into the device matrix (the reference unpickles everything twice, :63-95), each GPU runs the fused
scan on its own stream, writes its top-k into a packed block that is peer-copied into device 0's
merge buffer, device 0 merges once every block's event has fired (no host synchronisation between
the GPUs), and /rebuild swaps the index atomically (a search in flight keeps the old index alive).
"""
from __future__ import annotations

import json
import os
import pickle
import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .engine import MipsEngine, merge_packed, merge_topk, packed_result_buffer

try:  # FastAPI is only needed for the HTTP shell, not for the index itself
    from fastapi import FastAPI, HTTPException, Request, Response
    from pydantic import BaseModel

    class RetrieveRequest(BaseModel):   # build_server/server_start.py:18-21
        query_embs: list
        bsz: int = 1
        topk: int = 10

    class rebuildRequest(BaseModel):    # build_server/server_start.py:23-25
        checkpoint_path: str
        response_url: str
except ImportError:  # pragma: no cover
    FastAPI = None


def parse_retrieve_request(body: bytes) -> dict:
    """``{"query_embs": [bsz*dim floats], "bsz": int = 1, "topk": int = 10}`` -> dict with an fp32 array.  Raises
    HTTPException(422) for what pydantic would reject (missing list, wrong types, bsz not dividing the list)."""
    import ctypes

    def bad(msg):
        if FastAPI is not None:
            raise HTTPException(status_code=422, detail=msg)
        raise ValueError(msg)

    arr = None
    a = body.find(b"[", body.find(b'"query_embs"') + 1) if b'"query_embs"' in body else -1
    b = body.find(b"]", a + 1) if a >= 0 else -1
    if a >= 0 and b > a and body.find(b"[", a + 1, b) < 0:            # a flat list: the fast scanner applies
        from . import _native as N
        lib = N.load()
        out = np.empty((b - a) // 2 + 1, dtype=np.float32)             # every number needs >= 2 bytes ("1,")
        n = lib.mips_parse_float_list(body[a:b + 1], b + 1 - a, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), out.size)
        if n >= 0:
            arr = out[:n]
            body = body[:a] + b"null" + body[b + 1:]
    try:
        req = json.loads(body)
    except ValueError:
        bad("body is not JSON")
    if not isinstance(req, dict):
        bad("body must be a JSON object")
    if arr is None:
        if not isinstance(req.get("query_embs"), list):
            bad("query_embs: a list is required")
        try:
            arr = np.asarray(req["query_embs"], dtype=np.float32).reshape(-1)
        except (TypeError, ValueError):
            bad("query_embs must hold numbers")
    bsz, topk = req.get("bsz", 1), req.get("topk", 10)
    if not isinstance(bsz, int) or not isinstance(topk, int) or isinstance(bsz, bool) or isinstance(topk, bool):
        bad("bsz and topk must be integers")
    if bsz <= 0 or arr.size % bsz != 0:
        bad("query_embs does not hold bsz rows")
    return {"query_embs": arr, "bsz": bsz, "topk": topk}


def iter_embedding_stream(path: str):
    """Yields the batches appended by the reference's builder: each ``pickle.dump`` is a list of
    ``{"emb": np.ndarray[dim], "passage": dict}`` (build_server/index.py:108-111)."""
    with open(path, "rb") as f:
        while True:
            try:
                yield pickle.load(f)
            except EOFError:
                return


def append_embedding_batch(path: str, embs: np.ndarray, passages: Sequence[dict]) -> None:
    """Writes one batch in the reference's stream format (used by tests / tooling)."""
    with open(path, "ab") as f:
        pickle.dump([{"emb": embs[i], "passage": passages[i]} for i in range(len(passages))], f)


def get_pkl_files_in_directory(directory: str) -> List[str]:
    """build_server/server_start.py:164-170 (os.walk order is not sorted there; it is here)."""
    out = []
    for root, _, files in os.walk(directory):
        for file in sorted(files):
            if file.endswith(".pkl"):
                out.append(os.path.join(root, file))
    return sorted(out)


class B200ServerIndex(object):
    """Drop-in for ``DistributedFaissIndex`` (build_server/server_start.py:30-163)."""

    def __init__(self, embedding_file_pathes: Sequence[str], checkpoint_pathes=None, gpu_ids: Optional[Sequence[int]] = None,
                 dtype: torch.dtype = torch.float16, dimension: Optional[int] = None):
        n_dev = torch.cuda.device_count() if torch.cuda.is_available() else 0
        if n_dev == 0:
            raise RuntimeError("B200ServerIndex needs CUDA devices; there is no CPU fallback")
        gpu_ids = list(range(min(len(embedding_file_pathes), n_dev))) if gpu_ids is None else list(gpu_ids)
        for g in gpu_ids:
            if g >= n_dev:
                raise ValueError(f"GPU {g} not available (Total GPUs: {n_dev})")   # server_start.py:50-51
        if len(embedding_file_pathes) > len(gpu_ids):
            raise ValueError("more embedding files than GPUs: the reference maps file i to GPU i")
        self.doc_map = {}
        self.shards: List[MipsEngine] = []
        self.dtype = dtype
        last_id = 0
        for file_idx, path in enumerate(embedding_file_pathes):
            dev = torch.device("cuda", gpu_ids[file_idx])
            chunks = []
            first = last_id
            for batch in iter_embedding_stream(path):                  # one pass (reference: two)
                embs = np.stack([np.asarray(d["emb"]) for d in batch]).astype(np.float16, copy=False)
                for d in batch:
                    self.doc_map[last_id] = d["passage"]                  # server_start.py:84-87
                    last_id += 1
                chunks.append(torch.from_numpy(embs))
            if not chunks:
                continue
            dim = int(chunks[0].shape[1]) if dimension is None else int(dimension)
            n = sum(c.shape[0] for c in chunks)
            store = torch.empty(n, dim, dtype=dtype, device=dev)
            at = 0
            for c in chunks:
                store[at:at + c.shape[0]].copy_(c, non_blocking=True)
                at += c.shape[0]
            eng = MipsEngine(dim, dtype, dev)
            eng.bind(store, id_base=first, id_stride=1)                   # successive ids across shards
            self.shards.append(eng)
        if not self.shards:
            raise ValueError("no embeddings found")
        self._finish_init(last_id)

    def _finish_init(self, ntotal: int) -> None:
        self.dimension = self.shards[0].dim
        self.ntotal = ntotal
        self._lock = threading.Lock()          # one search at a time owns the staging buffers below
        self._stage = {}                       # (batch, k) -> per-shard packed blocks + device-0 merge buffer
        self._streams = [torch.cuda.Stream(device=e.device) for e in self.shards]

    @classmethod
    def from_tensors(cls, stores: Sequence[torch.Tensor], doc_map, dtype: torch.dtype = torch.float16) -> "B200ServerIndex":
        """Index over matrices that already live on their GPUs ([n_i, dim] each; shard i = ``stores[i]``, ids are
        successive across shards like IndexShards(successive_ids=True)).  ``doc_map``: id -> passage."""
        self = cls.__new__(cls)
        self.doc_map, self.shards, self.dtype = doc_map, [], dtype
        first = 0
        for st in stores:
            eng = MipsEngine(int(st.shape[1]), dtype, st.device)
            eng.bind(st, id_base=first, id_stride=1)
            self.shards.append(eng)
            first += int(st.shape[0])
        self._finish_init(first)
        return self

    def _staging(self, batch: int, k: int):
        key = (batch, k)
        st = self._stage.get(key)
        if st is None:
            if len(self._stage) > 8:
                self._stage.clear()
            dev0 = self.shards[0].device
            merged, _, _ = packed_result_buffer(batch, k, dev0, lists=len(self.shards))
            local = [packed_result_buffer(batch, k, e.device) for e in self.shards]
            st = self._stage[key] = (merged, local)
        return st

    @torch.no_grad()
    def search(self, query_embs: torch.Tensor, topk: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """normalize_L2(q) -> exact IP top-k on every shard -> merge on the first device.

        Every GPU works on its own stream: queries in, fused scan + select straight into a packed
        [scores | ids] block, one peer copy of that block (12*B*k bytes) into device 0's merge buffer, an
        event; device 0 waits for the events and merges all blocks in one launch.  Nothing synchronises with the
        host until the caller reads the result."""
        q = query_embs.to(torch.float32)
        b = int(q.shape[0])
        if b == 0 or any(e.n_local < topk for e in self.shards):
            return self._search_general(q, topk)
        dev0 = self.shards[0].device
        if q.device.type == "cpu" and not q.is_pinned():
            q = q.pin_memory()
        with self._lock:
            merged, local = self._staging(b, topk)
            cur0 = torch.cuda.current_stream(dev0)
            start = torch.cuda.Event()
            start.record(cur0)
            done = []
            for g, eng in enumerate(self.shards):
                stream = self._streams[g]
                stream.wait_event(start)                        # the staging buffers are free again
                with torch.cuda.device(eng.device), torch.cuda.stream(stream):
                    _, ls, li = local[g]
                    eng.search(q.to(eng.device, non_blocking=True), topk, normalize=True, out=(ls[0], li[0]))
                    merged[g].copy_(local[g][0][0], non_blocking=True)       # peer copy into device 0
                    ev = torch.cuda.Event()
                    ev.record(stream)
                    done.append(ev)
            for ev in done:
                cur0.wait_event(ev)
            with torch.cuda.device(dev0):
                if len(self.shards) == 1:
                    _, ls, li = local[0]
                    return ls[0].clone(), li[0].clone()
                return merge_packed(merged, b, topk, topk)

    @torch.no_grad()
    def _search_general(self, q: torch.Tensor, topk: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Shards smaller than topk (tiny indices) / empty batches: per-shard k differs, lists are padded."""
        parts_s, parts_i = [], []
        for eng in self.shards:
            with torch.cuda.device(eng.device):
                s, i = eng.search(q.to(eng.device, non_blocking=True), min(topk, eng.n_local), normalize=True)
            parts_s.append(s)
            parts_i.append(i)
        if len(self.shards) == 1 and parts_s[0].shape[1] == topk:
            return parts_s[0], parts_i[0]
        dev0 = self.shards[0].device
        kmax = max(p.shape[1] for p in parts_s)
        ss = torch.full((len(parts_s), q.shape[0], kmax), float("-inf"), device=dev0)
        ii = torch.full((len(parts_s), q.shape[0], kmax), -1, dtype=torch.int64, device=dev0)
        for n, (s, i) in enumerate(zip(parts_s, parts_i)):
            torch.cuda.synchronize(s.device)
            ss[n, :, : s.shape[1]] = s.to(dev0)
            ii[n, :, : i.shape[1]] = i.to(dev0)
        with torch.cuda.device(dev0):
            return merge_topk(ss, ii, topk)

    def search_knn(self, query_embs, topk):
        """build_server/server_start.py:139-163: returns (all_docs, all_scores) nested lists."""
        if int(topk) > self.ntotal:
            raise RuntimeError("selected index k out of range")
        D, I = self.search(torch.as_tensor(query_embs), int(topk))
        ids = I.cpu().numpy()
        all_scores = D.cpu().tolist()                              # float(...) of every score (:160)
        if hasattr(self.doc_map, "as_object_array") or isinstance(self.doc_map, dict):
            all_docs = self._doc_table()[ids].tolist()             # one gather instead of B*k dict lookups (:157-159)
        else:
            all_docs = [[self.doc_map[int(i)] for i in row] for row in ids]
        return all_docs, all_scores

    def _doc_table(self) -> np.ndarray:
        key = (id(self.doc_map), len(self.doc_map))
        if getattr(self, "_doc_table_key", None) != key:
            if hasattr(self.doc_map, "as_object_array"):
                tab = self.doc_map.as_object_array()
            else:
                tab = np.empty(len(self.doc_map), dtype=object)
                for i in range(len(self.doc_map)):
                    tab[i] = self.doc_map[i]
            self._doc_table_arr, self._doc_table_key = tab, key
        return self._doc_table_arr


class IndexHolder(object):
    """Atomically swappable reference to the live index (the reference mutates a module global while
    `async def retrieve` may be running, server_start.py:181-196)."""

    def __init__(self, index=None):
        self._index = index
        self._lock = threading.Lock()

    def get(self):
        with self._lock:
            return self._index

    def swap(self, index):
        with self._lock:
            old, self._index = self._index, index
        return old


def create_app(holder: IndexHolder, rebuild_fn=None, notify=None):
    """FastAPI app with the reference's two routes (server_start.py:181-196)."""
    if FastAPI is None:
        raise RuntimeError("fastapi / pydantic are required for the HTTP server")
    app = FastAPI()

    def _json_answer(relevant_docs, scores) -> "Response":
        # one json.dumps of the nested lists; FastAPI's default path (jsonable_encoder walking B*k passage dicts,
        # then dumps) costs ~4x as much for a 64 x 100 answer
        return Response(content=json.dumps([relevant_docs, scores]), media_type="application/json")

    @app.post("/retrieve")
    async def retrieve(request: Request):
        """The reference's route and schema (RetrieveRequest: query_embs, bsz = 1, topk = 10; :18-21,181-189).  The
        body is parsed by hand: the flat float list goes through the C scanner (mips_parse_float_list), the rest
        through json — pydantic validation of 65k list items is most of the reference route's latency."""
        index = holder.get()
        if index is None:
            raise HTTPException(status_code=500, detail="Index is not ready")      # :184-185
        req = parse_retrieve_request(await request.body())
        query_embs = torch.from_numpy(req["query_embs"]).view(req["bsz"], -1)         # :186
        relevant_docs, scores = index.search_knn(query_embs, req["topk"])             # :188
        return _json_answer(relevant_docs, scores)                                  # :189

    # Binary fast path next to the reference's JSON one: the request body is the raw little-endian
    # [bsz, dim] query matrix (fp32 or fp16), so a batch of 64 x 1024 floats is 256 KB of bytes instead
    # of ~1.2 MB of decimal text that has to be parsed float by float.
    def _queries_from_body(body: bytes, bsz: int, dtype: str) -> torch.Tensor:
        np_dtype = {"fp32": np.float32, "fp16": np.float16}.get(dtype)
        if np_dtype is None or bsz <= 0 or len(body) % (bsz * np.dtype(np_dtype).itemsize) != 0:
            raise HTTPException(status_code=422, detail="body must be bsz x dim little-endian fp32/fp16 values")
        return torch.from_numpy(np.frombuffer(body, dtype=np_dtype).reshape(bsz, -1).copy())

    @app.post("/retrieve_bin")
    async def retrieve_bin(request: Request, bsz: int = 1, topk: int = 10, dtype: str = "fp32"):
        """Same answer as /retrieve ([docs, scores] as JSON), binary request."""
        index = holder.get()
        if index is None:
            raise HTTPException(status_code=500, detail="Index is not ready")
        relevant_docs, scores = index.search_knn(_queries_from_body(await request.body(), bsz, dtype), topk)
        return _json_answer(relevant_docs, scores)

    @app.post("/search_bin")
    async def search_bin(request: Request, bsz: int = 1, topk: int = 10, dtype: str = "fp32"):
        """Binary both ways, for callers that hold the passage store themselves: the response body is
        bsz*topk fp32 scores followed by bsz*topk int64 passage ids (insertion order)."""
        index = holder.get()
        if index is None:
            raise HTTPException(status_code=500, detail="Index is not ready")
        D, I = index.search(_queries_from_body(await request.body(), bsz, dtype), topk)
        payload = D.detach().cpu().numpy().astype("<f4").tobytes() + I.detach().cpu().numpy().astype("<i8").tobytes()
        return Response(content=payload, media_type="application/octet-stream")

    @app.post("/rebuild")
    def rebuild(request: rebuildRequest):
        if rebuild_fn is None:
            raise HTTPException(status_code=501, detail="rebuild is not configured")
        holder.swap(rebuild_fn(request.checkpoint_path))                             # :193-195
        if notify is not None:
            notify(request.response_url, {"status": "success"})                      # :196
        else:
            import requests
            requests.post(request.response_url, json={"status": "success"})
        return {"status": "success"}

    return app


def serve(embedding_dir: str, host: str = "0.0.0.0", port: int = 29501, gpu_ids=None):  # pragma: no cover
    """`python -m` entry: the reference hard-codes directory and port (server_start.py:171-175,201)."""
    import uvicorn

    files = get_pkl_files_in_directory(embedding_dir)
    holder = IndexHolder(B200ServerIndex(files, None, gpu_ids=gpu_ids))
    uvicorn.run(create_app(holder, rebuild_fn=lambda ckpt: B200ServerIndex(files, ckpt, gpu_ids=gpu_ids)), host=host, port=port)
