"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's exact-MIPS retrieval path.

This module is the *oracle*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker / the timed CPU baseline.  The product (``jsa-rag_b200``) never imports it and has no
CPU fallback.

Parity pin: the reference ships **no** tests, golden vectors or fixtures for this path
(SURVEY.md §4).  The pin is therefore (a) ``tests/golden/*.npz`` — outputs of the *unmodified*
reference (``/root/reference/src/index.py`` imported under two stubs by ``oracle/ref_import.py``)
on seeded inputs, frozen by ``oracle/make_golden.py``; and (b) a live comparison against that
import whenever ``/root/reference`` is present (``tests/test_oracle.py``).

The arithmetic of the reference lives in a third-party dependency, PyTorch
(``pytorch==1.11.0 cudatoolkit=11.3``, reference ``README.md:6``): ``torch.matmul`` on fp16
operands and ``torch.topk``.  The container's torch 2.11 stands in; semantics are the same:
fp16 inputs, fp32 accumulation, **fp16-rounded output**, selection on the rounded values,
tie order unspecified.  The server path's arithmetic lives in faiss-gpu 1.7.2
(``README.md:7``, not vendored): ``faiss.normalize_L2`` + ``GpuIndexFlatIP`` (fp16 storage,
fp32 queries, exact inner product, descending top-k, ids in insertion order).

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np
import torch


# --------------------------------------------------------------------------------------------
# src/index.py:50-54 — storage: fp16 matrix in [D, N_local] layout
# --------------------------------------------------------------------------------------------
def make_embeddings_dn(embeddings_nd: torch.Tensor) -> torch.Tensor:
    """``torch.zeros(dim, N, dtype=fp16)`` then ``emb[:, a:b] = x.T`` (src/index.py:52, src/rag.py:120)."""
    n, d = embeddings_nd.shape
    emb = torch.zeros(d, n, dtype=torch.float16)
    emb[:, :] = embeddings_nd.T  # slice assignment casts to fp16 with round-to-nearest-even
    return emb


# --------------------------------------------------------------------------------------------
# src/index.py:114-121 — _compute_scores_and_indices
# --------------------------------------------------------------------------------------------
def compute_scores_and_indices(allqueries: torch.Tensor, embeddings_dn: torch.Tensor, topk: int
                               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """scores = matmul(q.half(), E); topk(scores, k, dim=1)  (src/index.py:118-119).

    Returns fp16 scores [B, k] (descending) and int64 local row indices [B, k].
    Raises RuntimeError when ``topk > N_local`` exactly like ``torch.topk``.
    A bf16 matrix (BASELINE configs[3]; not something the reference itself can hold, it always ``.half()``s) is
    scored by the same two lines with ``.bfloat16()`` — a documented extension, not a restatement.
    """
    if embeddings_dn.dtype == torch.bfloat16:
        scores = torch.matmul(allqueries.bfloat16(), embeddings_dn)
        scores, indices = torch.topk(scores, topk, dim=1)
        return scores, indices
    scores = torch.matmul(allqueries.half(), embeddings_dn)
    scores, indices = torch.topk(scores, topk, dim=1)
    return scores, indices


def compute_scores_and_indices_numpy(queries: np.ndarray, embeddings_dn: np.ndarray, topk: int
                                     ) -> Tuple[np.ndarray, np.ndarray]:
    """Same as above in numpy: fp16 operands, fp32 accumulate, fp16-rounded scores, select on those.

    Ties are broken by ascending row index (stable sort) — one valid instance of the
    reference's unspecified tie order.
    """
    q16 = queries.astype(np.float16).astype(np.float32)
    e = embeddings_dn.astype(np.float32)
    s = (q16 @ e).astype(np.float16)
    if topk > s.shape[1]:
        raise RuntimeError("selected index k out of range")
    order = np.argsort(-s.astype(np.float32), axis=1, kind="stable")[:, :topk]
    return np.take_along_axis(s, order, axis=1), order.astype(np.int64)


# --------------------------------------------------------------------------------------------
# src/index.py:123-158 — search_knn, single process (dist not initialised)
# --------------------------------------------------------------------------------------------
def search_knn_single(queries: torch.Tensor, embeddings_dn: torch.Tensor, doc_map: dict, topk: int):
    """Single-rank search_knn: returns (docs, scores) — docs first (src/index.py:158)."""
    scores, indices = compute_scores_and_indices(queries, embeddings_dn, topk)   # :132
    return search_knn_tail(scores, indices, doc_map, topk)


def search_knn_tail(scores: torch.Tensor, indices: torch.Tensor, doc_map, topk: int):
    """The host part of search_knn after the arithmetic (src/index.py:133-134,152-157): B*k doc_map lookups, the
    (single-rank: redundant) second topk and the Python re-indexing.  Separate so that bench.py can time it alone."""
    indices = indices.tolist()                                                  # :133
    docs = [[doc_map[x] for x in row] for row in indices]                        # :134
    _, sub = torch.topk(scores, topk, dim=1)                                     # :152
    scores = scores.tolist()                                                     # :153
    sub = sub.tolist()                                                           # :154
    scores = [[scores[k][j] for j in idx] for k, idx in enumerate(sub)]          # :156
    docs = [[docs[k][j] for j in idx] for k, idx in enumerate(sub)]              # :157
    return docs, scores


# --------------------------------------------------------------------------------------------
# src/index.py:135-157 — cross-rank merge (restated without NCCL / pickling)
# --------------------------------------------------------------------------------------------
def merge_rank_results(per_rank_scores: Sequence[torch.Tensor], per_rank_ids: Sequence[torch.Tensor],
                       topk: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Concatenate each rank's [b, k] block **in rank order** along dim 1, then torch.topk.

    The reference routes (scores, pickled docs) with 2*W gathers (:139-142), concatenates in
    rank order (:143-151) and re-selects (:152-157).  Ids stand in for the docs.
    """
    scores = torch.cat(list(per_rank_scores), dim=1)     # :145
    ids = torch.cat(list(per_rank_ids), dim=1)           # :147-151 (docs merged in rank order)
    top, sub = torch.topk(scores, topk, dim=1)           # :152
    return top, torch.gather(ids, 1, sub)                # :156-157


def shard_rows_round_robin(n_total: int, world_size: int, rank: int) -> np.ndarray:
    """Global rows owned by ``rank`` when passages are loaded from jsonl (src/index_io.py:41)."""
    return np.arange(rank, n_total, world_size, dtype=np.int64)


def shard_rows_contiguous(n_total_shards_sizes: Sequence[int], world_size: int, rank: int) -> np.ndarray:
    """Global rows owned by ``rank`` after load_index (src/index.py:97-100): shard files
    [rank*spw, (rank+1)*spw) concatenated in order."""
    total = len(n_total_shards_sizes)
    assert total % world_size == 0, "N workers must be a multiple of shards to save"   # :96
    spw = total // world_size
    starts = np.concatenate([[0], np.cumsum(n_total_shards_sizes)])
    lo, hi = starts[rank * spw], starts[(rank + 1) * spw]
    return np.arange(lo, hi, dtype=np.int64)


def search_sharded(queries_per_rank: Sequence[torch.Tensor], embeddings_nd: torch.Tensor, world_size: int,
                   topk: int, sharding: str = "round_robin") -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Emulates a W-rank search_knn (src/index.py:123-158) in one process.

    Every rank sees *all* queries (varsize_all_gather, src/dist_utils.py:47-71), scores its own
    row shard, and rank r receives the rows of its own queries from every shard, concatenated in
    rank order, then re-selects.  Returns, per rank, (scores fp16 [b_r, k], global ids int64 [b_r, k]).
    """
    n = embeddings_nd.shape[0]
    allq = torch.cat(list(queries_per_rank), dim=0)                       # :128
    sizes = np.cumsum([0] + [int(q.shape[0]) for q in queries_per_rank])  # :129-130
    shard_scores, shard_ids = [], []
    for r in range(world_size):
        if sharding == "round_robin":
            rows = shard_rows_round_robin(n, world_size, r)
        else:
            per = math.ceil(n / world_size)
            rows = np.arange(r * per, min(n, (r + 1) * per), dtype=np.int64)
        emb_dn = make_embeddings_dn(embeddings_nd[torch.from_numpy(rows)])
        s, i = compute_scores_and_indices(allq, emb_dn, topk)              # :132
        shard_scores.append(s)
        shard_ids.append(torch.from_numpy(rows)[i])
    out = []
    for r in range(world_size):
        sl = slice(int(sizes[r]), int(sizes[r + 1]))
        out.append(merge_rank_results([s[sl] for s in shard_scores], [i[sl] for i in shard_ids], topk))
    return out


# --------------------------------------------------------------------------------------------
# build_server/index.py:217-261 — 3-tuple variant (adds gathered passage embeddings)
# --------------------------------------------------------------------------------------------
def gather_result_embeddings(embeddings_dn: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """embeddings[:, indices.view(-1)].T.view(B, k, D)  (build_server/index.py:228-229)."""
    e = embeddings_dn[:, indices.reshape(-1)]
    return e.transpose(0, 1).contiguous().view(indices.size(0), indices.size(1), -1)


# --------------------------------------------------------------------------------------------
# src/rag.py:228-246 — tail of retrieve_with_rerank (after the encoder)
# --------------------------------------------------------------------------------------------
def rerank_tail(query_emb: torch.Tensor, passage_emb: torch.Tensor, topk: int):
    """einsum("id,ijd->ij") -> sort descending -> first topk -> gather embeddings (src/rag.py:228-233) and
    the two statistics of src/rag.py:236-240.  Pinned by tests/golden/rerank_*.npz, which hold outputs of the
    unmodified RAG.retrieve_with_rerank (oracle/ref_import.run_reference_rerank).
    Returns (scores [B,k], positions [B,k], emb [B,k,D], mrr, mrr_rev)."""
    bsz = query_emb.shape[0]
    scores = torch.einsum("id,ijd->ij", query_emb, passage_emb)
    sorted_scores, sorted_ids = torch.sort(scores, dim=-1, descending=True)
    top_s, top_i = sorted_scores[:, :topk], sorted_ids[:, :topk]
    emb = torch.gather(passage_emb, 1, top_i.unsqueeze(2).expand(bsz, topk, passage_emb.size(-1)))
    mrr = 1 / (top_i.float() + 1).mean(-1).mean().item()
    _, rev = torch.sort(sorted_ids, dim=-1)
    mrr_rev = 1 / (rev[:, :topk].float() + 1).mean(-1).mean().item()
    return top_s, top_i, emb, mrr, mrr_rev


# --------------------------------------------------------------------------------------------
# src/index.py:195-223 — DistributedFAISSIndex with faiss_index_type="flat" 
# --------------------------------------------------------------------------------------------
def faiss_flat_search(queries: torch.Tensor, embeddings_dn_fp16: torch.Tensor, topk: int
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
    """GpuIndexFlatIP over ``_cast_to_torch32(self.embeddings.T)`` searched with ``_cast_to_torch32(allqueries)``,
    scores cast ``.half()`` (src/index.py:205,217,223).  The stored matrix is the fp16 one of init_embeddings
    (src/index.py:52, inherited), so the vectors faiss holds are fp32 copies of fp16-rounded values; the QUERIES stay
    fp32 (the flat DistributedIndex rounds them to fp16, :118) and the products are accumulated in fp32.
    faiss-gpu 1.7.2 is not installed here: PARITY UNPINNED — this restates the documented semantics of IndexFlatIP
    (exact inner products, larger first)."""
    s = queries.float() @ embeddings_dn_fp16.float()
    top, idx = torch.topk(s, topk, dim=1)
    return top.half(), idx


# --------------------------------------------------------------------------------------------
# build_server/server_start.py:139-163 — faiss server search (normalise queries, exact IP)
# --------------------------------------------------------------------------------------------
def normalize_l2(x: np.ndarray) -> np.ndarray:
    """faiss.normalize_L2 (server_start.py:142): per-row x / ||x||_2 in fp32; zero rows unchanged."""
    x = np.ascontiguousarray(x, dtype=np.float32).copy()
    nrm = np.sqrt((x.astype(np.float32) ** 2).sum(axis=1, dtype=np.float32))
    nz = nrm > 0
    x[nz] = x[nz] / nrm[nz, None]
    return x


def server_search(query_embs: np.ndarray, embeddings_nd_fp16: np.ndarray, topk: int
                  ) -> Tuple[np.ndarray, np.ndarray]:
    """D, I = IndexShards(GpuIndexFlatIP fp16).search(normalize_L2(q), k)  (server_start.py:141-145).

    Vectors are stored in fp16 (GpuClonerOptions.useFloat16, :56), queries stay fp32, distances are
    fp32 inner products, ids are global insertion order (successive_ids=True, :45).
    """
    q = normalize_l2(query_embs)
    e = embeddings_nd_fp16.astype(np.float16).astype(np.float32)
    s = q @ e.T
    if topk > s.shape[1]:
        raise RuntimeError("selected index k out of range")
    order = np.argsort(-s, axis=1, kind="stable")[:, :topk]
    return np.take_along_axis(s, order, axis=1).astype(np.float32), order.astype(np.int64)


# --------------------------------------------------------------------------------------------
# Exact scores + the tolerance-aware comparator
# --------------------------------------------------------------------------------------------
def exact_scores(queries: np.ndarray, embeddings_nd: np.ndarray, q_dtype=np.float16) -> np.ndarray:
    """fp64 inner products of the *stored* operands (queries rounded to the index dtype first,
    like ``allqueries.half()`` at src/index.py:118).  Ground truth for near-tie decisions."""
    q = np.asarray(queries)
    if q_dtype is not None:
        q = _round_to(q, q_dtype)
    return q.astype(np.float64) @ np.asarray(embeddings_nd).astype(np.float64).T


def _round_to(x: np.ndarray, dtype) -> np.ndarray:
    if dtype == "bf16":
        return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).bfloat16().float().numpy()
    return x.astype(dtype)


def compare_topk(engine_ids: np.ndarray, engine_scores: np.ndarray, ref_ids: np.ndarray,
                 ref_scores: np.ndarray, exact: np.ndarray, rtol: float = 1e-3, atol: float = 1e-6) -> dict:
    """Tolerance-aware top-k comparison.  PASS iff for every query row:

    (i)   every engine id has exact score >= t - tol, t = reference's k-th returned score;
    (ii)  every id whose exact score is > t + tol is returned by the engine;
    (iii) engine scores are within rtol of the exact score of the id they are attached to, and
          of the reference's score for ids both returned;
    (iv)  engine rows are non-increasing and ids are unique.
    Returns a report dict; ``report["ok"]`` is the verdict and ``near_tie_diffs`` counts id-set
    differences that are attributable to near-ties.
    """
    engine_ids = np.asarray(engine_ids)
    ref_ids = np.asarray(ref_ids)
    es = np.asarray(engine_scores, dtype=np.float64)
    rs = np.asarray(ref_scores, dtype=np.float64)
    b, k = engine_ids.shape
    rep = dict(ok=True, rows=b, k=k, id_set_equal_rows=0, near_tie_diffs=0, errors=[])
    for r in range(b):
        ex = exact[r]
        t = rs[r, -1]
        tol = rtol * abs(t) + atol
        eid = engine_ids[r]
        if len(set(eid.tolist())) != k:
            rep["errors"].append((r, "duplicate ids"))
        if np.any(np.diff(es[r]) > 1e-7 * np.maximum(1.0, np.abs(es[r][:-1]))):
            rep["errors"].append((r, "scores not non-increasing"))
        if np.any(ex[eid] < t - tol):
            rep["errors"].append((r, f"id below reference k-th score: min {ex[eid].min()} vs t {t}"))
        must = np.nonzero(ex > t + tol)[0]
        missing = np.setdiff1d(must, eid)
        if missing.size:
            rep["errors"].append((r, f"{missing.size} clearly-better ids missing, e.g. {missing[:4]}"))
        err = np.abs(es[r] - ex[eid])
        if np.any(err > rtol * np.abs(ex[eid]) + atol):
            rep["errors"].append((r, f"score vs exact: max err {err.max()}"))
        common, ei, ri = np.intersect1d(eid, ref_ids[r], return_indices=True)
        if common.size:
            err2 = np.abs(es[r][ei] - rs[r][ri])
            if np.any(err2 > rtol * np.abs(rs[r][ri]) + atol):
                rep["errors"].append((r, f"score vs reference: max err {err2.max()}"))
        ndiff = k - common.size
        if ndiff == 0:
            rep["id_set_equal_rows"] += 1
        rep["near_tie_diffs"] += ndiff
    rep["ok"] = not rep["errors"]
    return rep


# --------------------------------------------------------------------------------------------
# src/index.py:62-88 — save_index shard geometry (file naming + column ranges)
# --------------------------------------------------------------------------------------------
def shard_ranges(n_embeddings: int, total_saved_shards: int, world_size: int = 1, rank: int = 0):
    """[(shard_id, start, end)] exactly as save_index computes them (src/index.py:73-80)."""
    assert total_saved_shards % world_size == 0, "N workers must be a multiple of shards to save"
    spw = total_saved_shards // world_size
    per = math.ceil(n_embeddings / spw)
    out = []
    for shard_ind, start in enumerate(range(0, n_embeddings, per)):
        out.append((shard_ind + rank * spw, start, min(start + per, n_embeddings)))
    return out


# --------------------------------------------------------------------------------------------
# Timed CPU baseline (bench.py cpu_baseline / --impl reference): the reference's two lines
# --------------------------------------------------------------------------------------------
def cpu_search_arith(queries: torch.Tensor, embeddings_dn: torch.Tensor, topk: int):
    """The arithmetic of the reference CPU path (src/index.py:118-119) for timing."""
    with torch.no_grad():
        return compute_scores_and_indices(queries, embeddings_dn, topk)
