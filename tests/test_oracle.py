"""The oracle (CPU restatement of the reference path) against the frozen outputs of the UNMODIFIED
reference (tests/golden, written by oracle/make_golden.py) and, when /root/reference is present,
against the live reference."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, GOLDEN_DIR, load_golden
from oracle import flat_index_oracle as O
from oracle import ref_import


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_golden(name):
    g = load_golden(name)
    k = int(g["k"])
    e = torch.from_numpy(g["embeddings"])
    q = torch.from_numpy(g["queries"])
    emb_dn = O.make_embeddings_dn(e)
    scores, idx = O.compute_scores_and_indices(q, emb_dn, k)
    # the fp16 score rows must be bit-identical (tie order may differ, values may not)
    assert np.array_equal(scores.numpy(), g["scores"].astype(np.float16))
    exact = O.exact_scores(g["queries"], g["embeddings"])
    rep = O.compare_topk(idx.numpy(), scores.float().numpy(), g["ids"], g["scores"].astype(np.float32), exact)
    assert rep["ok"], rep["errors"][:3]
    # numpy restatement: same fp16 score multiset per row, ids within tie tolerance
    s2, i2 = O.compute_scores_and_indices_numpy(g["queries"], emb_dn.numpy(), k)
    rep2 = O.compare_topk(i2, s2.astype(np.float32), g["ids"], g["scores"].astype(np.float32), exact)
    assert rep2["ok"], rep2["errors"][:3]
    assert np.abs(s2.astype(np.float32) - g["scores"].astype(np.float32)).max() <= 1e-3 * np.abs(g["scores"]).max()


def test_search_knn_single_returns_docs_first():
    g = load_golden("flat_n1003_d768_b8_k20")
    n = int(g["n"])
    doc_map = {i: {"id": str(i)} for i in range(n)}
    docs, scores = O.search_knn_single(torch.from_numpy(g["queries"]), O.make_embeddings_dn(torch.from_numpy(g["embeddings"])),
                                       doc_map, 20)
    assert isinstance(docs[0][0], dict) and isinstance(scores[0][0], float)
    ids = np.array([[int(d["id"]) for d in row] for row in docs])
    exact = O.exact_scores(g["queries"], g["embeddings"])
    assert O.compare_topk(ids, np.array(scores), g["ids"], g["scores"].astype(np.float32), exact)["ok"]


def test_pinned_behaviours():
    b = np.load(os.path.join(GOLDEN_DIR, "behaviours.npz"))
    assert int(b["empty_docs_len"]) == 0 and int(b["empty_scores_len"]) == 0
    assert "selected index k out of range" in str(b["k_too_large_error"])
    e = torch.randn(64, 768)
    with pytest.raises(RuntimeError, match="selected index k out of range"):
        O.compute_scores_and_indices(torch.randn(2, 768), O.make_embeddings_dn(e), 65)
    with pytest.raises(RuntimeError, match="selected index k out of range"):
        O.compute_scores_and_indices_numpy(np.random.randn(2, 768), O.make_embeddings_dn(e).numpy(), 65)
    # shard file geometry of save_index(dir, 4) with N=1003 (src/index.py:74-80)
    assert list(b["shard_files"]) == [f"embeddings.{i}.pt" for i in range(4)] + [f"passages.{i}.pt" for i in range(4)]
    rng = O.shard_ranges(1003, 4)
    assert [(s, e - st) for s, st, e in rng] == [(0, 251), (1, 251), (2, 251), (3, 250)]
    for (sid, st, en), rep in zip(rng, b["shard_shapes"]):
        assert f"(768, {en - st})" in str(rep) and "float16" in str(rep) and "True" in str(rep)


@pytest.mark.parametrize("sharding", ["round_robin", "contiguous"])
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_merge_equals_single_index(world, sharding):
    """src/index.py:135-157: concatenating per-rank top-k in rank order and re-selecting gives the
    global top-k (up to fp16 ties)."""
    g = load_golden("flat_n1003_d768_b8_k20")
    e = torch.from_numpy(g["embeddings"]).float()
    q = torch.from_numpy(g["queries"])
    sizes = [3, 5] if world == 2 else [1, 3, 0, 4]
    qs = list(torch.split(q, sizes))
    out = O.search_sharded(qs, e, world, 20, sharding)
    ids = torch.cat([o[1] for o in out]).numpy()
    scores = torch.cat([o[0] for o in out]).float().numpy()
    exact = O.exact_scores(g["queries"], g["embeddings"])
    rep = O.compare_topk(ids, scores, g["ids"], g["scores"].astype(np.float32), exact)
    assert rep["ok"], rep["errors"][:3]


def test_server_search_normalises_queries_only():
    g = load_golden("flat_n300_d1024_b5_k10")
    q = g["queries"] * np.array([[3.0], [0.5], [10.0], [1.0], [7.0]], dtype=np.float32)
    d, i = O.server_search(q, g["embeddings"], 10)
    d1, i1 = O.server_search(g["queries"], g["embeddings"], 10)
    assert np.array_equal(i, i1) and np.allclose(d, d1, rtol=1e-5)
    z = O.normalize_l2(np.zeros((2, 8), dtype=np.float32))
    assert np.all(z == 0)


def test_comparator_rejects_wrong_results():
    g = load_golden("flat_n1003_d768_b8_k20")
    exact = O.exact_scores(g["queries"], g["embeddings"])
    ids = g["ids"].copy()
    sc = g["scores"].astype(np.float32)
    assert O.compare_topk(ids, sc, g["ids"], sc, exact)["ok"]
    bad = ids.copy()
    bad[0, 0] = int(np.argmin(exact[0]))  # clearly not a neighbour
    assert not O.compare_topk(bad, sc, g["ids"], sc, exact)["ok"]
    bad_s = sc.copy()
    bad_s[1, 3] *= 1.01
    assert not O.compare_topk(ids, bad_s, g["ids"], sc, exact)["ok"]


@pytest.mark.skipif(not ref_import.reference_available(), reason="/root/reference not present (GPU box)")
def test_oracle_against_live_reference():
    """Runs the unmodified reference (stub-imported) on fresh seeded inputs."""
    from oracle import make_golden
    e16, q = make_golden.synth_inputs(3000, 768, 12, seed=99)
    ids, scores, idx = make_golden.run_reference(e16, q, 50)
    s, i = O.compute_scores_and_indices(torch.from_numpy(q), O.make_embeddings_dn(torch.from_numpy(e16)), 50)
    assert np.array_equal(s.numpy().astype(np.float32), scores)
    exact = O.exact_scores(q, e16)
    assert O.compare_topk(i.numpy(), s.float().numpy(), ids, scores, exact)["ok"]
    assert idx.search_knn(torch.from_numpy(q[:0]), 5) == ([], [])


def test_faiss_flat_mode_is_within_tolerance_of_the_flat_index():
    """index_mode="faiss", faiss_index_type="flat" is served by the same engine.  The reference's faiss
    flat path differs from its torch flat path only in keeping the QUERIES in fp32 (src/index.py:217 vs :118); on
    unit-norm Contriever-shaped vectors that moves a top-k score by < 1e-3 relative (north_star's tolerance), and the
    `.half()` both paths apply to the returned scores (:223 / :118) rounds by up to 4.9e-4 on its own."""
    g = torch.Generator().manual_seed(11)
    e = torch.nn.functional.normalize(torch.randn(60_000, 768, generator=g), dim=1)
    q = torch.nn.functional.normalize(torch.randn(32, 768, generator=g), dim=1)
    e_dn = O.make_embeddings_dn(e)
    fs, fi = O.faiss_flat_search(q, e_dn, 100)
    ts, ti = O.compute_scores_and_indices(q, e_dn, 100)
    exact = O.exact_scores(q.numpy(), e_dn.T.contiguous().numpy(), q_dtype=np.float32)     # fp32 queries, fp16-stored rows
    rep = O.compare_topk(ti.numpy(), ts.float().numpy(), fi.numpy(), fs.float().numpy(), exact, rtol=1e-3)
    assert rep["ok"], rep["errors"][:3]
    unrounded = torch.gather(q @ e_dn.float(), 1, fi)
    rel = ((fs.float() - unrounded).abs() / unrounded.abs()).max().item()
    assert rel <= 2 ** -11 * 1.001                                                          # fp16 output rounding: half an ulp
