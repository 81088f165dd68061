"""Re-rank tail (reference src/rag.py:228-246): oracle vs the frozen outputs of the unmodified reference (CPU),
and the fused CUDA op through the C ABI vs both (GPU)."""
import ctypes
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import flat_index_oracle as O
from oracle import ref_import

RERANK_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "rerank_*.npz")))
_DT = {"float32": torch.float32, "bfloat16": torch.bfloat16, "float16": torch.float16}


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    dt = _DT[str(g["dtype"])]
    g["q"] = torch.from_numpy(g["query_emb"]).to(dt)
    g["p"] = torch.from_numpy(g["passage_emb"]).to(dt)
    return g


def _check_against_exact(pos, scores, g, rel=1e-3):
    """Tolerance-aware comparison: the k-th exact score bounds every returned candidate, every
    candidate clearly above it is present, scores agree within `rel`, rows are non-increasing."""
    exact = np.einsum("id,ijd->ij", g["q"].double().numpy(), g["p"].double().numpy())
    k = pos.shape[1]
    for i in range(pos.shape[0]):
        kth = np.sort(exact[i])[::-1][k - 1]
        tol = rel * max(abs(kth), 1e-3) + (2e-2 if str(g["dtype"]) != "float32" else 0) * np.abs(exact[i]).max()
        assert len(set(pos[i].tolist())) == k
        assert (exact[i, pos[i]] >= kth - tol).all()
        must = set(np.nonzero(exact[i] > kth + tol)[0].tolist())
        assert must <= set(pos[i].tolist())
        assert np.abs(scores[i] - exact[i, pos[i]]).max() <= tol + rel * np.abs(exact[i]).max()
        assert (np.diff(scores[i]) <= 1e-6).all()


def test_cases_exist():
    assert len(RERANK_CASES) >= 4


@pytest.mark.parametrize("name", RERANK_CASES)
def test_oracle_matches_reference_golden(name):
    g = _load(name)
    k = int(g["topk"])
    s, pos, emb, mrr, mrr_rev = O.rerank_tail(g["q"], g["p"], k)
    if str(g["dtype"]) == "float32":
        assert np.array_equal(pos.numpy(), g["positions"])
    _check_against_exact(pos.numpy(), s.float().numpy(), g)
    assert np.allclose(s.float().numpy(), g["scores"], rtol=1e-6, atol=1e-7)
    if "emb" in g:
        ref_pos = torch.from_numpy(g["positions"])
        want = torch.gather(g["p"], 1, ref_pos.unsqueeze(2).expand(-1, -1, g["p"].shape[-1])).float().numpy()
        assert np.array_equal(want, g["emb"])                      # what the reference gathered = cand[pos]
    assert mrr == pytest.approx(float(g["mrr"]), rel=1e-6 if str(g["dtype"]) == "float32" else 0.2)
    assert mrr_rev == pytest.approx(float(g["mrr_rev"]), rel=1e-6 if str(g["dtype"]) == "float32" else 0.2)


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not mounted")
def test_oracle_matches_live_reference():
    torch.manual_seed(5)
    q, p = torch.randn(4, 48), torch.randn(4, 33, 48)
    out_p, out_s, _, emb, stats = ref_import.run_reference_rerank(q, p, 9)
    s, pos, e2, mrr, mrr_rev = O.rerank_tail(q, p, 9)
    assert [[d["id"] - i * 33 for d in row] for i, row in enumerate(out_p)] == pos.tolist()
    assert np.allclose(np.array(out_s), s.numpy(), rtol=1e-6)
    assert torch.equal(emb, e2)
    assert stats["MRR"][0] == pytest.approx(mrr) and stats["MRR_rev"][0] == pytest.approx(mrr_rev)


def test_rerank_symbols_exported(eng):
    lib = eng._native.load()
    assert lib.mips_max_rerank_candidates() == 1024
    # argument validation happens before any CUDA call, so it can be exercised without a device
    assert lib.mips_rerank(0, None, 8, None, 2, 1, 4, 8, 5, None, None, None, None, None) == eng._native.MIPS_EKRANGE
    assert lib.mips_rerank(0, None, 8, None, 2, 1, 2000, 8, 5, None, None, None, None, None) == eng._native.MIPS_EINVAL
    assert lib.mips_rerank(0, None, 8, None, 7, 1, 4, 8, 2, None, None, None, None, None) == eng._native.MIPS_EINVAL
    assert lib.mips_rerank(0, None, 8, None, 2, 0, 4, 8, 2, None, None, None, None, None) == eng._native.MIPS_OK   # empty batch


def test_rerank_has_no_cpu_fallback(eng):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        eng.rerank_topk(torch.randn(2, 8), torch.randn(2, 3, 8), 2)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", RERANK_CASES)
def test_gpu_rerank_matches_golden(eng, name):
    g = _load(name)
    k = int(g["topk"])
    q, p = g["q"].cuda(), g["p"].cuda()
    s, pos, rank, emb = eng.rerank_topk(q, p, k, want_rank=True)
    pos_h, s_h = pos.cpu().numpy(), s.cpu().numpy()
    _check_against_exact(pos_h, s_h, g)
    if str(g["dtype"]) == "float32":
        assert np.array_equal(pos_h, g["positions"])                     # unmodified reference, exact order
        assert np.allclose(s_h, g["scores"], rtol=2e-5, atol=1e-6)
    # gathered embeddings are bit-exact copies of the winners
    want = torch.gather(p, 1, pos.unsqueeze(2).expand(-1, -1, p.shape[-1]))
    assert torch.equal(emb, want)
    if "emb" in g and str(g["dtype"]) == "float32":
        assert np.array_equal(emb.float().cpu().numpy(), g["emb"])
    # rank is the inverse permutation of the full order; its first k entries are the returned positions
    L = p.shape[1]
    r = rank.cpu().numpy()
    assert all(sorted(r[i].tolist()) == list(range(L)) for i in range(r.shape[0]))
    assert all((r[i, pos_h[i]] == np.arange(k)).all() for i in range(r.shape[0]))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(1, 1, 8, 1), (3, 7, 20, 7), (64, 100, 768, 20), (5, 128, 1024, 128),
                                   (2, 1000, 72, 50), (4, 1024, 64, 1024), (9, 33, 770, 5)])
def test_gpu_rerank_vs_oracle(eng, dtype, shape):
    b, L, d, k = shape
    gen = torch.Generator().manual_seed(b * 1000 + L)
    q = torch.nn.functional.normalize(torch.randn(b, d, generator=gen), dim=-1).to(dtype)
    p = torch.nn.functional.normalize(torch.randn(b, L, d, generator=gen), dim=-1).to(dtype)
    s, pos, _, emb = eng.rerank_topk(q.cuda(), p.cuda(), k)
    g = {"q": q, "p": p, "dtype": str(dtype).replace("torch.", "")}
    _check_against_exact(pos.cpu().numpy(), s.cpu().numpy(), g)
    so, po, eo, _, _ = O.rerank_tail(q.float(), p.float(), k)             # oracle on the same (rounded) inputs
    same = (po == pos.cpu()).float().mean().item()
    assert same > 0.98                                                      # fp32 summation order only
    assert torch.equal(emb.cpu(), torch.gather(p, 1, pos.cpu().unsqueeze(2).expand(-1, -1, d)))


@pytest.mark.gpu
def test_gpu_rerank_ties_and_strides(eng):
    # duplicate candidates: equal scores come back in ascending position order (a total order, unlike torch.sort)
    q = torch.randn(2, 64).cuda()
    base = torch.randn(2, 4, 64)
    p = base.repeat(1, 3, 1).cuda()                                         # positions j, j+4, j+8 identical
    s, pos, _, _ = eng.rerank_topk(q, p, 12)
    pos_h = pos.cpu().numpy()
    for i in range(2):
        for t in range(0, 12, 3):
            trio = pos_h[i, t:t + 3]
            assert (np.diff(trio) == 4).all() and trio[0] < 4
    # strided query rows (a column slice of a wider tensor) are consumed in place
    wide = torch.randn(2, 200).cuda()
    s2, pos2, _, _ = eng.rerank_topk(wide[:, 10:74], p, 12)
    s3, pos3, _, _ = eng.rerank_topk(wide[:, 10:74].contiguous(), p, 12)
    assert torch.equal(pos2, pos3) and torch.equal(s2, s3)


@pytest.mark.gpu
def test_gpu_rerank_errors_and_wrapper(eng):
    q, p = torch.randn(3, 32).cuda(), torch.randn(3, 10, 32).cuda()
    with pytest.raises(RuntimeError, match="selected index k out of range"):
        eng.rerank_topk(q, p, 11)
    with pytest.raises(ValueError):
        eng.rerank_topk(q, torch.randn(3, 2000, 32).cuda(), 5)
    passages = [[{"id": i * 10 + j} for j in range(10)] for i in range(3)]
    stats = {}
    out_p, out_s, emb = eng.rerank_passages(q, p.view(30, 32), passages, 4, iter_stats=stats)
    so, po, eo, mrr, mrr_rev = O.rerank_tail(q.cpu(), p.cpu(), 4)
    assert [[d["id"] - i * 10 for d in row] for i, row in enumerate(out_p)] == po.tolist()
    assert np.allclose(np.array(out_s), so.numpy(), rtol=1e-5, atol=1e-6)
    assert torch.equal(emb.cpu(), eo)
    assert stats["MRR"][0] == pytest.approx(mrr) and stats["MRR_rev"][0] == pytest.approx(mrr_rev)
    assert stats["MRR"][1] == 3
    assert eng.rerank_passages(q[:0], p[:0], [], 4)[:2] == ([], [])
