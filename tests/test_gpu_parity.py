"""Parity of the sm_100a path (called through the C ABI) against the oracle and the golden vectors
frozen from the unmodified reference.  Bar (BASELINE.json north_star): identical top-k passage-id
sets except documented near-ties (score gap below 1e-3 relative), scores within 1e-3 relative —
`oracle.flat_index_oracle.compare_topk` implements exactly that."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from oracle import flat_index_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-3  # tolerance stated by north_star


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    torch.cuda.set_device(0)
    return torch.device("cuda:0")


def _engine(eng, e, dtype=torch.float16, **kw):
    m = eng.MipsEngine(e.shape[1], dtype, e.device)
    m.bind(e, **kw)
    return m


def _synth(n, d, b, seed, dev, dtype=torch.float16):
    g = torch.Generator(device=dev).manual_seed(seed)
    e = torch.empty(n, d, dtype=dtype, device=dev)
    for s in range(0, n, 1 << 20):
        c = torch.randn(min(1 << 20, n - s), d, generator=g, device=dev)
        e[s:s + c.shape[0]] = torch.nn.functional.normalize(c, dim=1).to(dtype)
    q = torch.nn.functional.normalize(torch.randn(b, d, generator=g, device=dev), dim=1)
    return e, q


def _torch_ref(e, q, k, dtype=torch.float16):
    """fp32 scores of the stored operands on the same device + (score desc, id asc) order."""
    s = q.to(dtype).float() @ e.float().T
    order = torch.argsort(-s, dim=1, stable=True)[:, :k]
    return torch.gather(s, 1, order), order


# ------------------------------------------------------------------ golden vectors (reference outputs)
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_vectors_through_cabi(eng, dev, name):
    g = load_golden(name)
    k = int(g["k"])
    e = torch.from_numpy(g["embeddings"]).to(dev)
    m = _engine(eng, e)
    s, i = m.search(torch.from_numpy(g["queries"]).to(dev), k)
    exact = O.exact_scores(g["queries"], g["embeddings"])
    rep = O.compare_topk(i.cpu().numpy(), s.cpu().numpy(), g["ids"], g["scores"].astype(np.float32), exact, rtol=RTOL)
    assert rep["ok"], rep["errors"][:3]
    # same through the HOST-buffer entry point of the C ABI
    s2, i2 = m.search_host(torch.from_numpy(g["queries"]), k)
    assert torch.equal(s2, s.cpu()) and torch.equal(i2, i.cpu())


def test_drop_in_search_knn_matches_reference_outputs(eng, dev):
    g = load_golden("flat_n1003_d768_b8_k20")
    n = int(g["n"])
    passages = [{"id": str(i), "title": f"t{i}", "text": f"passage {i}"} for i in range(n)]
    idx = eng.B200Index()
    idx.init_embeddings(passages, dim=768)
    assert idx.embeddings.is_cuda and tuple(idx.embeddings.shape) == (768, n)
    for a in range(0, n, 256):                                   # write site of src/rag.py:108-121
        chunk = torch.from_numpy(g["embeddings"][a:a + 256]).to(dev)
        idx.embeddings[:, a:a + chunk.shape[0]] = chunk.T
    iter_stats = {}
    idx.iter_stats = iter_stats                                                  # src/rag.py:143: the trainer's dict
    docs, scores = idx.search_knn(torch.from_numpy(g["queries"]).to(dev), 20)   # docs first (src/index.py:158)
    # runtime/search (src/rag.py:170) is CUDA-synchronised here and split into device time and host tail
    assert set(iter_stats) == {"runtime/search", "runtime/search_device", "runtime/search_host_tail"}
    assert iter_stats["runtime/search"][0] >= iter_stats["runtime/search_device"][0] > 0 and iter_stats["runtime/search"][1] == 1
    assert len(docs) == 8 and len(docs[0]) == 20 and isinstance(docs[0][0], dict) and isinstance(scores[0][0], float)
    ids = np.array([[int(d["id"]) for d in row] for row in docs])
    exact = O.exact_scores(g["queries"], g["embeddings"])
    rep = O.compare_topk(ids, np.array(scores), g["ids"], g["scores"].astype(np.float32), exact, rtol=RTOL)
    assert rep["ok"], rep["errors"][:3]
    # scores are returned fp16-rounded like the reference's: most are bit-identical to the golden ones
    same = np.mean(np.sort(np.array(scores, dtype=np.float32), axis=1) == np.sort(g["scores"].astype(np.float32), axis=1))
    assert same > 0.9
    assert idx.search_knn(torch.empty(0, 768, device=dev), 20) == ([], [])          # empty batch
    with pytest.raises(RuntimeError, match="selected index k out of range"):        # torch.topk's error
        idx.search_knn(torch.from_numpy(g["queries"]).to(dev), n + 1)
    d3, s3, emb = idx.search_knn(torch.from_numpy(g["queries"]).to(dev), 5, return_embeddings=True)
    ids3 = torch.tensor([[int(d["id"]) for d in row] for row in d3])
    assert tuple(emb.shape) == (8, 5, 768) and torch.equal(emb.cpu(), torch.from_numpy(g["embeddings"])[ids3])
    tw = eng.B200IndexWithEmbeddings()
    tw.init_embeddings(passages, dim=768)
    tw.embeddings[:, :] = torch.from_numpy(g["embeddings"]).to(dev).T
    assert len(tw.search_knn(torch.from_numpy(g["queries"]).to(dev), 5)) == 3       # build_server/index.py:261


def test_in_place_refresh_needs_no_rebind(eng, dev):
    """The index refresh of the reference rewrites `index.embeddings[:, a:b]` in place (src/rag.py:108-121,
    train.py:189-204); the next search must see the new values without any extra call."""
    e, q = _synth(20000, 768, 8, 41, dev)
    idx = eng.B200Index()
    idx.init_embeddings([{"id": str(i)} for i in range(20000)], dim=768)
    idx.embeddings[:, :] = e.T
    _, i0 = idx.search(q, 10)
    e2, _ = _synth(20000, 768, 8, 42, dev)
    for a in range(0, 20000, 4096):
        idx.embeddings[:, a:a + 4096] = e2[a:a + 4096].T
    _, i1 = idx.search(q, 10)
    assert torch.equal(i1, _torch_ref(e2, q, 10)[1]) and not torch.equal(i0, i1)


def test_save_load_roundtrip_on_device(eng, dev, tmp_path):
    g = load_golden("flat_n300_d1024_b5_k10")
    passages = [{"id": str(i)} for i in range(300)]
    idx = eng.B200Index()
    idx.init_embeddings(passages, dim=1024)
    idx.embeddings[:, :] = torch.from_numpy(g["embeddings"]).to(dev).T
    idx.save_index(str(tmp_path), 3)
    idx2 = eng.B200Index()
    idx2.load_index(str(tmp_path), 3)
    d1, s1 = idx.search_knn(torch.from_numpy(g["queries"]).to(dev), 10)
    d2, s2 = idx2.search_knn(torch.from_numpy(g["queries"]).to(dev), 10)
    assert d1 == d2 and s1 == s2
    # shards stream to the device in column chunks (here 37 columns at a time): same matrix
    import importlib
    mod = importlib.import_module("jsa-rag_b200.index")
    saved, mod._LOAD_CHUNK = mod._LOAD_CHUNK, 37
    try:
        idx3 = eng.B200Index()
        idx3.load_index(str(tmp_path), 3)
    finally:
        mod._LOAD_CHUNK = saved
    assert torch.equal(idx3.embeddings, idx.embeddings) and idx3.doc_map == idx.doc_map


# ------------------------------------------------------------------ oracle on seeded inputs, edge cases
@pytest.mark.parametrize("n,d,b,k,dtype", [
    (1, 768, 1, 1, torch.float16), (127, 768, 3, 127, torch.float16), (128, 768, 64, 128, torch.float16),
    (129, 768, 65, 20, torch.float16), (5000, 1024, 7, 100, torch.float16), (30011, 768, 130, 10, torch.float16),
    (60000, 768, 64, 100, torch.bfloat16), (148 * 128 * 5 + 77, 768, 33, 100, torch.float16),
    (250000, 768, 64, 100, torch.float16), (20000, 64, 16, 50, torch.float16), (20000, 256, 5, 128, torch.bfloat16),
])
def test_matches_fp32_oracle_bitwise_ids(eng, dev, n, d, b, k, dtype):
    """fp32-accumulated scores of identical stored operands: ids must agree with the (score desc,
    id asc) order except where two fp32 scores differ by less than accumulation-order noise."""
    e, q = _synth(n, d, b, 1000 + n % 97, dev, dtype)
    m = _engine(eng, e, dtype)
    s, i = m.search(q, k)
    rs, ri = _torch_ref(e, q, k, dtype)
    assert bool((s[:, 1:] <= s[:, :-1]).all()), "scores not sorted descending"
    assert int(i.min()) >= 0 and int(i.max()) < n
    got = torch.gather(q.to(dtype).float() @ e.float().T, 1, i)
    assert float((s - got).abs().max()) <= 2e-6 * max(1.0, float(got.abs().max())) + 1e-6
    exact = (q.to(dtype).double() @ e.double().T).cpu().numpy()
    rep = O.compare_topk(i.cpu().numpy(), s.cpu().numpy(), ri.cpu().numpy(), rs.cpu().numpy(), exact, rtol=1e-5, atol=1e-6)
    assert rep["ok"], rep["errors"][:3]
    assert rep["near_tie_diffs"] <= max(1, b * k // 200)


@pytest.mark.parametrize("n,b,k", [(3000, 5, 1000), (200_000, 64, 1000), (50_000, 130, 129), (100_000, 9, 512),
                                   (1_500_000, 64, 1024), (1024, 3, 1024)])
def test_large_k(eng, dev, n, b, k):
    """k in (128, 1024] (BASELINE configs[4] sweeps k = 1000): 2048-slot lists + streaming select."""
    e, q = _synth(n, 768, b, 77 + k, dev)
    m = _engine(eng, e)
    s, i = m.search(q, k)
    rs, ri = _torch_ref(e, q, k)
    assert bool((s[:, 1:] <= s[:, :-1]).all())
    exact = (q.half().double() @ e.double().T).cpu().numpy()
    rep = O.compare_topk(i.cpu().numpy(), s.cpu().numpy(), ri.cpu().numpy(), rs.cpu().numpy(), exact, rtol=1e-5, atol=1e-6)
    assert rep["ok"], rep["errors"][:3]
    assert rep["near_tie_diffs"] <= max(2, b * k // 200)
    # ties at large k: duplicated rows, (score desc, id asc) order must be exact
    if n <= 50_000:
        e2 = e[: n // 4].repeat(4, 1)
        m.bind(e2)
        s2, i2 = m.search(q[:3], k)
        rs2, ri2 = _torch_ref(e2, q[:3], k)
        assert torch.equal(i2, ri2)
    # cross-rank merge at large k
    W = 4
    parts = [_engine(eng, e[r::W].contiguous(), id_base=r, id_stride=W).search(q, k) for r in range(W)] if n >= W * k else None
    if parts:
        ms, mi = eng.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), k)
        assert torch.equal(mi, i) and torch.equal(ms, s)


def test_ties_are_broken_by_ascending_id(eng, dev):
    z = torch.zeros(5000, 768, dtype=torch.float16, device=dev)          # init_embeddings() state: all scores 0
    m = _engine(eng, z)
    s, i = m.search(torch.randn(7, 768, device=dev), 20)
    assert bool((s == 0).all()) and bool((i == torch.arange(20, device=dev)).all())
    e, q = _synth(3000, 768, 9, 5, dev)
    e2 = e.repeat(40, 1)                                                  # every row 40 times
    m.bind(e2)
    s, i = m.search(q, 100)
    rs, ri = _torch_ref(e2, q, 100)
    assert torch.equal(i, ri) and float((s - rs).abs().max()) < 1e-5


def test_global_id_mapping_and_gather(eng, dev):
    e, q = _synth(4096, 768, 4, 3, dev)
    m = _engine(eng, e, id_base=3, id_stride=8)      # round-robin shard of rank 3 of 8 (src/index_io.py:41)
    s, i = m.search(q, 10)
    rs, ri = _torch_ref(e, q, 10)
    assert torch.equal(i, ri * 8 + 3)
    rows = m.gather_rows(ri)
    assert torch.equal(rows.view(4, 10, 768), e[ri])
    m.bind(e[:, :], id_base=1000, id_stride=1)       # contiguous shard starting at row 1000
    assert torch.equal(m.search(q, 10)[1], ri + 1000)


def test_query_dtypes_and_normalisation(eng, dev):
    """allqueries.half() (src/index.py:118) for fp32 / bf16 / fp16 queries; normalize=True is
    faiss.normalize_L2 on the queries only (build_server/server_start.py:142)."""
    e, q = _synth(20000, 768, 12, 11, dev)
    m = _engine(eng, e)
    s32, i32 = m.search(q, 30)
    s16, i16 = m.search(q.half(), 30)
    assert torch.equal(i32, i16) and torch.equal(s32, s16)
    sb, ib = m.search(q.bfloat16(), 30)
    rs, ri = _torch_ref(e, q.bfloat16().float(), 30)
    assert float((sb - rs).abs().max()) < 1e-5
    scaled = q * torch.linspace(0.1, 30, 12, device=dev)[:, None]
    sn, inn = m.search(scaled, 30, normalize=True)
    d, i = O.server_search(scaled.cpu().numpy(), e.cpu().numpy(), 30)
    exact = O.exact_scores(O.normalize_l2(scaled.cpu().numpy()), e.cpu().numpy(), q_dtype=None)
    rep = O.compare_topk(inn.cpu().numpy(), sn.cpu().numpy(), i, d, exact, rtol=RTOL)
    assert rep["ok"], rep["errors"][:3]
    # non-contiguous (row-strided) queries
    wide = torch.zeros(12, 1024, device=dev)
    wide[:, :768] = q
    s_str, i_str = m.search(wide[:, :768], 30)
    assert torch.equal(i_str, i32)


def test_error_codes_and_limits(eng, dev):
    e, q = _synth(500, 768, 2, 1, dev)
    m = _engine(eng, e)
    with pytest.raises(RuntimeError, match="selected index k out of range"):
        m.search(q, 501)
    big, _ = _synth(2000, 768, 1, 2, dev)
    mb = _engine(eng, big)
    with pytest.raises(ValueError):
        mb.search(q, eng._native.load().mips_max_k() + 1)
    with pytest.raises(ValueError):
        m.search(q[:, :100], 5)
    with pytest.raises(ValueError):
        m.bind(e.float())
    s, i = m.search(q[:0], 5)
    assert tuple(s.shape) == (0, 5)
    fresh = eng.MipsEngine(768, torch.float16, dev)
    with pytest.raises(RuntimeError, match="mips_bind_index"):
        fresh.search(q, 5)
    assert m.last_launch_count() == 0 and m.search(q, 5) and m.last_launch_count() >= 3


def test_merge_kernel_matches_reference_merge(eng, dev):
    """mips_merge_topk == concat per-rank blocks in rank order + torch.topk (src/index.py:143-157),
    with (score desc, id asc) tie order."""
    torch.manual_seed(0)
    for L, b, k_in, k_out in [(8, 64, 100, 100), (2, 5, 20, 20), (4, 3, 128, 50), (1, 9, 10, 10), (16, 2, 7, 7)]:
        s = torch.sort(torch.randn(L, b, k_in, device=dev), dim=2, descending=True).values
        s[0, :, : min(3, k_in)] = s[L - 1, :, : min(3, k_in)]              # exact cross-list ties
        s = torch.sort(s, dim=2, descending=True).values
        ids = torch.stack([torch.randperm(100000, device=dev)[: b * k_in].view(b, k_in) + 100000 * l for l in range(L)])
        ms, mi = eng.merge_topk(s, ids, k_out)
        ref_s, ref_i = O.merge_rank_results(list(s.cpu()), list(ids.cpu()), k_out)
        assert torch.equal(ms.cpu(), ref_s)
        flat_s = s.permute(1, 0, 2).reshape(b, -1).cpu().double().numpy()
        flat_i = ids.permute(1, 0, 2).reshape(b, -1).cpu().numpy()
        want = np.stack([flat_i[r][np.lexsort((flat_i[r], -flat_s[r]))[:k_out]] for r in range(b)])
        assert np.array_equal(mi.cpu().numpy(), want)
    # padding entries (id < 0) are ignored
    s = torch.tensor([[[3.0, 1.0, -float("inf")]], [[2.0, -float("inf"), -float("inf")]]], device=dev)
    i = torch.tensor([[[5, 7, -1]], [[9, -1, -1]]], device=dev)
    ms, mi = eng.merge_topk(s, i, 3)
    assert mi.tolist() == [[5, 9, 7]] and ms.tolist() == [[3.0, 2.0, 1.0]]


def test_emulated_shards_merge_to_the_global_answer(eng, dev):
    """Row-sharding property (the 8-GPU layout on one device): merging the per-shard top-k of W
    round-robin shards equals the top-k of the whole index, ids included."""
    e, q = _synth(80000, 768, 64, 21, dev)
    m = _engine(eng, e)
    gs, gi = m.search(q, 100)
    W = 8
    parts = []
    for r in range(W):
        mr = _engine(eng, e[r::W].contiguous(), id_base=r, id_stride=W)
        parts.append(mr.search(q, 100))
    ms, mi = eng.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), 100)
    assert torch.equal(mi, gi) and torch.equal(ms, gs)


# ------------------------------------------------------------------ size-independent properties at scale
def _properties(eng, dev, n, b, k, seed):
    e, q = _synth(n, 768, b, seed, dev)
    planted = torch.randint(0, n, (b,), device=dev, generator=torch.Generator(device=dev).manual_seed(seed + 1))
    q = torch.nn.functional.normalize(e[planted].float() + 0.02 * q, dim=1)      # known nearest neighbours
    m = _engine(eng, e)
    s, i = m.search(q, k)
    s2, i2 = m.search(q, k)
    assert torch.equal(s, s2) and torch.equal(i, i2), "search is not deterministic / idempotent"
    assert torch.equal(i[:, 0], planted), "planted nearest neighbour not ranked first"
    assert bool((s[:, 1:] <= s[:, :-1]).all())
    assert all(len(set(r.tolist())) == k for r in i.cpu()), "duplicate ids"
    # returned scores are the true inner products of the returned rows
    got = torch.einsum("bd,bkd->bk", q.half().float(), e[i].float())
    assert float((s - got).abs().max()) <= 2e-5 * max(1.0, float(got.abs().max()))   # fp32 summation order only
    # no row of a random 2M-row sample beats the k-th returned score (completeness of the scan)
    samp = torch.randint(0, n, (min(n, 2_000_000),), device=dev)
    for a in range(0, samp.numel(), 1 << 19):
        rows = samp[a:a + (1 << 19)]
        sc = q.half().float() @ e[rows].float().T
        better = sc > (s[:, -1:] * (1 + 1e-6) + 1e-7)
        if better.any():
            bq, bc = better.nonzero(as_tuple=True)
            assert bool((i[bq] == rows[bc][:, None]).any(1).all()), "a better row is missing from the result"
    return m, e, q


def test_properties_at_shard_size(eng, dev):
    _properties(eng, dev, 4_125_000, 64, 100, 7)          # one of 8 shards of the 33M index


@pytest.mark.skipif(os.environ.get("JSA_SKIP_FULL_SIZE") == "1", reason="disabled by JSA_SKIP_FULL_SIZE")
def test_properties_at_full_size(eng, dev):
    """BASELINE configs[1]: 33M x 768 fp16 (50.7 GB), batch 64, top-100 on one B200."""
    free, total = torch.cuda.mem_get_info()
    if free < 70 * (1 << 30):
        pytest.skip("not enough free device memory for the 33M x 768 index")
    _properties(eng, dev, 33_000_000, 64, 100, 9)


def test_raw_c_abi_with_caller_workspace(eng, dev):
    """The boundary exactly as INTEGRATION.md binds it: plain pointers, caller-owned workspace."""
    import ctypes
    lib = eng._native.load()
    e, q = _synth(30000, 768, 70, 4, dev)
    h = ctypes.c_void_p()
    assert lib.mips_create(ctypes.byref(h), 0, 768, 0) == 0
    try:
        assert lib.mips_bind_index(h, ctypes.c_void_p(e.data_ptr()), e.shape[0], e.stride(0), 5, 3) == 0
        need = ctypes.c_size_t(0)
        assert lib.mips_workspace_bytes(h, 70, 50, ctypes.byref(need)) == 0 and need.value > 0
        ws = torch.empty(need.value + 1024, dtype=torch.uint8, device=dev)
        off = (-ws.data_ptr()) % 1024
        s = torch.empty(70, 50, device=dev)
        i = torch.empty(70, 50, dtype=torch.int64, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        args = (h, ctypes.c_void_p(q.data_ptr()), 2, 768, 70, 50, 0, ctypes.c_void_p(s.data_ptr()), ctypes.c_void_p(i.data_ptr()))
        assert lib.mips_search_local(*args, ctypes.c_void_p(ws.data_ptr() + off), need.value, st) == 0
        rs, ri = _torch_ref(e, q, 50)
        torch.cuda.synchronize()
        exact = (q.half().double() @ e.double().T).cpu().numpy()
        assert bool(((i - 5) % 3 == 0).all())
        rep = O.compare_topk(((i - 5) // 3).cpu().numpy(), s.cpu().numpy(), ri.cpu().numpy(), rs.cpu().numpy(), exact,
                             rtol=1e-5, atol=1e-6)
        assert rep["ok"] and rep["near_tie_diffs"] <= 3, rep["errors"][:3]
        assert lib.mips_last_launch_count(h) >= 3
        # too small a workspace is refused, nothing is launched
        assert lib.mips_search_local(*args, ctypes.c_void_p(ws.data_ptr() + off), 4096, st) == eng._native.MIPS_EWORKSPACE
        assert "workspace" in eng._native.last_error(h)
    finally:
        lib.mips_destroy(h)


@pytest.mark.parametrize("n,d,b,k,dtype", [(5000, 768, 9, 20, torch.float16), (70001, 768, 64, 100, torch.float16),
                                           (30000, 1024, 70, 50, torch.bfloat16), (130, 64, 3, 5, torch.float16)])
def test_reference_layout_is_consumed_in_place(eng, dev, n, d, b, k, dtype):
    """A [dim, n_local] tensor in the reference's own layout (src/index.py:52) is searched zero-copy as an
    MN-major tcgen05 operand; results are bit-identical to the K-major copy."""
    e, q = _synth(n, d, b, 31, dev, dtype)
    n_pad = (n + 7) // 8 * 8                       # TMA needs 16-byte aligned row pitch
    e_dn = torch.zeros(d, n_pad, dtype=dtype, device=dev)
    e_dn[:, :n] = e.T
    m_k = _engine(eng, e, dtype)
    m_mn = eng.MipsEngine(d, dtype, dev)
    m_mn.bind(e_dn[:, :n].t())                     # [n, d] view with unit stride along n
    s1, i1 = m_k.search(q, k)
    s2, i2 = m_mn.search(q, k)
    assert torch.equal(i1, i2) and torch.equal(s1, s2)
    assert torch.equal(m_mn.gather_rows(i2[:, :2]), e[i2[:, :2].reshape(-1)])
    if n % 8 == 0 or n == 5000:
        idx = eng.B200Index(dtype=dtype, layout="dn")            # drop-in object storing the reference layout
        idx.init_embeddings([{"id": str(j)} for j in range(n)], dim=d)
        idx.embeddings[:, :] = e.T
        assert idx.embeddings.is_contiguous()
        s3, i3 = idx.search(q, k)
        assert torch.equal(i3, i1) and torch.equal(s3, s1)


def test_cuda_graph_replay_of_a_search(eng, dev):
    e, q = _synth(300_000, 768, 64, 17, dev)
    idx = eng.B200Index()
    idx._store = e
    idx._set_sharding("round_robin")
    run = idx.make_graphed_search(64, 100)
    s0, i0 = idx.search(q, 100)
    s1, i1 = run(q)
    torch.cuda.synchronize()
    assert torch.equal(s0, s1) and torch.equal(i0, i1)
    q2 = torch.roll(q, 1, 0)
    s2, i2 = run(q2)
    torch.cuda.synchronize()
    assert torch.equal(i2, torch.roll(i0, 1, 0))


@pytest.mark.parametrize("n,b,k", [(150_000, 300, 100), (150_000, 512, 20), (64, 260, 10), (1_000_000, 1024, 100), (40_000, 129, 500)])
def test_multi_block_launches_match_single_block(eng, dev, n, b, k):
    """Batches > 128 run on tcgen05 CTA pairs (cta_group::2, UMMA M = 256: 1, 2 or 4 pair blocks per launch, each
    CTA loading half of every passage tile); flag 128 = the round-1 path (2 or 4 single-CTA query blocks per launch
    sharing tiles through the L2).  Both must be bit-identical to one block per launch (debug flag 32) and agree
    with the oracle."""
    e, q = _synth(n, 768, b, 5 + b, dev)
    m = _engine(eng, e)
    s, i = m.search(q, k)
    m.debug_config(32, False)
    s1, i1 = m.search(q, k)
    m.debug_config(128, False)
    s2, i2 = m.search(q, k)
    m.debug_config(0, False)
    assert torch.equal(i, i1) and torch.equal(s, s1)
    assert torch.equal(i2, i1) and torch.equal(s2, s1)
    rs, ri = _torch_ref(e, q, k)
    exact = (q.half().double() @ e.double().T).cpu().numpy()
    rep = O.compare_topk(i.cpu().numpy(), s.cpu().numpy(), ri.cpu().numpy(), rs.cpu().numpy(), exact, rtol=1e-5, atol=1e-6)
    assert rep["ok"], rep["errors"][:3]


@pytest.mark.parametrize("n,b,k", [(200_000, 64, 100), (200_000, 100, 17), (1_200_000, 20, 600), (300_000, 130, 128)])
def test_in_stream_compaction_paths(eng, dev, n, b, k):
    """Without the sampled pre-passes (debug flag 4) every list overflows repeatedly, so the in-stream and
    final compactions (register-resident for k <= 128, streaming for larger k) carry the search."""
    e, q = _synth(n, 768, b, 99 + k, dev)
    m = _engine(eng, e)
    s0, i0 = m.search(q, k)
    m.debug_config(4, False)
    s1, i1 = m.search(q, k)
    m.debug_config(0, False)
    assert torch.equal(i0, i1) and torch.equal(s0, s1), "seeded and unseeded searches disagree"
    rs, ri = _torch_ref(e, q, k)
    exact = (q.half().double() @ e.double().T).cpu().numpy()
    rep = O.compare_topk(i1.cpu().numpy(), s1.cpu().numpy(), ri.cpu().numpy(), rs.cpu().numpy(), exact, rtol=1e-5, atol=1e-6)
    assert rep["ok"], rep["errors"][:3]


@pytest.mark.parametrize("n,d,b,k,dtype", [
    (300_000, 768, 64, 100, torch.float16), (300_000, 768, 17, 128, torch.float16), (1_500_000, 768, 64, 100, torch.float16),
    (400_000, 768, 300, 100, torch.bfloat16), (400_000, 768, 512, 37, torch.float16), (250_000, 1024, 100, 10, torch.float16),
    (2_000_000, 256, 130, 1, torch.float16), (148 * 64 * 5 + 1, 768, 64, 100, torch.float16),
])
def test_in_kernel_seeding_matches_host_prepass(eng, dev, n, d, b, k, dtype):
    """k <= 128: the sampled pre-pass runs inside the full-shard scan (top-4 per CTA and query over its first
    tiles, token-tagged flags, one warp per query takes the k-th best).  Results must be bit-identical to the
    separate sampled scan + select launches (debug flag 64) and to the unseeded scan (flag 4), with fewer launches."""
    e, q = _synth(n, d, b, 7 + k, dev, dtype)
    m = _engine(eng, e, dtype)
    s0, i0 = m.search(q, k)
    launches = m.last_launch_count()
    for _ in range(3):                      # the flag words are reused by every launch (fresh token each time)
        s0b, i0b = m.search(q, k)
        assert torch.equal(i0, i0b) and torch.equal(s0, s0b)
    m.debug_config(64, False)
    s1, i1 = m.search(q, k)
    launches_host = m.last_launch_count()
    m.debug_config(4, False)
    s2, i2 = m.search(q, k)
    m.debug_config(0, False)
    assert torch.equal(i0, i1) and torch.equal(s0, s1), "in-kernel and host-side seeding disagree"
    assert torch.equal(i0, i2) and torch.equal(s0, s2), "seeded and unseeded searches disagree"
    assert launches <= launches_host and (b > 128 or launches == 3)
    rs, ri = _torch_ref(e, q, k, dtype)
    exact = (q.to(dtype).double() @ e.double().T).cpu().numpy()
    rep = O.compare_topk(i0.cpu().numpy(), s0.cpu().numpy(), ri.cpu().numpy(), rs.cpu().numpy(), exact, rtol=1e-5, atol=1e-6)
    assert rep["ok"], rep["errors"][:3]


def test_async_host_search_keeps_requests_in_flight(eng, dev):
    """mips_search_host_async only enqueues: several requests in flight on one stream, answers valid after the
    event recorded behind each of them, identical to the synchronous call."""
    e, q = _synth(200_000, 768, 40, 81, dev)
    m = _engine(eng, e)
    qs = [torch.roll(q, r, 0).cpu().pin_memory() for r in range(4)]
    want = [m.search_host(x, 20) for x in qs]
    outs = [(torch.empty(40, 20).pin_memory(), torch.empty(40, 20, dtype=torch.int64).pin_memory()) for _ in qs]
    evs = []
    for x, o in zip(qs, outs):
        m.search_host(x, 20, out=o, wait=False)
        ev = torch.cuda.Event(); ev.record(); evs.append(ev)
    for ev, o, w in zip(evs, outs, want):
        ev.synchronize()
        assert torch.equal(o[1], w[1]) and torch.equal(o[0], w[0])


def test_concurrent_searches_on_two_streams(eng, dev):
    """Two engines searching at the same time on two streams share the SMs, so neither scan has all of its CTAs
    resident: the in-kernel seeding must not depend on that (bounded, token-tagged hand-offs; late CTAs are left
    out and the scan proceeds with a lower seed).  Results stay exact."""
    e1, q1 = _synth(600_000, 768, 64, 71, dev)
    e2, q2 = _synth(500_000, 768, 48, 72, dev)
    m1, m2 = _engine(eng, e1), _engine(eng, e2)
    ref1, ref2 = m1.search(q1, 100), m2.search(q2, 50)
    torch.cuda.synchronize()
    st1, st2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    outs = []
    for _ in range(12):
        with torch.cuda.stream(st1):
            outs.append((0, m1.search(q1, 100)))
        with torch.cuda.stream(st2):
            outs.append((1, m2.search(q2, 50)))
    torch.cuda.synchronize()
    for which, (s, i) in outs:
        rs, ri = (ref1, ref2)[which]
        assert torch.equal(i, ri) and torch.equal(s, rs)


def test_concurrent_pair_scans_on_two_streams(eng, dev):
    """The same with the CTA-pair kernel (clusters of 2, two pair blocks per launch): pairs that share a tile sequence
    keep in lock-step through progress words, and a sibling whose SMs are held by the other stream's kernel must only
    cost a bounded wait (the lock-step is dropped), never a hang or a wrong answer."""
    e1, q1 = _synth(700_000, 768, 512, 81, dev)
    e2, q2 = _synth(600_000, 768, 300, 82, dev)
    m1, m2 = _engine(eng, e1), _engine(eng, e2)
    ref1, ref2 = m1.search(q1, 100), m2.search(q2, 20)
    ref1, ref2 = tuple(t.clone() for t in ref1), tuple(t.clone() for t in ref2)
    torch.cuda.synchronize()
    st1, st2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    outs = []
    for _ in range(6):
        with torch.cuda.stream(st1):
            outs.append((0, m1.search(q1, 100)))
        with torch.cuda.stream(st2):
            outs.append((1, m2.search(q2, 20)))
    torch.cuda.synchronize()
    for which, (s, i) in outs:
        rs, ri = (ref1, ref2)[which]
        assert torch.equal(i, ri) and torch.equal(s, rs)


@pytest.mark.parametrize("k", [100, 700])
def test_adversarial_row_order(eng, dev, k):
    """Scores that grow with the row number defeat the seeded thresholds (the sample is the worst part
    of the index): correctness must then come from the running thresholds and compactions alone."""
    n, b = 400_000, 40
    g = torch.Generator(device=dev).manual_seed(3)
    u = torch.nn.functional.normalize(torch.randn(b, 768, generator=g, device=dev), dim=1)
    base = torch.nn.functional.normalize(torch.randn(n, 768, generator=g, device=dev), dim=1)
    ramp = torch.linspace(0.0, 1.0, n, device=dev)[:, None]
    e = (0.3 * base + ramp * u.mean(0, keepdim=True)).half()          # later rows score higher for every query
    q = u + u.mean(0, keepdim=True)
    m = _engine(eng, e)
    s, i = m.search(q, k)
    rs, ri = _torch_ref(e, q, k)
    exact = (q.half().double() @ e.double().T).cpu().numpy()
    rep = O.compare_topk(i.cpu().numpy(), s.cpu().numpy(), ri.cpu().numpy(), rs.cpu().numpy(), exact, rtol=1e-5, atol=1e-6)
    assert rep["ok"], rep["errors"][:3]
    assert int(i.min()) > n // 4          # the winners sit in the late, high-scoring part of the index


@pytest.mark.parametrize("n,d,b,k,dtype", [
    (90_000, 1024, 512, 100, torch.float16),     # K tail of the queries in shared memory (SS-mode pair MMAs)
    (90_000, 64, 1500, 10, torch.bfloat16),      # one K chunk per stage, 4 + 2 pair blocks
    (31, 768, 256, 31, torch.float16),           # fewer tiles than pairs
    (200_001, 768, 257, 100, torch.bfloat16),    # second CTA pair almost empty
    (120_000, 768, 640, 300, torch.float16),     # big-k lists on pairs
    (700_000, 768, 1024, 16, torch.float16),     # 4 pair blocks, in-kernel seeding (18 lists x 4 >= 4 k)
])
def test_cta_pair_scan_matches_oracle(eng, dev, n, d, b, k, dtype):
    e, q = _synth(n, d, b, 77 + b, dev, dtype)
    m = _engine(eng, e, dtype)
    s, i = m.search(q, k)
    m.debug_config(32, False)
    s1, i1 = m.search(q, k)
    m.debug_config(2048, False)                  # 4 pair blocks per launch (automatic only for very long shards)
    s4, i4 = m.search(q, k)
    m.debug_config(0, False)
    assert torch.equal(i, i1) and torch.equal(s, s1)
    assert torch.equal(i4, i1) and torch.equal(s4, s1)
    rs, ri = _torch_ref(e, q, k, dtype)
    exact = (q.to(dtype).double() @ e.double().T).cpu().numpy()
    rep = O.compare_topk(i.cpu().numpy(), s.cpu().numpy(), ri.cpu().numpy(), rs.cpu().numpy(), exact, rtol=1e-5, atol=1e-6)
    assert rep["ok"], rep["errors"][:3]


def test_graph_survives_a_larger_eager_search(eng, dev):
    """A captured search holds raw pointers into the engine workspace.  A later, larger eager search on the same
    index outgrows that workspace; it must be retired, not freed, until the graph is released."""
    e, q = _synth(200_000, 768, 640, 23, dev)
    idx = eng.B200Index()
    idx._store = e
    idx._set_sharding("round_robin")
    run = idx.make_graphed_search(16, 10)
    want_s, want_i = idx.search(q[:16], 10)
    want_s, want_i = want_s.clone(), want_i.clone()
    big_s, big_i = idx.search(q, 1000)                 # batch 640, k 1000: a much larger workspace
    filler = [torch.full((1 << 22,), 7, dtype=torch.int64, device=dev) for _ in range(8)]   # would reuse freed memory
    s1, i1 = run(q[:16])
    torch.cuda.synchronize()
    assert torch.equal(i1, want_i) and torch.equal(s1, want_s)
    s2, i2 = idx.search(q, 1000)
    assert torch.equal(i2, big_i) and torch.equal(s2, big_s)
    run.release()
    run.release()                                      # idempotent
    s3, i3 = idx.search(q[:16], 10)
    assert torch.equal(i3, want_i)
    del filler


def test_handles_of_different_dims_share_a_device(eng, dev):
    """The dynamic shared-memory limit is a per-function attribute: a dim-768 handle created after a dim-1024 one
    must not lower it."""
    e1, q1 = _synth(20_000, 1024, 9, 3, dev)
    m1 = _engine(eng, e1)
    e2, q2 = _synth(20_000, 768, 9, 4, dev)
    m2 = _engine(eng, e2)
    for m, e, q in ((m1, e1, q1), (m2, e2, q2), (m1, e1, q1)):
        s, i = m.search(q, 10)
        rs, ri = _torch_ref(e, q, 10)
        assert torch.equal(i, ri)


def test_faiss_flat_mode_matches_its_restatement(eng, dev, tmp_path):
    """index_mode="faiss", faiss_index_type="flat" (src/index.py:164-223) runs on the same engine: against
    the restated faiss semantics (fp32 queries over the fp16-stored matrix, scores .half()) the answer stays inside
    north_star's tolerance (1e-3 relative) — the only arithmetic difference is the fp16 rounding of the queries."""
    class Opt:
        index_mode, faiss_index_type, faiss_code_size = "faiss", "flat", None
        retriever_model_path, load_index_path, passages, max_passages = "facebook/contriever", None, [], -1
    index, _ = eng.load_or_initialize_index(Opt())
    n, k = 80_000, 100
    e, q = _synth(n, 768, 48, 3, dev)
    index.init_embeddings([{"id": str(i)} for i in range(n)], dim=768)
    index.embeddings[:, :] = e.T
    docs, scores = index.search_knn(q, k)
    ids = np.array([[int(d["id"]) for d in row] for row in docs])
    e_dn = O.make_embeddings_dn(e.cpu())
    fs, fi = O.faiss_flat_search(q.cpu(), e_dn, k)
    exact = O.exact_scores(q.cpu().numpy(), e.cpu().numpy(), q_dtype=np.float32)
    rep = O.compare_topk(ids, np.array(scores), fi.numpy(), fs.float().numpy(), exact, rtol=RTOL)
    assert rep["ok"], rep["errors"][:3]


def test_randomised_configurations(eng, dev):
    """Seeded random mix of shapes, dtypes, layouts, id mappings and k (interactions between the small /
    big-k modes, M=64 / M=128 / multi-block launches, the smem K tail of dim 1024 and both layouts)."""
    import random
    rnd = random.Random(20260718)
    for trial in range(20):
        d = rnd.choice([64, 256, 768, 768, 1024])
        n = rnd.choice([1, 63, 65, 1000, 9473, 40_000, 150_000, 300_001])
        b = rnd.choice([1, 7, 64, 65, 128, 129, 300, 520, 1100])
        k = min(n, rnd.choice([1, 10, 100, 128, 129, 400, 1024]))
        dtype = rnd.choice([torch.float16, torch.bfloat16])
        layout_dn = rnd.random() < 0.3
        base, stride = rnd.choice([(0, 1), (3, 8), (1_000_000_007, 1)])
        e, q = _synth(n, d, b, 1000 + trial, dev, dtype)
        m = eng.MipsEngine(d, dtype, dev)
        if layout_dn:
            n_pad = (n + 7) // 8 * 8
            e_dn = torch.zeros(d, n_pad, dtype=dtype, device=dev)
            e_dn[:, :n] = e.T
            m.bind(e_dn[:, :n].t(), id_base=base, id_stride=stride)
        else:
            m.bind(e, id_base=base, id_stride=stride)
        s, i = m.search(q, k)
        assert bool(((i - base) % stride == 0).all()), (trial, "id mapping")
        rows = (i - base) // stride
        rs, ri = _torch_ref(e, q, k, dtype)
        exact = (q.to(dtype).double() @ e.double().T).cpu().numpy()
        rep = O.compare_topk(rows.cpu().numpy(), s.cpu().numpy(), ri.cpu().numpy(), rs.cpu().numpy(), exact, rtol=1e-5, atol=1e-6)
        assert rep["ok"], (trial, n, d, b, k, dtype, layout_dn, rep["errors"][:2])
        m.close()
