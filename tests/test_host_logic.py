"""CPU-only tests of the host layer: C-ABI surface, storage / shard I/O compatible with the reference,
loud failure without the CUDA device, and the world_size-2 plumbing over gloo."""
import ctypes
import importlib
import json
import os
import pickle
import re
import socket

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from oracle import ref_import


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_cabi_exports_every_declared_symbol(eng):
    header = open(os.path.join(ROOT, "include", "jsa_mips.h")).read()
    declared = set(re.findall(r"\b(mips_[a-z_0-9]+)\s*\(", header))
    declared.discard("mips_handle")
    assert {"mips_create", "mips_bind_index", "mips_search_local", "mips_merge_topk", "mips_gather_rows",
            "mips_search_host", "mips_workspace_bytes", "mips_last_error", "mips_destroy"} <= declared
    lib = ctypes.CDLL(eng._native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in jsa_mips.h but not exported"
    assert set(eng._native.SYMBOLS) == declared, "ctypes table and header disagree"
    n = eng._native.load()
    assert n.mips_abi_version() == 1 and n.mips_max_k() >= 100 and n.mips_max_dim() >= 1024


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(eng):
    n = eng._native.load()
    h = ctypes.c_void_p()
    rc = n.mips_create(ctypes.byref(h), 0, 768, 0)
    assert rc == eng._native.MIPS_EUNSUPPORTED and "no CPU fallback" in eng._native.last_error(None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        eng.MipsEngine(768)
    idx = eng.B200Index()
    idx.init_embeddings([{"id": str(i)} for i in range(10)], dim=768)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        idx.search_knn(torch.randn(2, 768), 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        eng.merge_topk(torch.zeros(2, 1, 4), torch.zeros(2, 1, 4, dtype=torch.int64), 4)


def test_create_rejects_bad_arguments(eng):
    n = eng._native.load()
    h = ctypes.c_void_p()
    assert n.mips_create(ctypes.byref(h), 0, 100, 0) == eng._native.MIPS_EINVAL   # dim not a multiple of 64
    assert "multiple of 64" in eng._native.last_error(None)
    assert n.mips_create(ctypes.byref(h), 0, 768, 7) == eng._native.MIPS_EINVAL   # unknown dtype
    assert n.mips_create(None, 0, 768, 0) == eng._native.MIPS_EINVAL


def test_storage_is_reference_shaped_view(eng):
    idx = eng.B200Index(device="cpu")
    assert idx.embeddings is None and idx.doc_map == {} and idx.is_index_trained()
    passages = [{"id": str(i), "text": f"p{i}"} for i in range(37)]
    idx.init_embeddings(passages, dim=768)
    emb = idx.embeddings
    assert tuple(emb.shape) == (768, 37) and emb.dtype == torch.float16 and float(emb.abs().sum()) == 0.0
    x = torch.randn(10, 768)
    idx.embeddings[:, 5:15] = x.T                 # src/rag.py:120
    assert torch.equal(idx._store[5:15], x.half()) and torch.equal(idx.embeddings[:, 5:15], x.half().T)
    assert idx._store.is_contiguous() and len(idx.doc_map) == 37 and idx.doc_map[3]["id"] == "3"
    idx.embeddings = torch.randn(768, 12)         # attribute assignment like load_index in the reference
    assert tuple(idx._store.shape) == (12, 768)


def test_reference_layout_storage_option(eng, tmp_path):
    """layout="dn": `.embeddings` is a real contiguous [dim, n] tensor like the reference's; same files on disk."""
    idx = eng.B200Index(device="cpu", layout="dn")
    idx.init_embeddings([{"id": str(i)} for i in range(24)], dim=768)
    assert idx.embeddings.is_contiguous() and tuple(idx.embeddings.shape) == (768, 24) and tuple(idx._store.shape) == (24, 768)
    x = torch.randn(24, 768)
    idx.embeddings[:, :] = x.T
    assert torch.equal(idx._store, x.half())
    idx.save_index(str(tmp_path), 2)
    a = eng.B200Index(device="cpu")
    a.load_index(str(tmp_path), 2)
    b = eng.B200Index(device="cpu", layout="dn")
    b.load_index(str(tmp_path), 2)
    assert torch.equal(a._store, x.half()) and torch.equal(b._store, x.half()) and b.embeddings.is_contiguous()
    b.embeddings = torch.randn(768, 8)
    assert b.embeddings.is_contiguous() and tuple(b._store.shape) == (8, 768)


def test_shard_files_match_reference_format(eng, tmp_path):
    g = load_golden("flat_n1003_d768_b8_k20")
    n = int(g["n"])
    passages = [{"id": str(i), "title": f"t{i}", "text": f"passage {i}"} for i in range(n)]
    idx = eng.B200Index(device="cpu")
    idx.init_embeddings(passages, dim=768)
    idx.embeddings[:, :] = torch.from_numpy(g["embeddings"]).T
    idx.save_index(str(tmp_path), 4)
    files = sorted(os.listdir(tmp_path))
    assert files == [f"embeddings.{i}.pt" for i in range(4)] + [f"passages.{i}.pt" for i in range(4)]
    at = 0
    for i, width in enumerate([251, 251, 251, 250]):
        t = torch.load(tmp_path / f"embeddings.{i}.pt")
        assert tuple(t.shape) == (768, width) and t.dtype == torch.float16 and t.is_contiguous()
        assert torch.equal(t, torch.from_numpy(g["embeddings"][at:at + width]).T)
        with open(tmp_path / f"passages.{i}.pt", "rb") as f:
            assert pickle.load(f) == passages[at:at + width]    # raw pickle, not torch.save (src/index.py:84-85)
        at += width
    idx2 = eng.B200Index(device="cpu")
    idx2.load_index(str(tmp_path), 4)
    assert torch.equal(idx2._store, idx._store) and idx2.doc_map == idx.doc_map and idx2._sharding == "contiguous"
    # embeddings are always rewritten, passages only when missing / overwrite requested (src/index.py:82)
    passages[0]["text"] = "changed"
    idx.doc_map[0] = passages[0]
    idx.save_index(str(tmp_path), 4)
    with open(tmp_path / "passages.0.pt", "rb") as f:
        assert pickle.load(f)[0]["text"] == "passage 0"
    idx.save_index(str(tmp_path), 4, overwrite_saved_passages=True)
    with open(tmp_path / "passages.0.pt", "rb") as f:
        assert pickle.load(f)[0]["text"] == "changed"


@pytest.mark.skipif(not ref_import.reference_available(), reason="/root/reference not present (GPU box)")
def test_shard_files_interchange_with_live_reference(eng, tmp_path):
    from oracle import make_golden
    e16, q = make_golden.synth_inputs(203, 768, 2, seed=5)
    _, _, ref_idx = make_golden.run_reference(e16, q, 5)
    d_ref, d_ours = tmp_path / "ref", tmp_path / "ours"
    d_ref.mkdir(), d_ours.mkdir()
    ref_idx.save_index(str(d_ref), 2)                  # written by the unmodified reference
    ours = eng.B200Index(device="cpu")
    ours.load_index(str(d_ref), 2)
    assert torch.equal(ours.embeddings, ref_idx.embeddings) and ours.doc_map == ref_idx.doc_map
    ours.save_index(str(d_ours), 2)
    for f in sorted(os.listdir(d_ref)):
        if f.startswith("embeddings"):
            assert torch.equal(torch.load(d_ref / f), torch.load(d_ours / f))
            assert os.path.getsize(d_ref / f) == os.path.getsize(d_ours / f)
        else:
            assert open(d_ref / f, "rb").read() == open(d_ours / f, "rb").read()


def test_load_passages_and_factory(eng, tmp_path):
    p = tmp_path / "p.jsonl"
    with open(p, "w") as f:
        for i in range(7):
            rec = {"id": str(i), "title": "T", "text": "x"}
            if i % 2:
                rec["section"] = "S"
            f.write(json.dumps(rec) + "\n")
    ps = eng.load_passages([str(p)])
    assert len(ps) == 7 and ps[1]["title"] == "T: S" and ps[0]["title"] == "T"   # src/index_io.py:30-31
    assert len(eng.load_passages([str(p)], maxload=3)) == 3

    class Opt:
        index_mode = "flat"; load_index_path = None; use_file_passages = False
        passages = [str(p)]; max_passages = -1; retriever_model_path = "facebook/contriever"; save_index_n_shards = 1
    index, passages = eng.load_or_initialize_index(Opt())
    assert isinstance(index, eng.B200Index) and tuple(index.embeddings.shape) == (768, 7) and passages == ps
    Opt.retriever_model_path = "BAAI/bge-large-en"
    assert eng.load_or_initialize_index(Opt())[0].embeddings.shape[0] == 1024     # src/index_io.py:92
    Opt.index_mode = "b200"
    assert isinstance(eng.load_or_initialize_index(Opt())[0], eng.B200Index)
    Opt.index_mode = "annoy"
    with pytest.raises(ValueError, match="unsupported index mode"):
        eng.load_or_initialize_index(Opt())
    Opt.index_mode = "faiss"; Opt.faiss_index_type = "ivfpq"
    with pytest.raises(ValueError):
        eng.load_or_initialize_index(Opt())


def test_embeddings_setter_updates_the_sharding(eng):
    """Direct assignment (which the setter supports) must refresh shard sizes and the global-id mapping."""
    idx = eng.B200Index(device="cpu")
    idx.is_in_gpu = False
    idx.embeddings = torch.randn(768, 37)
    assert idx._all_counts == [37] and (idx._id_base, idx._id_stride) == (0, 1)
    idx.embeddings = torch.randn(768, 11)
    assert idx._all_counts == [11]


def test_candidate_packing_single_process(eng):
    s = torch.randn(3, 5)
    i = torch.randint(0, 1 << 40, (3, 5))
    gs, gi = eng.dist_utils.all_gather_candidates(s, i)
    assert torch.equal(gs[0], s) and torch.equal(gi[0], i)
    assert eng.dist_utils.get_world_size() == 1 and eng.dist_utils.get_rank() == 0
    assert torch.equal(eng.dist_utils.varsize_all_gather(s), s)


@pytest.mark.parametrize("mode", ["round_robin", "contiguous"])
def test_two_rank_search_over_gloo(mode, tmp_path):
    import torch.multiprocessing as mp
    import _dist_worker
    mp.spawn(_dist_worker.worker, args=(2, _free_port(), str(tmp_path), mode), nprocs=2, join=True)


def test_filter_results_by_id(eng):
    """src/tasks/base.py:96-148: drop the source passage, re-append violators when short, cut to topk."""
    docs = [[{"id": "a"}, {"id": "b"}, {"id": "c"}, {"id": "d"}], [{"id": "x"}, {"id": "y"}, {"id": "z"}, {"id": "w"}]]
    scores = [[4.0, 3.0, 2.0, 1.0], [8.0, 7.0, 6.0, 5.0]]
    meta = [{"id": "b"}, {"id": "none"}]
    p, s = eng.filter_results_by_id(meta, docs, scores, 2)
    assert [[d["id"] for d in r] for r in p] == [["a", "c"], ["x", "y"]] and [list(r) for r in s] == [[4.0, 2.0], [8.0, 7.0]]
    p, s = eng.filter_results_by_id(meta, docs, scores, 4)           # short after filtering: violator appended back
    assert [d["id"] for d in p[0]] == ["a", "c", "d", "b"] and list(s[0]) == [4.0, 2.0, 1.0, 3.0]
    p, s = eng.filter_results_by_id(None, docs, scores, 3)           # padding instance: plain cut
    assert [len(r) for r in p] == [3, 3] and s[1] == [8.0, 7.0, 6.0]


def test_vectorised_filter_matches_the_reference_loop(eng):
    """Randomised: the id-array filter must give what the reference's per-pair loop gives (restated here)."""
    import random
    from importlib import import_module
    F = import_module("jsa-rag_b200.filtering")
    rnd = random.Random(7)
    for trial in range(50):
        b, kk, topk = rnd.randint(1, 6), rnd.randint(1, 12), rnd.randint(1, 12)
        docs = [[{"id": str(rnd.randint(0, 5)), "n": j} for j in range(kk)] for _ in range(b)]
        scores = [[float(kk - j) for j in range(kk)] for _ in range(b)]
        meta = [{"id": str(rnd.randint(0, 5))} for _ in range(b)]
        want_p, want_s = [], []
        for m, ps, ss in zip(meta, docs, scores):                      # src/tasks/base.py:127-146
            keep = [(p, s) for p, s in zip(ps, ss) if p["id"] != m["id"]]
            viol = [(p, s) for p, s in zip(ps, ss) if p["id"] == m["id"]]
            both = keep + viol
            want_p.append([p for p, _ in both][:topk]); want_s.append([s for _, s in both][:topk])
        got_p, got_s = F.filter_results_by_id(meta, docs, scores, topk)
        assert [list(r) for r in got_p] == want_p and [list(r) for r in got_s] == want_s
        assert all(a is b_ for ra, rb in zip(got_p, want_p) for a, b_ in zip(ra, rb))     # the same dict objects
    pos, kept = F.filter_positions(np.array([5, 9]), np.array([[5, 1, 5, 2], [1, 2, 3, 4]]), 3)
    assert pos.tolist() == [[1, 3, 0], [0, 1, 2]] and kept.tolist() == [2, 4]
    ragged_p, ragged_s = F.filter_results_by_id([{"id": "a"}, {"id": "b"}], [[{"id": "a"}, {"id": "c"}], [{"id": "b"}]],
                                                [[2.0, 1.0], [3.0]], 2)
    assert [[d["id"] for d in r] for r in ragged_p] == [["c", "a"], ["b"]]


def test_passage_store_round_trip(eng):
    from importlib import import_module
    PS = import_module("jsa-rag_b200.passages").PassageStore
    shards = [[{"id": str(3 * i + r), "title": f"t{i}", "text": "x" * (i % 7)} for i in range(20 + r)] for r in range(3)]
    st = PS.from_shards(shards)
    assert [st.shard_len(r) for r in range(3)] == [20, 21, 22]
    assert st.get(2, 21) == shards[2][21] and st.get(0, 0) == shards[0][0]
    owners = np.array([0, 2, 1, 1]); rows = np.array([19, 0, 20, 3])
    assert st.get_many(owners, rows) == [shards[o][r] for o, r in zip(owners, rows)]
    st.close()


def test_peer_exchange_is_optional_and_never_a_cpu_path(eng, monkeypatch):
    """Without an NCCL process group (single process, gloo, CPU) no exchange object is created — callers keep the
    collective path — and the mode switch is read from the environment."""
    import importlib
    ex = importlib.import_module("jsa-rag_b200.exchange")
    assert ex.make_peer_exchange("cpu", 1 << 20) is None
    monkeypatch.setenv("JSA_MIPS_EXCHANGE", "nccl")
    assert ex.exchange_mode() == "nccl"
    monkeypatch.delenv("JSA_MIPS_EXCHANGE")
    assert ex.exchange_mode() == "p2p"
    lib = eng._native.load()
    assert lib.mips_xchg_handle_bytes() == 64
    # argument validation happens before any CUDA call
    import ctypes
    h = ctypes.c_void_p()
    assert lib.mips_xchg_create(ctypes.byref(h), 0, 3, 2, 1024) == eng._native.MIPS_EINVAL      # rank >= world
    assert lib.mips_xchg_create(ctypes.byref(h), 0, 0, 17, 1024) == eng._native.MIPS_EINVAL     # world > 16
    assert lib.mips_xchg_merge(None, None, 0, 0, 1, 1, 1, None, None, None) == eng._native.MIPS_EINVAL
