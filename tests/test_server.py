"""Index-server path (reference build_server/server_start.py + src/post.py): wire format and stream
loader on CPU; the search itself on the GPU."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import flat_index_oracle as O


class _StubIndex:
    def __init__(self):
        self.calls = []

    def search_knn(self, q, topk):
        self.calls.append((tuple(q.shape), topk))
        return [[{"id": str(j)} for j in range(topk)] for _ in range(q.shape[0])], [[1.0 / (j + 1) for j in range(topk)] for _ in range(q.shape[0])]


def test_http_contract_and_client(eng):
    from fastapi.testclient import TestClient
    holder = eng.IndexHolder(None)
    notified = []
    app = eng.create_app(holder, rebuild_fn=lambda ckpt: _StubIndex(), notify=lambda url, body: notified.append((url, body)))
    client = TestClient(app)
    # 500 while the index is missing (server_start.py:184-185); the client prints and returns None (src/post.py:30-31)
    r = client.post("/retrieve", json={"query_embs": [0.0] * 8, "bsz": 2, "topk": 3})
    assert r.status_code == 500
    q = torch.arange(8, dtype=torch.float64).view(2, 4)

    class _Sess:
        def post(self, url, json):
            return client.post("/retrieve", json=json)
    assert eng.call_retrieve_api(q, 3, session=_Sess()) is None
    # /rebuild swaps the index and notifies response_url (server_start.py:191-196)
    r = client.post("/rebuild", json={"checkpoint_path": "ckpt", "response_url": "http://cb"})
    assert r.status_code == 200 and notified == [("http://cb", {"status": "success"})]
    stub = holder.get()
    docs, scores = eng.call_retrieve_api(q, 3, session=_Sess())
    assert stub.calls == [((2, 4), 3)]                       # query_embs.view(bsz, -1)
    assert docs == [[{"id": "0"}, {"id": "1"}, {"id": "2"}]] * 2 and scores[0] == [1.0, 0.5, 1.0 / 3]
    # defaults of the request model: bsz=1, topk=10 (server_start.py:18-21)
    r = client.post("/retrieve", json={"query_embs": [0.0] * 4})
    assert r.status_code == 200 and len(r.json()[0]) == 1 and len(r.json()[0][0]) == 10


def test_binary_endpoints(eng):
    import numpy as np
    from fastapi.testclient import TestClient

    class _TensorStub(_StubIndex):
        def search(self, q, topk):
            b = q.shape[0]
            return (torch.arange(b * topk, dtype=torch.float32).view(b, topk), torch.arange(b * topk).view(b, topk) + 7)

    holder = eng.IndexHolder(_TensorStub())
    client = TestClient(eng.create_app(holder))
    q = torch.randn(3, 8)
    r = client.post("/retrieve_bin", params={"bsz": 3, "topk": 4}, content=q.numpy().tobytes())
    assert r.status_code == 200 and holder.get().calls[-1] == ((3, 8), 4) and len(r.json()[0]) == 3

    class _Sess:                                        # requests.Session stand-in over the ASGI test client
        def post(self, url, params=None, data=None, headers=None, json=None):
            return client.post(url[url.index("/", 8):], params=params, content=data, headers=headers, json=json)
    from importlib import import_module
    rc = import_module("jsa-rag_b200.client").RetrieveClient("http://testserver/retrieve", session=_Sess())
    s_, i_ = rc.search_binary(q, 4)
    assert s_.shape == (3, 4) and i_.dtype == np.int64 and i_[0, 0] == 7 and s_[2, 3] == 11.0
    docs_b, scores_b = rc.retrieve_binary(q, 4)
    assert len(docs_b) == 3 and len(scores_b[0]) == 4
    r = client.post("/retrieve_bin", params={"bsz": 3, "topk": 4, "dtype": "fp16"}, content=q.half().numpy().tobytes())
    assert r.status_code == 200 and holder.get().calls[-1] == ((3, 8), 4)
    r = client.post("/search_bin", params={"bsz": 3, "topk": 4}, content=q.numpy().tobytes())
    body = r.content
    assert r.status_code == 200 and len(body) == 3 * 4 * 12
    assert np.frombuffer(body[:48], dtype="<f4").tolist() == list(range(12))
    assert np.frombuffer(body[48:], dtype="<i8").tolist() == [x + 7 for x in range(12)]
    assert client.post("/retrieve_bin", params={"bsz": 5, "topk": 4}, content=q.numpy().tobytes()).status_code == 422


def test_embedding_stream_roundtrip(eng, tmp_path):
    g = load_golden("flat_n300_d1024_b5_k10")
    path = str(tmp_path / "embeddings_0.pkl")
    passages = [{"id": str(i), "text": f"p{i}"} for i in range(300)]
    for a in range(0, 300, 64):                              # appended batch by batch (build_server/index.py:108-111)
        eng.append_embedding_batch(path, g["embeddings"][a:a + 64], passages[a:a + 64])
    batches = list(eng.iter_embedding_stream(path))
    assert [len(b) for b in batches] == [64, 64, 64, 64, 44]
    assert batches[1][3]["passage"] == passages[67] and np.array_equal(batches[1][3]["emb"], g["embeddings"][67])
    assert eng.get_pkl_files_in_directory(str(tmp_path)) == [path]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_server_index_has_no_cpu_fallback(eng, tmp_path):
    path = str(tmp_path / "embeddings_0.pkl")
    eng.append_embedding_batch(path, np.zeros((4, 64), dtype=np.float16), [{"id": str(i)} for i in range(4)])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        eng.B200ServerIndex([path])


@pytest.mark.gpu
def test_server_search_matches_faiss_semantics(eng, tmp_path):
    """normalize_L2(queries) -> exact IP over fp16 vectors -> ids in insertion order (server_start.py:139-163)."""
    from fastapi.testclient import TestClient
    g = load_golden("flat_n300_d1024_b5_k10")
    e16 = g["embeddings"]
    passages = [{"id": str(i), "text": f"p{i}"} for i in range(300)]
    n_files = min(2, torch.cuda.device_count())
    cuts = [0, 300] if n_files == 1 else [0, 170, 300]
    files = []
    for f in range(n_files):
        path = str(tmp_path / f"embeddings_{f}.pkl")
        for a in range(cuts[f], cuts[f + 1], 50):
            b = min(a + 50, cuts[f + 1])
            eng.append_embedding_batch(path, e16[a:b], passages[a:b])
        files.append(path)
    index = eng.B200ServerIndex(files, None)
    assert index.ntotal == 300 and index.dimension == 1024 and index.doc_map[299] == passages[299]
    q = g["queries"] * np.array([[3.0], [0.5], [10.0], [1.0], [7.0]], dtype=np.float32)   # un-normalised on purpose
    docs, scores = index.search_knn(torch.from_numpy(q), 10)
    d_ref, i_ref = O.server_search(q, e16, 10)
    ids = np.array([[int(d["id"]) for d in row] for row in docs])
    exact = O.exact_scores(O.normalize_l2(q), e16, q_dtype=None)
    rep = O.compare_topk(ids, np.array(scores), i_ref, d_ref, exact, rtol=1e-3)
    assert rep["ok"], rep["errors"][:3]
    # through HTTP with the reference's client
    client = TestClient(eng.create_app(eng.IndexHolder(index)))

    class _Sess:
        def post(self, url, json):
            return client.post("/retrieve", json=json)
    docs2, scores2 = eng.call_retrieve_api(torch.from_numpy(q), 10, session=_Sess())
    assert docs2 == docs and np.allclose(np.array(scores2), np.array(scores))
    with pytest.raises(RuntimeError, match="selected index k out of range"):
        index.search_knn(torch.from_numpy(q), 301)


def test_retrieve_request_parser(eng):
    """The hand-written body parser of POST /retrieve accepts what the reference's pydantic model accepts
    (build_server/server_start.py:18-21) and rejects the rest with 422."""
    import json
    from importlib import import_module
    from fastapi import HTTPException
    S = import_module("jsa-rag_b200.server")
    rng = np.random.default_rng(1)
    q = rng.standard_normal(6 * 16).astype(np.float32)
    r = S.parse_retrieve_request(json.dumps({"query_embs": q.tolist(), "bsz": 6, "topk": 5}).encode())
    assert r["bsz"] == 6 and r["topk"] == 5 and np.array_equal(r["query_embs"], q)          # C scanner, exact
    r = S.parse_retrieve_request(json.dumps({"topk": 3, "query_embs": [1, 2.5, -3e-2, 4E1]}).encode())
    assert r["bsz"] == 1 and r["query_embs"].tolist() == [1.0, 2.5, np.float32(-0.03), 40.0]
    r = S.parse_retrieve_request(b'{"query_embs": [[1.0, 2.0], [3.0, 4.0]], "bsz": 2}')       # nested: general json path
    assert r["query_embs"].tolist() == [1.0, 2.0, 3.0, 4.0] and r["topk"] == 10
    r = S.parse_retrieve_request(b'{"query_embs": [1.0, NaN], "bsz": 1}')                     # not a plain number: json path
    assert np.isnan(r["query_embs"][1])
    for bad in (b"not json", b"[1, 2]", b'{"bsz": 2}', b'{"query_embs": "x"}', b'{"query_embs": [1, 2, 3], "bsz": 2}',
                b'{"query_embs": [1.0], "bsz": 1.5}', b'{"query_embs": [1.0], "bsz": 0}', b'{"query_embs": ["a"]}'):
        with pytest.raises(HTTPException) as ei:
            S.parse_retrieve_request(bad)
        assert ei.value.status_code == 422
