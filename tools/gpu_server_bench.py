"""Index-server path at the reference's shape (build_server/server_start.py:31-35,139-189): dim 1024, one shard per
GPU, batch 64, top-100, L2-normalised queries.  Reports q/s and p50 / p90 latency of
  * B200ServerIndex.search            (device ids/scores, the fused multi-GPU flow)
  * B200ServerIndex.search_knn        (reference return type: docs + float scores)
  * POST /retrieve  (reference JSON schema), /retrieve_bin, /search_bin through the ASGI app (in-process TestClient)
and checks the merged answer against a single engine over the concatenated shards.
SRV_ROWS (rows per GPU, default 8M), SRV_GPUS (default all), SRV_B (64), SRV_K (100), SRV_ITERS (30)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import jsa_rag_b200 as eng
from bench import PooledDocMap

n_gpu = int(os.environ.get("SRV_GPUS", torch.cuda.device_count()))
rows = int(os.environ.get("SRV_ROWS", 8_000_000))
B, K, D = int(os.environ.get("SRV_B", 64)), int(os.environ.get("SRV_K", 100)), 1024
iters = int(os.environ.get("SRV_ITERS", 30))

stores = []
for g in range(n_gpu):
    dev = torch.device("cuda", g)
    gen = torch.Generator(device=dev).manual_seed(10 + g)
    st = torch.empty(rows, D, dtype=torch.float16, device=dev)
    for a in range(0, rows, 1 << 20):
        c = torch.randn(min(1 << 20, rows - a), D, generator=gen, device=dev)
        st[a:a + c.shape[0]] = (c * (0.5 + torch.rand(c.shape[0], 1, generator=gen, device=dev))).half()   # NOT normalised
    stores.append(st)
server = eng.B200ServerIndex.from_tensors(stores, PooledDocMap(rows * n_gpu))
q = torch.randn(B, D, generator=torch.Generator().manual_seed(3)) * 3.0

# ---- parity of the fused multi-GPU flow: one engine over a sample of every shard == the server over the same sample
sample = 200_000
small = eng.B200ServerIndex.from_tensors([s[:sample] for s in stores], PooledDocMap(sample * n_gpu))
ss, si = small.search(q, K)
cat = torch.cat([s[:sample].to("cuda:0") for s in stores])
one = eng.MipsEngine(D, torch.float16, torch.device("cuda:0")); one.bind(cat)
os_, oi = one.search(q.to("cuda:0"), K, normalize=True)
torch.cuda.synchronize()
same = bool(torch.equal(si, oi)) and bool(torch.equal(ss, os_))
qn = torch.nn.functional.normalize(q.to("cuda:0"), dim=1)
exact = qn.half().float() @ cat.float().T
rs, ri = torch.topk(exact, K, dim=1)
set_eq = sum(set(a.tolist()) == set(c.tolist()) for a, c in zip(si, ri))
print(json.dumps({"check": "server merge == single engine over the concatenated shards (bit-identical)", "ok": same,
                  "id_sets_equal_to_fp32_topk": f"{set_eq}/{B}"}), flush=True)
del one, cat, small, exact


def timeit(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    lat = []
    t_all = time.perf_counter()
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        lat.append(time.perf_counter() - t0)
    total = time.perf_counter() - t_all
    lat.sort()
    return {"qps": B * n / total, "p50_ms": 1e3 * lat[len(lat) // 2], "p90_ms": 1e3 * lat[int(len(lat) * 0.9)]}


def dev_search():
    s, i = server.search(q, K)
    s.cpu(); i.cpu()


algo_bytes = rows * D * 2
out = {"shape": {"gpus": n_gpu, "rows_per_gpu": rows, "dim": D, "batch": B, "k": K},
       "hbm_floor_ms": algo_bytes / 6542.1e9 * 1e3}
out["search (device flow + D2H)"] = timeit(dev_search, iters)
out["search_knn (docs + scores)"] = timeit(lambda: server.search_knn(q, K), iters)

from fastapi.testclient import TestClient
holder = eng.IndexHolder(server)
client = TestClient(eng.create_app(holder))
payload = {"query_embs": q.reshape(-1).tolist(), "bsz": B, "topk": K}
body_json = json.dumps(payload).encode()        # what the reference client sends (src/post.py:10-28), encoded once:
body32 = q.numpy().astype("<f4").tobytes()      # the timings below are the server's side of a request


def http_json():
    r = client.post("/retrieve", content=body_json, headers={"content-type": "application/json"})
    assert r.status_code == 200
    return r


def http_bin():
    r = client.post(f"/retrieve_bin?bsz={B}&topk={K}&dtype=fp32", content=body32)
    assert r.status_code == 200
    return r


def http_search_bin():
    r = client.post(f"/search_bin?bsz={B}&topk={K}&dtype=fp32", content=body32)
    assert r.status_code == 200
    return r.content


out["POST /retrieve (reference JSON)"] = timeit(http_json, max(5, iters // 3))
out["POST /retrieve_bin"] = timeit(http_bin, max(5, iters // 3))
out["POST /search_bin"] = timeit(http_search_bin, iters)
docs, scores = http_json().json()
out["request_bytes"] = {"json": len(body_json), "binary": len(body32)}
out["response_bytes"] = {"json": len(http_json().content), "search_bin": len(http_search_bin())}
raw = http_search_bin()
ids = np.frombuffer(raw[B * K * 4:], dtype="<i8").reshape(B, K)
ds, di = server.search(q, K)
out["routes_agree"] = bool((ids == di.cpu().numpy()).all()) and len(docs) == B and len(docs[0]) == K
print(json.dumps(out), flush=True)
