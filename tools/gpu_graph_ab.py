"""Eager vs CUDA-graph replay of one search (single GPU), back to back and with a host sync per search."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jsa_rag_b200 as eng
dev = torch.device("cuda:0")
n = int(os.environ.get("DBG_N", 4_125_000))
g = torch.Generator(device=dev).manual_seed(1)
e = torch.empty(n, 768, dtype=torch.float16, device=dev)
for s in range(0, n, 1 << 20):
    c = torch.randn(min(1 << 20, n - s), 768, generator=g, device=dev)
    e[s:s + c.shape[0]] = torch.nn.functional.normalize(c, dim=1).half()
idx = eng.B200Index(); idx._store = e; idx._set_sharding("round_robin")
for b in [int(x) for x in os.environ.get("DBG_B", "64,256,512,1024").split(",")]:
    q = torch.nn.functional.normalize(torch.randn(b, 768, generator=g, device=dev), dim=1)
    run = idx.make_graphed_search(b, 100)
    ref_s, ref_i = idx.search(q, 100)
    gs, gi = run(q); torch.cuda.synchronize()
    same = torch.equal(gi, ref_i)
    res = {}
    for name, fn in (("eager", lambda: idx.search(q, 100)), ("graph", lambda: run(q))):
        for mode in ("back-to-back", "sync-per-search"):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(30):
                fn()
                if mode == "sync-per-search": torch.cuda.synchronize()
            torch.cuda.synchronize()
            res[(name, mode)] = (time.perf_counter() - t0) / 30 * 1e3
    print(f"B={b} graph==eager {same}: " + "  ".join(f"{k[0]}/{k[1]} {v:.3f} ms" for k, v in res.items()), flush=True)
    run.release()
