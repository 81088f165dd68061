import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jsa_rag_b200 as eng
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
n = 4_125_000
e = torch.empty(n, 768, dtype=torch.float16, device=dev)
for s in range(0, n, 1 << 20):
    c = torch.randn(min(1 << 20, n - s), 768, generator=g, device=dev)
    e[s:s + c.shape[0]] = torch.nn.functional.normalize(c, dim=1).half()
q = torch.nn.functional.normalize(torch.randn(64, 768, generator=g, device=dev), dim=1)
m = eng.MipsEngine(768, torch.float16, dev); m.bind(e)
idx = eng.B200Index(); idx._store = e; idx._set_sharding("round_robin")
for name, fn in (("engine.search", lambda: m.search(q, 100)), ("index.search", lambda: idx.search(q, 100))):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200): fn()
    t_issue = (time.perf_counter() - t0) / 200
    torch.cuda.synchronize()
    t_total = (time.perf_counter() - t0) / 200
    print(f"{name}: CPU issue {t_issue*1e6:.0f} us/call, wall {t_total*1e6:.0f} us/call", flush=True)
# CUDA-graph replay of the same call
out_s = torch.empty(64, 100, device=dev); out_i = torch.empty(64, 100, dtype=torch.int64, device=dev)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): m.search(q, 100, out=(out_s, out_i))
torch.cuda.current_stream().wait_stream(s)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    m.search(q, 100, out=(out_s, out_i))
ref_s, ref_i = m.search(q, 100)
gr.replay(); torch.cuda.synchronize()
print("graph replay equals eager:", torch.equal(out_s, ref_s), torch.equal(out_i, ref_i))
t0 = time.perf_counter()
for _ in range(200): gr.replay()
t_issue = (time.perf_counter() - t0) / 200
torch.cuda.synchronize()
print(f"graph replay: CPU issue {t_issue*1e6:.0f} us/call, wall {(time.perf_counter()-t0)/200*1e6:.0f} us/call")
