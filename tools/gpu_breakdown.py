"""Where does the scan kernel's time go?  Times the search with the select / MMA stages disabled and
prints the per-CTA cycle counters.  Developer tool (B200 only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jsa_rag_b200 as eng

dev = torch.device("cuda:0")
n = int(os.environ.get("DBG_N", 4_125_000)); b = int(os.environ.get("DBG_B", 64)); k = int(os.environ.get("DBG_K", 100))
g = torch.Generator(device=dev).manual_seed(1)
e = torch.empty(n, 768, dtype=torch.float16, device=dev)
for s in range(0, n, 1 << 20):
    c = torch.randn(min(1 << 20, n - s), 768, generator=g, device=dev)
    e[s:s + c.shape[0]] = torch.nn.functional.normalize(c, dim=1).half()
q = torch.nn.functional.normalize(torch.randn(b, 768, generator=g, device=dev), dim=1)
m = eng.MipsEngine(768, torch.float16, dev); m.bind(e)


def timeit(iters=10):
    for _ in range(3): m.search(q, k)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(iters): m.search(q, k)
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / iters


for name, flags in [("full", 0)]:
    m.debug_config(flags, False)
    ms = timeit()
    print(f"{name:10s}: {ms:.3f} ms  {n*1536/ms/1e6:.0f} GB/s", flush=True)
    st = m.debug_config(flags, True)
    m.search(q, k); torch.cuda.synchronize()
    stf = st.double()
    print("   per-CTA mean:", {nm: round(v, 0) for nm, v in zip(m.STAT_NAMES, stf.mean(0).tolist())})
    print("   per-CTA max :", {nm: round(v, 0) for nm, v in zip(m.STAT_NAMES, stf.max(0).values.tolist())})
m.debug_config(0, False)
