"""Scan bandwidth with the index in the reference layout [dim, n] (MN-major operand) vs [n, dim]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jsa_rag_b200 as eng
dev = torch.device("cuda:0")
n = int(os.environ.get("DBG_N", 33_000_000))
g = torch.Generator(device=dev).manual_seed(1)
q = torch.nn.functional.normalize(torch.randn(64, 768, generator=g, device=dev), dim=1)
for layout in ("dn", "nd"):
    e = torch.empty((768, n) if layout == "dn" else (n, 768), dtype=torch.float16, device=dev)
    for s in range(0, n, 1 << 20):
        c = torch.nn.functional.normalize(torch.randn(min(1 << 20, n - s), 768, generator=g, device=dev), dim=1).half()
        if layout == "dn": e[:, s:s + c.shape[0]] = c.T
        else: e[s:s + c.shape[0]] = c
    m = eng.MipsEngine(768, torch.float16, dev); m.bind(e.t() if layout == "dn" else e)
    for _ in range(3): m.search(q, 100)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10): m.search(q, 100)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    print(f"layout [{layout}] n={n}: {ms:.3f} ms/search  {n*1536/ms/1e6:.0f} GB/s", flush=True)
    m.close(); del e
