"""Sustained-throughput A/B: many back-to-back searches (power-capped regime), M=64 vs forced M=128."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jsa_rag_b200 as eng
dev = torch.device("cuda:0")
n = int(os.environ.get("DBG_N", 33_000_000)); k = 100
g = torch.Generator(device=dev).manual_seed(1)
e = torch.empty(n, 768, dtype=torch.float16, device=dev)
for s in range(0, n, 1 << 20):
    c = torch.randn(min(1 << 20, n - s), 768, generator=g, device=dev)
    e[s:s + c.shape[0]] = torch.nn.functional.normalize(c, dim=1).half()
m = eng.MipsEngine(768, torch.float16, dev); m.bind(e)
for b in (64, 128, 1):
    q = torch.nn.functional.normalize(torch.randn(b, 768, generator=g, device=dev), dim=1)
    for name, flags in (("m64/auto", 0), ("force-m128", 16)):
        m.debug_config(flags, False)
        for _ in range(5): m.search(q, k)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        iters = 60
        t0.record()
        for _ in range(iters): m.search(q, k)
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / iters
        smi = os.popen("nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader").read().strip()
        print(f"B={b:4d} {name:11s}: {ms:.3f} ms/search  {n*1536/ms/1e6:.0f} GB/s  {b/ms*1e3:.0f} q/s   [{smi}]", flush=True)
