import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, jsa_rag_b200 as eng
dev = torch.device("cuda:0"); g = torch.Generator(device=dev).manual_seed(1)
n = int(os.environ.get("DBG_N", 33_000_000))
e = torch.empty(n, 768, dtype=torch.float16, device=dev)
for s in range(0, n, 1 << 20):
    c = torch.randn(min(1 << 20, n - s), 768, generator=g, device=dev)
    e[s:s + c.shape[0]] = torch.nn.functional.normalize(c, dim=1).half()
m = eng.MipsEngine(768, torch.float16, dev); m.bind(e)
for b in (64, 256, 512):
    q = torch.nn.functional.normalize(torch.randn(b, 768, generator=g, device=dev), dim=1)
    for _ in range(2): m.search(q, 100)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(6): m.search(q, 100)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 6
    print(f"CPS={os.environ.get('JSA_MIPS_CPS','default')} B={b}: {ms:.2f} ms  {b/ms*1e3:.0f} q/s", flush=True)
