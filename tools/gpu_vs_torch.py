"""Like-for-like GPU baseline: the reference's own two lines (src/index.py:118-119: torch.matmul on fp16 +
torch.topk) on the same B200, same index, next to the fused engine."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, jsa_rag_b200 as eng
dev = torch.device("cuda:0"); g = torch.Generator(device=dev).manual_seed(1)
n = int(os.environ.get("DBG_N", 33_000_000)); k = 100
e = torch.empty(n, 768, dtype=torch.float16, device=dev)
for s in range(0, n, 1 << 20):
    c = torch.randn(min(1 << 20, n - s), 768, generator=g, device=dev)
    e[s:s + c.shape[0]] = torch.nn.functional.normalize(c, dim=1).half()
m = eng.MipsEngine(768, torch.float16, dev); m.bind(e)
emb_dn = e.t()        # [768, n] view = the reference's operand layout (values identical)
for b in [int(x) for x in os.environ.get("DBG_BS", "64,256").split(",")]:
    q = torch.nn.functional.normalize(torch.randn(b, 768, generator=g, device=dev), dim=1)
    def ref():
        return torch.topk(torch.matmul(q.half(), emb_dn), k, dim=1)
    def ours():
        return m.search(q, k)
    out = {}
    for name, fn, iters in (("reference lines (torch.matmul + torch.topk)", ref, 3), ("fused engine", ours, 10)):
        try:
            for _ in range(2): fn()
            torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(iters): r = fn()
            t1.record(); torch.cuda.synchronize()
            out[name] = t0.elapsed_time(t1) / iters
            print(f"n={n} B={b} k={k} {name}: {out[name]:.2f} ms  {b/out[name]*1e3:.0f} q/s", flush=True)
        except Exception as ex:
            print(f"n={n} B={b} {name}: failed: {type(ex).__name__}: {str(ex)[:120]}", flush=True)
    if len(out) == 2:
        print(f"   speed-up {list(out.values())[0] / list(out.values())[1]:.1f}x", flush=True)
