"""BASELINE configs[4] sweep on one GPU: k in {10,100,1000} x batch in {1,64,128,512,4096}; also bf16 (configs[3])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jsa_rag_b200 as eng
dev = torch.device("cuda:0")
n = int(os.environ.get("DBG_N", 33_000_000))
for dtype in (torch.float16, torch.bfloat16):
    g = torch.Generator(device=dev).manual_seed(1)
    e = torch.empty(n, 768, dtype=dtype, device=dev)
    for s in range(0, n, 1 << 20):
        c = torch.randn(min(1 << 20, n - s), 768, generator=g, device=dev)
        e[s:s + c.shape[0]] = torch.nn.functional.normalize(c, dim=1).to(dtype)
    m = eng.MipsEngine(768, dtype, dev); m.bind(e)
    combos = [(1, 100), (64, 10), (64, 100), (64, 1000), (128, 100), (512, 100), (1024, 100), (4096, 10), (4096, 100), (4096, 1000)]
    if dtype == torch.bfloat16:
        combos = [(64, 100), (512, 100)]
    for b, k in combos:
        q = torch.nn.functional.normalize(torch.randn(b, 768, generator=g, device=dev), dim=1)
        iters = 3 if b >= 512 else 10
        for _ in range(2): m.search(q, k)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(iters): m.search(q, k)
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / iters
        passes = (b + 127) // 128
        print(f"{str(dtype)[6:]:9s} n={n} B={b:5d} k={k:5d}: {ms:9.3f} ms  {b/ms*1e3:9.0f} q/s  {passes} pass(es)  "
              f"{n*1536*passes/ms/1e6:6.0f} GB/s streamed  launches={m.last_launch_count()}", flush=True)
    m.close(); del e
