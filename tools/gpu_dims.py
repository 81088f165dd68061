import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jsa_rag_b200 as eng
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
total_bytes = 12 << 30
for d, dt in [(1024, torch.float16), (1024, torch.bfloat16), (768, torch.float16), (512, torch.float16), (256, torch.float16), (128, torch.float16), (64, torch.float16)]:
    n = total_bytes // (d * 2)
    e = torch.empty(n, d, dtype=dt, device=dev)
    for s in range(0, n, 1 << 21):
        c = torch.randn(min(1 << 21, n - s), d, generator=g, device=dev)
        e[s:s + c.shape[0]] = torch.nn.functional.normalize(c, dim=1).to(dt)
    m = eng.MipsEngine(d, dt, dev); m.bind(e)
    for b in (64, 128):
        q = torch.nn.functional.normalize(torch.randn(b, d, generator=g, device=dev), dim=1)
        for _ in range(3): m.search(q, 100)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10): m.search(q, 100)
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 10
        print(f"dim={d:5d} {str(dt)[6:]:8s} n={n:9d} B={b:3d}: {ms:7.3f} ms  {n*d*2/ms/1e6:6.0f} GB/s  {b/ms*1e3:8.0f} q/s", flush=True)
    m.close(); del e
