import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jsa_rag_b200 as eng
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
for t in (1, 2, 4, 7):
    n = 148 * 64 * t
    e = torch.nn.functional.normalize(torch.randn(n, 768, generator=g, device=dev), dim=1).half()
    q = torch.nn.functional.normalize(torch.randn(64, 768, generator=g, device=dev), dim=1)
    m = eng.MipsEngine(768, torch.float16, dev); m.bind(e)
    for _ in range(3): m.search(q, 100)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(50): m.search(q, 100)
    t1.record(); torch.cuda.synchronize()
    print(f"tiles/CTA={t} n={n}: {t0.elapsed_time(t1)/50*1e3:.1f} us per search ({m.last_launch_count()} launches)", flush=True)
