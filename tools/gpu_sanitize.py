"""Small searches for compute-sanitizer (memcheck / racecheck): every code path once, tiny sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jsa_rag_b200 as eng
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)
for n, d, b, k, dt in [(700, 768, 5, 10, torch.float16), (40000, 768, 70, 100, torch.float16), (9000, 1024, 3, 20, torch.bfloat16),
                       (30000, 768, 4, 600, torch.float16), (148 * 64 * 9, 768, 64, 100, torch.float16)]:
    e = torch.nn.functional.normalize(torch.randn(n, d, generator=g, device=dev), dim=1).to(dt)
    q = torch.nn.functional.normalize(torch.randn(b, d, generator=g, device=dev), dim=1)
    m = eng.MipsEngine(d, dt, dev); m.bind(e)
    s, i = m.search(q, k)
    ref = torch.topk(q.to(dt).float() @ e.float().T, k, dim=1)
    torch.cuda.synchronize()
    print(n, d, b, k, "ok" if float((s - ref.values).abs().max()) < 1e-4 else "MISMATCH", flush=True)
    if k <= 128:
        ms, mi = eng.merge_topk(torch.stack([s, s - 1]), torch.stack([i, i + n]), k)
    m.gather_rows(i[:, :3])
    m.close()
torch.cuda.synchronize()
print("done")
