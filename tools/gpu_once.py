"""Two searches (one warm-up) over a synthetic shard: the command profiled under ncu."""
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
for _ in range(int(os.environ.get("DBG_ITERS", 2))):
    s_, i_ = m.search(q, k)
torch.cuda.synchronize()
print("ok", float(s_[0, 0]), int(i_[0, 0]), "launches", m.last_launch_count())
