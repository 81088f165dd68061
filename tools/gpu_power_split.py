"""Where does the power go?  Sustained back-to-back scans (33M x 768, B=64) with parts of the kernel switched off:
full search, no select epilogue (flag 1), no MMAs = pure TMA streaming (flag 2).  Prints ms/search, GB/s and the
SM clock / power sampled while the loop runs."""
import sys, os, threading, time
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
q = torch.nn.functional.normalize(torch.randn(64, 768, generator=g, device=dev), dim=1)


def sample(stop, out):
    while not stop.is_set():
        out.append(os.popen("nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits").read().strip())
        time.sleep(0.1)


for name, flags in (("full", 0), ("no-select", 1), ("no-mma", 2), ("no-mma,no-select", 3), ("full", 0)):
    m.debug_config(flags, False)
    for _ in range(5): m.search(q, k)
    torch.cuda.synchronize()
    stop, smp = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, smp)); th.start()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    iters = 120
    t0.record()
    for _ in range(iters): m.search(q, k)
    t1.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    ms = t0.elapsed_time(t1) / iters
    tail = smp[len(smp) // 2:]
    clk = sorted(float(x.split(",")[0]) for x in tail)[len(tail) // 2]
    pw = sorted(float(x.split(",")[1]) for x in tail)[len(tail) // 2]
    print(f"{name:18s}: {ms:.3f} ms/search  {n*1536/ms/1e6:.0f} GB/s   median SM clock {clk:.0f} MHz, power {pw:.0f} W", flush=True)
