"""Large-batch scan under sustained load: ms/search, TFLOP/s, SM clock and power sampled while the loop runs, and the
per-role cycle counters of one launch (who waits for whom).  DBG_B (512), DBG_N (33M), DBG_FLAGS (0; 128 = no pairs),
DBG_DTYPE (fp16|bf16), DBG_SECS (seconds of back-to-back searches, 3)."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jsa_rag_b200 as eng

dev = torch.device("cuda:0")
n = int(os.environ.get("DBG_N", 33_000_000)); k = int(os.environ.get("DBG_K", 100))
dtype = torch.bfloat16 if os.environ.get("DBG_DTYPE", "fp16") == "bf16" else torch.float16
g = torch.Generator(device=dev).manual_seed(1)
e = torch.empty(n, 768, dtype=dtype, device=dev)
for s in range(0, n, 1 << 20):
    c = torch.randn(min(1 << 20, n - s), 768, generator=g, device=dev)
    e[s:s + c.shape[0]] = torch.nn.functional.normalize(c, dim=1).to(dtype)
m = eng.MipsEngine(768, dtype, dev); m.bind(e)


def sample(stop, out):
    while not stop.is_set():
        out.append(os.popen("nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits").read().strip())
        time.sleep(0.1)


for b in [int(x) for x in os.environ.get("DBG_B", "512").split(",")]:
    q = torch.nn.functional.normalize(torch.randn(b, 768, generator=g, device=dev), dim=1)
    for flags in [int(x) for x in os.environ.get("DBG_FLAGS", "0").split(",")]:
        m.debug_config(flags, False)
        for _ in range(3): m.search(q, k)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); m.search(q, k); t1.record(); torch.cuda.synchronize()
        one = t0.elapsed_time(t1)
        iters = max(3, int(float(os.environ.get("DBG_SECS", 3)) * 1e3 / one))
        stop, smp = threading.Event(), []
        th = threading.Thread(target=sample, args=(stop, smp)); th.start()
        t0.record()
        for _ in range(iters): m.search(q, k)
        t1.record(); torch.cuda.synchronize()
        stop.set(); th.join()
        ms = t0.elapsed_time(t1) / iters
        tail = smp[len(smp) // 2:] or smp
        clk = sorted(float(x.split(",")[0]) for x in tail)[len(tail) // 2]
        pw = sorted(float(x.split(",")[1]) for x in tail)[len(tail) // 2]
        tf = 2.0 * b * n * 768 / (ms * 1e-3) / 1e12
        print(f"B={b} flags={flags} {str(dtype)[6:]}: {ms:.3f} ms/search ({iters} back to back) {b/ms*1e3:.0f} q/s {tf:.1f} TFLOP/s "
              f"= {tf/1405.9:.3f} of sustained peak, {n*1536/ms/1e6:.0f} GB/s index; median SM clock {clk:.0f} MHz, power {pw:.0f} W; "
              f"tensor-pipe share at that clock {tf*1e12/(148*8192*clk*1e6):.3f}", flush=True)
        st = m.debug_config(flags, True)
        m.search(q, k); torch.cuda.synchronize()
        stf = st.double()
        act = stf[:, 8] > 0
        print("   per-CTA mean:", {nm: round(v, 0) for nm, v in zip(m.STAT_NAMES, stf[act].mean(0).tolist())})
        print("   per-CTA max :", {nm: round(v, 0) for nm, v in zip(m.STAT_NAMES, stf[act].max(0).values.tolist())}, flush=True)
m.debug_config(0, False)
