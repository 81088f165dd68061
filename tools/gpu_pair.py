"""CTA-pair (cta_group::2) scan on a B200: ids bit-identical to the one-block-per-launch path (debug flag 32) and to
torch fp32 on the stored operands, then timing at the full index against the round-1 multi-block path (flag 128).
DBG_N (default 33M) sets the timed index size, DBG_QUICK=1 skips the timing."""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jsa_rag_b200 as eng

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
print(torch.cuda.get_device_name(0), flush=True)


def synth(n, d, b, seed, dtype):
    g = torch.Generator(device=dev).manual_seed(seed)
    e = torch.empty(n, d, dtype=dtype, device=dev)
    for s in range(0, n, 1 << 20):
        c = torch.randn(min(1 << 20, n - s), d, generator=g, device=dev)
        e[s:s + c.shape[0]] = torch.nn.functional.normalize(c, dim=1).to(dtype)
    q = torch.nn.functional.normalize(torch.randn(b, d, generator=g, device=dev), dim=1)
    return e, q


def check(n, d, b, k, dtype=torch.float16, seed=0):
    e, q = synth(n, d, b, seed, dtype)
    m = eng.MipsEngine(d, dtype, dev)
    m.bind(e)
    s, i = m.search(q, k)
    launches = m.last_launch_count()
    torch.cuda.synchronize()
    m.debug_config(32)                       # one query block per launch, no pairs
    s1, i1 = m.search(q, k)
    torch.cuda.synchronize()
    m.debug_config(0)
    ref = q.to(dtype).float() @ e.float().T
    rs, ri = torch.topk(ref, k, dim=1)
    got_exact = torch.gather(ref, 1, i.clamp(0, n - 1))
    err = (s - got_exact).abs().max().item()
    set_eq = sum(set(a.tolist()) == set(c.tolist()) for a, c in zip(i, ri))
    same = bool(torch.equal(i, i1)) and bool(torch.equal(s, s1))
    print(f"n={n} d={d} b={b} k={k} {str(dtype)[6:]}: bit-identical to flag32={same} score_err={err:.2e} "
          f"idset_equal_torch={set_eq}/{b} sorted={bool((s[:, 1:] <= s[:, :-1]).all())} launches={launches}", flush=True)
    if not same:
        bad = (i != i1).any(dim=1).nonzero().flatten()[:4].tolist()
        print("  first differing queries:", bad)
        for r in bad[:2]:
            print("   pair ", i[r, :8].tolist(), [round(x, 4) for x in s[r, :8].tolist()])
            print("   flag32", i1[r, :8].tolist(), [round(x, 4) for x in s1[r, :8].tolist()])
    m.close()
    return same and err < 1e-3


ok = True
cases = [(20000, 768, 256, 100), (20000, 768, 129, 10), (20000, 768, 200, 100), (50000, 768, 512, 100),
         (50000, 768, 300, 20), (50000, 768, 700, 100), (31, 768, 256, 5), (100, 768, 512, 100), (4097, 768, 1024, 128),
         (30000, 1024, 512, 100), (30000, 64, 512, 100), (30000, 512, 384, 100), (40000, 768, 512, 100, torch.bfloat16),
         (60000, 768, 512, 1000), (400000, 768, 512, 100), (300000, 768, 1500, 100), (300000, 768, 1024, 10),
         (2_000_000, 768, 512, 100)]
for c in cases:
    try:
        ok = check(*c) and ok
    except Exception:
        traceback.print_exc()
        print("FAILED case", c, flush=True)
        ok = False
        break
print("PAIR PARITY", "OK" if ok else "FAILED", flush=True)
if not ok or os.environ.get("DBG_QUICK"):
    sys.exit(0 if ok else 1)

n = int(os.environ.get("DBG_N", 33_000_000))
for dtype in (torch.float16, torch.bfloat16):
    e, _ = synth(n, 768, 1, 1, dtype)
    m = eng.MipsEngine(768, dtype, dev)
    m.bind(e)
    for b in ((192, 256, 512, 1024) if dtype == torch.float16 else (512,)):
        q = torch.nn.functional.normalize(torch.randn(b, 768, device=dev), dim=1)
        for flags, name in ((0, "pairs"), (128, "r1 multi-block")):
            m.debug_config(8 | flags)
            for _ in range(2):
                m.search(q, 100)
            torch.cuda.synchronize()
            m.scan_times_ms()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            iters = 5
            t0.record()
            for _ in range(iters):
                m.search(q, 100)
            t1.record(); torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / iters
            scans = m.scan_times_ms()
            tf = 2.0 * b * n * 768 / (ms * 1e-3) / 1e12
            print(f"{str(dtype)[6:]:8s} n={n} B={b:5d} {name:15s}: {ms:8.3f} ms/search {b/ms*1e3:8.0f} q/s {tf:7.1f} TFLOP/s "
                  f"({tf/1405.9:.3f} of sustained peak) scan launches/search={len(scans)//iters} "
                  f"scan sum={sum(scans)/iters:.3f} ms launches={m.last_launch_count()}", flush=True)
    m.close()
    del e
