"""Developer smoke/debug on a B200: correctness vs torch fp32 on the same device + a first timing."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jsa_rag_b200 as eng

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)


def make(n, d, b, seed, dtype=torch.float16):
    g = torch.Generator(device=dev).manual_seed(seed)
    e = torch.nn.functional.normalize(torch.randn(n, d, generator=g, device=dev), dim=1).to(dtype)
    q = torch.nn.functional.normalize(torch.randn(b, d, generator=g, device=dev), dim=1)
    return e, q


def check(n, d, b, k, dtype=torch.float16, seed=0, verbose=False):
    e, q = make(n, d, b, seed, dtype)
    m = eng.MipsEngine(d, dtype, dev)
    m.bind(e)
    s, i = m.search(q, k)
    torch.cuda.synchronize()
    ref = q.to(dtype).float() @ e.float().T
    rs, ri = torch.topk(ref, k, dim=1)
    got_exact = torch.gather(ref, 1, i.clamp(0, n - 1))
    err = (s - got_exact).abs().max().item()
    set_eq = sum(set(a.tolist()) == set(c.tolist()) for a, c in zip(i, ri))
    kth_gap = (s[:, -1] - rs[:, -1]).abs().max().item()
    print(f"n={n} d={d} b={b} k={k} {dtype}: score_err={err:.3e} kth_gap={kth_gap:.3e} idset_equal={set_eq}/{b} "
          f"sorted={bool((s[:, 1:] <= s[:, :-1]).all())} idrange=({i.min().item()},{i.max().item()})", flush=True)
    if verbose or err > 1e-3:
        print(" engine s[0,:8]", s[0, :8].tolist()); print(" ref    s[0,:8]", rs[0, :8].tolist())
        print(" engine i[0,:8]", i[0, :8].tolist()); print(" ref    i[0,:8]", ri[0, :8].tolist())
    m.close()
    return err


cases = [(128, 768, 64, 128), (128, 768, 3, 16), (300, 768, 5, 10), (1000, 1024, 7, 20), (20000, 768, 64, 100),
         (20000, 768, 100, 100), (148 * 128 * 3 + 17, 768, 64, 100), (50000, 768, 64, 100, torch.bfloat16)]
for c in cases:
    try:
        check(*c, verbose=(c[0] == 128))
    except Exception:
        traceback.print_exc()
        print("FAILED case", c, flush=True)
        break

# ties: all-zero index (init_embeddings state) -> rows 0..k-1 with score 0; duplicated rows -> smallest ids win
try:
    z = torch.zeros(5000, 768, dtype=torch.float16, device=dev)
    m = eng.MipsEngine(768, torch.float16, dev); m.bind(z)
    q = torch.randn(7, 768, device=dev)
    s_, i_ = m.search(q, 20); torch.cuda.synchronize()
    print("zero index ok:", bool((i_ == torch.arange(20, device=dev)).all()), bool((s_ == 0).all()), flush=True)
    e, q = make(3000, 768, 9, 5)
    e2 = e.repeat(40, 1)            # every row appears 40 times: ids r, r+3000, ...
    m.bind(e2)
    s_, i_ = m.search(q, 100); torch.cuda.synchronize()
    ref = (q.half().float() @ e2.float().T)
    order = torch.argsort(-ref.double() * 1e6 + torch.arange(e2.shape[0], device=dev).double() * 1e-9, dim=1)[:, :100]
    rs = torch.gather(ref, 1, order)
    exact_ok = bool((torch.gather(ref, 1, i_) - s_).abs().max() < 1e-5)
    print("dup index: ids equal to (score desc, id asc) order:", bool((order == i_).all()), "scores ok", exact_ok, flush=True)
    m.close()
except Exception:
    traceback.print_exc()

# timing
try:
    n = int(os.environ.get("DBG_N", 4_000_000))
    e, q = make(n, 768, 64, 1)
    m = eng.MipsEngine(768, torch.float16, dev); m.bind(e)
    for _ in range(3): m.search(q, 100)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    iters = 10
    for _ in range(iters): m.search(q, 100)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / iters
    print(f"timing n={n}: {ms:.3f} ms/search  {n*1536/ms/1e6:.1f} GB/s  {64/ms*1e3:.0f} q/s launches={m.last_launch_count()}", flush=True)
    # torch baseline on GPU
    qh = q.half(); et = e.t().contiguous()
    for _ in range(2): torch.topk(qh @ et, 100, dim=1)
    torch.cuda.synchronize(); t0.record()
    for _ in range(3): torch.topk(qh @ et, 100, dim=1)
    t1.record(); torch.cuda.synchronize()
    print(f"torch matmul+topk same shape: {t0.elapsed_time(t1)/3:.3f} ms", flush=True)
except Exception:
    traceback.print_exc()
