"""Randomised soak of the search on one B200 (compute-sanitizer is closed on this pool, so wide randomised parity is
the memory-safety evidence): SOAK_TRIALS random (rows, dim, batch, k, dtype, layout, id mapping, flags) cases, each
checked with oracle.compare_topk against fp64 scores of the stored operands and, for batches > 128, for bit identity
with the one-block-per-launch path.  Prints one line per failure and a summary."""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import jsa_rag_b200 as eng
from oracle import flat_index_oracle as O   # checker

dev = torch.device("cuda:0")
seed = int(os.environ.get("SOAK_SEED", 20261018))
trials = int(os.environ.get("SOAK_TRIALS", 150))
rnd = random.Random(seed)
fails, t_start = 0, time.time()
shapes = {}
for trial in range(trials):
    d = rnd.choice([64, 128, 256, 512, 768, 768, 768, 1024])
    n = rnd.choice([1, 2, 63, 64, 65, 127, 1000, 4097, 9473, 40_000, 150_000, 300_001, 1_000_003])
    b = rnd.choice([1, 2, 7, 63, 64, 65, 100, 128, 129, 200, 255, 256, 257, 300, 512, 513, 640, 1025, 1300])
    k = min(n, rnd.choice([1, 2, 10, 20, 100, 100, 128, 129, 300, 1000, 1024]))
    if n * b > 600_000_000:
        b = max(1, 600_000_000 // n)
    dtype = rnd.choice([torch.float16, torch.float16, torch.bfloat16])
    layout_dn = rnd.random() < 0.2
    base, stride = rnd.choice([(0, 1), (3, 8), (5, 2), (1_000_000_007, 1)])
    flags = rnd.choice([0, 0, 0, 0, 64, 128, 512, 4])
    g = torch.Generator(device=dev).manual_seed(seed + trial)
    e = torch.nn.functional.normalize(torch.randn(n, d, generator=g, device=dev), dim=1)
    if rnd.random() < 0.15 and n > 100:          # duplicated rows: exact score ties
        e[n // 2:] = e[:n - n // 2].clone()
    e = e.to(dtype)
    q = torch.nn.functional.normalize(torch.randn(b, d, generator=g, device=dev), dim=1)
    m = eng.MipsEngine(d, dtype, dev)
    try:
        if layout_dn:
            n_pad = (n + 7) // 8 * 8
            e_dn = torch.zeros(d, n_pad, dtype=dtype, device=dev)
            e_dn[:, :n] = e.T
            m.bind(e_dn[:, :n].t(), id_base=base, id_stride=stride)
        else:
            m.bind(e, id_base=base, id_stride=stride)
        m.debug_config(flags)
        s, i = m.search(q, k)
        torch.cuda.synchronize()
        assert bool(((i - base) % stride == 0).all()), "id mapping"
        rows = (i - base) // stride
        exact = (q.to(dtype).double() @ e.double().T)
        order = torch.argsort(-exact, dim=1, stable=True)[:, :k]
        rs = torch.gather(exact, 1, order)
        rep = O.compare_topk(rows.cpu().numpy(), s.cpu().numpy(), order.cpu().numpy(), rs.cpu().numpy(), exact.cpu().numpy(),
                             rtol=1e-5, atol=2e-6)
        assert rep["ok"], rep["errors"][:2]
        if b > 128 and not layout_dn:
            m.debug_config(32)
            s1, i1 = m.search(q, k)
            assert torch.equal(i1, i) and torch.equal(s1, s), "differs from one block per launch"
        key = ("pair" if b > 128 and not layout_dn and not (flags & 128) else "single", "bigk" if k > 128 else "smallk")
        shapes[key] = shapes.get(key, 0) + 1
    except Exception as ex:  # noqa: BLE001
        fails += 1
        print(f"FAIL trial {trial}: n={n} d={d} b={b} k={k} {dtype} dn={layout_dn} ids=({base},{stride}) flags={flags}: {str(ex)[:300]}", flush=True)
    finally:
        m.close()
        del e, q
print(f"soak: {trials} trials, {fails} failures, seed {seed}, {time.time() - t_start:.0f} s, mix {shapes}", flush=True)
sys.exit(1 if fails else 0)
