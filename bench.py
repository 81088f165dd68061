#!/usr/bin/env python
"""bench.py — exact top-k MIPS throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" is one search of a batch of 64 queries (top-100) over the whole 33M x 768 fp16 index
(--searches-per-step 2 = the JSA training step's posterior + prior searches back to back, src/rag.py:1804-1825).
At N=1 the index lives on one B200 (BASELINE configs[1]); at N>1 it is row-sharded round-robin over
the ranks (configs[2]), every rank contributes batch/N queries, and a step is the reference's
distributed search_knn flow: query all-gather -> local fused scan+select -> ONE all-gather of
(score, id) candidates -> device merge.  Total work is fixed as N grows ("scaling": "strong").

Inputs are synthetic: unit-norm random passages / queries, fixed seeds.  The 50.7 GB
index is far larger than the 126 MB L2, so every timed step streams it from HBM.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "queries/sec exact top-100 MIPS, 33M x 768 fp16"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=33_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--searches-per-step", type=int, default=1)
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-api-e2e", action="store_true")
    ap.add_argument("--replicated-queries", action="store_true",
                    help="N > 1: every rank passes the same batch (a front end broadcasting a request to the shards) instead "
                         "of batch/N queries each; implied when the batch does not divide by N")
    ap.add_argument("--sweep", default="", help="comma list of batch[:k[:searches_per_step]] measured over ONE resident index, "
                                                "one JSON line each (BASELINE configs[2..4]); default: the single headline config")
    return ap.parse_args()


def workload(args, world):
    return {"workload": f"{args.rows/1e6:g}M x {args.dim} {args.dtype} index, batch {args.batch}, top-{args.k} exact inner product",
            "rows": args.rows, "dim": args.dim, "batch": args.batch, "k": args.k,
            "sharding": "single shard" if world == 1 else f"round-robin rows over {world} ranks",
            "l2": "index pass (rows*dim*2 bytes per rank) exceeds the 126 MB L2; no flush needed",
            "queries_per_rank": (args.batch // world if args.batch % world == 0 and not getattr(args, "replicated_queries", False)
                                 else f"{args.batch} (the same queries on every rank)"),
            "searches_per_step": args.searches_per_step}


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU arithmetic (torch.matmul fp16 + torch.topk, src/index.py:118-119)
# --------------------------------------------------------------------------------------------------
def cpu_reference_rate(args, reps, warm=1):
    """Times the oracle port on all host threads over a bounded row sample; returns q/s for the FULL index
    (exact scan cost is linear in rows)."""
    import numpy as np
    import torch
    from oracle import flat_index_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_s = min(args.cpu_sample_rows, args.rows)
    g = torch.Generator().manual_seed(1234)
    emb = torch.empty(args.dim, n_s, dtype=torch.float16)
    for a in range(0, n_s, 1 << 18):
        c = torch.randn(min(1 << 18, n_s - a), args.dim, generator=g)
        emb[:, a:a + c.shape[0]] = torch.nn.functional.normalize(c, dim=1).T
    q = torch.nn.functional.normalize(torch.randn(args.batch, args.dim, generator=torch.Generator().manual_seed(4321)), dim=1)
    for _ in range(warm):
        O.cpu_search_arith(q, emb, args.k)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        O.cpu_search_arith(q, emb, args.k)
        times.append(time.perf_counter() - t0)
    t = float(np.median(times))
    qps_full = args.batch / (t * args.rows / n_s)
    # the reference's whole search_knn (doc_map lookups + list building, src/index.py:133-134,152-157) on the same
    # sample: the arithmetic scales with the rows, the host tail (B*k lookups) does not
    pool = [{"id": str(i), "title": f"title {i}", "text": f"passage text {i}"} for i in range(1 << 12)]
    doc_map = _RefPool(pool, n_s)
    sc, ix = O.cpu_search_arith(q, emb, args.k)
    tails = []
    for _ in range(3):
        t0 = time.perf_counter()
        O.search_knn_tail(sc, ix, doc_map, args.k)
        tails.append(time.perf_counter() - t0)
    tail = float(np.median(tails))
    return {"value": qps_full, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n_s} of {args.rows} rows (fp16 [dim, n] layout), batch {args.batch}, top-{args.k}; "
                      f"median of {reps} passes, {t*1e3:.1f} ms each; scaled linearly to the full index",
            "seconds_per_sample_pass": t,
            "search_knn": {"value": args.batch / (t * args.rows / n_s + tail), "unit": UNIT,
                           "host_tail_ms": tail * 1e3,
                           "what": "reference search_knn incl. doc_map lookups and list building (src/index.py:123-158): "
                                   "arithmetic scaled to the full index + the measured, size-independent host tail"}}


class _RefPool:
    """doc_map stand-in for timing: n keys that map onto a small pool of passage dicts (a 33M-entry dict of dicts would
    take minutes and tens of GB to build; the lookups cost the same)."""

    def __init__(self, pool, n):
        self.pool, self.n = pool, n

    def __getitem__(self, i):
        return self.pool[i & (len(self.pool) - 1)]

    def __len__(self):
        return self.n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_rate(args, reps=max(1, args.steps), warm=max(1, min(args.warmup, 2)))
    world = args.gpus
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * args.batch / cb["value"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic", "impl": "reference", "config": workload(args, world),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "api_e2e": cb["search_knn"], "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [ln for (t, ln) in self.lines if t0 <= t <= t1 + 0.2] or [ln for (_, ln) in self.lines[-3:]]
        sm, mx, reasons = [], [], set()
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------------
class PooledDocMap:
    """doc_map for the api_e2e leg: n_local keys -> references into a pool of 2^17 synthetic passage dicts.  The host
    tail of search_knn (object gather, list building, score rounding) costs what it costs with distinct dicts; building
    33M distinct dicts would add minutes and tens of GB to a benchmark that does not read their text."""

    def __init__(self, n, pool_bits=17):
        self.n = n
        self.pool = [{"id": str(i), "title": f"title {i}", "text": f"synthetic passage {i}"} for i in range(1 << pool_bits)]

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return self.pool[i & (len(self.pool) - 1)]

    def as_object_array(self):
        import numpy as np
        tab = np.empty(len(self.pool), dtype=object)
        tab[:] = self.pool
        return tab[np.arange(self.n, dtype=np.int64) & (len(self.pool) - 1)]


def parity_checks(eng, index, args, world, rank, dev, tdtype, q_sets, repl=False):
    """Correctness evidence carried by the bench line (outside every timed region): the searches that were timed are
    checked against planted nearest neighbours on the full index, against the NCCL all-gather + merge path bit for
    bit (N > 1), and against the reference arithmetic (oracle, CPU) on a sub-shard with oracle.compare_topk."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from oracle import flat_index_oracle as O   # checker only

    checked, ok, notes = [], True, {}
    n_loc = int(index._store.shape[0])
    per = q_sets[0].shape[0]
    k = args.k

    # ---- (1) planted nearest neighbours over the full index: query = own passage row + 2 % noise
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    rows = torch.randint(0, n_loc, (per,), generator=g, device=dev)
    qp = index._store[rows].float()
    qp = torch.nn.functional.normalize(qp + 0.02 * torch.randn(qp.shape, generator=g, device=dev) / (args.dim ** 0.5), dim=1)
    want = index._id_base + rows * index._id_stride
    if repl:                                   # the same queries on every rank: rank 0's plants
        dist.broadcast(qp, 0)
        dist.broadcast(want, 0)
    s, i = index.search(qp, k, replicated=repl)
    planted_ok = bool((i[:, 0] == want).all()) and bool((s[:, 0] > 0.99).all()) and bool((s[:, 1:] <= s[:, :-1]).all())
    checked.append("planted nearest neighbours (full index, every rank's queries)")
    ok = ok and planted_ok

    # ---- (2) peer-exchange path == NCCL all-gather + merge path, bit for bit
    if world > 1:
        ref_s, ref_i = index.search(q_sets[0], k, replicated=repl)
        ref_s, ref_i = ref_s.clone(), ref_i.clone()
        saved = (getattr(index, "_xchg", None), getattr(index, "_xchg_q", None))
        if saved[0]:
            index._xchg, index._xchg_q = False, False
            n_s, n_i = index.search(q_sets[0], k, replicated=repl)
            index._xchg, index._xchg_q = saved
            same = torch.equal(n_s, ref_s) and torch.equal(n_i, ref_i)
            checked.append("nvlink peer exchange == nccl all-gather + merge (bit-identical scores and ids)")
            ok = ok and same
        else:
            notes["exchange"] = "nccl path only (no peer mapping): nothing to compare"

    # ---- (3) oracle on a sub-shard: the first rows of every rank's shard form a small index with the same sharding
    sub_total = int(min(2_000_000, max(100_000, (1 << 27) // max(1, args.batch))))
    sub_n = max(k, min(n_loc, sub_total // world))
    if world > 1:
        t = torch.tensor([sub_n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        sub_n = int(t.item())
    sub = eng.B200Index(dtype=tdtype)
    sub._store = index._store[:sub_n]
    sub._set_sharding("round_robin")
    sub.equal_batch = index.equal_batch
    ss, si = sub.search(q_sets[0], k, replicated=repl)
    if world > 1:
        all_e = torch.empty((world, sub_n, args.dim), dtype=tdtype, device=dev)
        dist.all_gather_into_tensor(all_e, sub._store.contiguous())
        if repl:
            qq = q_sets[0]                                                   # every rank already holds all rows
        else:
            sizes = eng.dist_utils.get_varsize(ss)
            ss = eng.dist_utils.varsize_all_gather(ss.contiguous(), sizes)      # rank order = query order
            si = eng.dist_utils.varsize_all_gather(si.contiguous(), sizes)
            qq = eng.dist_utils.varsize_all_gather(q_sets[0].contiguous(), sizes)
        emb = all_e.permute(1, 0, 2).reshape(sub_n * world, args.dim)     # global id = local * W + rank
        sub.close_exchange()
    else:
        qq, emb = q_sets[0], sub._store
    if rank == 0:
        exact = (qq.to(tdtype).float() @ emb.float().T).cpu().numpy()      # exact products, fp32 accumulation
        e_cpu = emb.cpu()
        e_dn = e_cpu.t().contiguous() if tdtype == torch.bfloat16 else O.make_embeddings_dn(e_cpu)
        torch.set_num_threads(os.cpu_count() or 1)
        r_s, r_i = O.compute_scores_and_indices(qq.cpu(), e_dn, k)         # src/index.py:118-119 on the host
        # north_star's tolerance (1e-3 relative) is stated against the reference's fp16 scores; the bf16 extension
        # returns bf16-rounded scores (8-bit mantissa: half an ulp is 2^-8 = 3.9e-3 relative), hence 4e-3 there
        rtol = 1e-3 if tdtype == torch.float16 else 4e-3
        rep = O.compare_topk(si.cpu().numpy(), ss.cpu().numpy(), r_i.numpy(), r_s.float().numpy(), exact, rtol=rtol)
        checked.append(f"oracle.compare_topk (rtol {rtol:g}) vs the reference arithmetic on a {sub_n * world}-row sub-index "
                       f"({rep['id_set_equal_rows']}/{rep['rows']} rows with identical id sets, {rep['near_tie_diffs']} near-tie differences)")
        ok = ok and rep["ok"]
        if not rep["ok"]:
            notes["oracle_errors"] = [str(e) for e in rep["errors"][:3]]
    if world > 1:
        t = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item() == 1.0)
    out = {"checked": checked, "ok": ok}
    out.update(notes)
    return out


def run_b200(args):
    import importlib
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = importlib.import_module("jsa-rag_b200")
    if not eng._native.is_built():
        importlib.import_module("jsa-rag_b200.build").build()

    tdtype = torch.float16 if args.dtype == "fp16" else torch.bfloat16
    n_loc = len(range(rank, args.rows, world))
    # ---- synthetic shard: rows rank, rank+W, ... of the global index (src/index_io.py:41 sharding) ----
    index = eng.B200Index(dtype=tdtype)
    index.init_embeddings([None] * 0, dim=args.dim)   # doc_map is set for the api_e2e leg only
    index._store = torch.empty(n_loc, args.dim, dtype=tdtype, device=dev)
    index._set_sharding("round_robin")
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    for a in range(0, n_loc, 1 << 20):
        c = torch.randn(min(1 << 20, n_loc - a), args.dim, generator=g, device=dev)
        index._store[a:a + c.shape[0]] = torch.nn.functional.normalize(c, dim=1).to(tdtype)
    del c
    def measure(args, first):
        sps = max(1, args.searches_per_step)
        # every rank contributes batch/N queries (reference training flow); a batch that does not divide by N is
        # searched the way a serving front end would: the same queries on every rank, no query exchange
        repl = world > 1 and (args.replicated_queries or args.batch % world != 0)
        sizes = [args.batch] * world if repl else [args.batch // world] * world
        offs = [0] * (world + 1) if repl else [sum(sizes[:r]) for r in range(world + 1)]
        per = sizes[rank]
        index.equal_batch = True
        q_sets, q_hosts = [], []
        for j in range(sps):     # one query set per search of a step (posterior / prior queries differ)
            q_all = torch.nn.functional.normalize(torch.randn(args.batch, args.dim, device=dev,
                                                              generator=torch.Generator(device=dev).manual_seed(4321 + j)), dim=1)
            q_mine = q_all[offs[rank]:offs[rank + 1]].contiguous() if (world > 1 and not repl) else q_all
            q_sets.append(q_mine)
            q_hosts.append(q_mine.cpu().pin_memory())

        engine = index._get_engine()
        dbg = int(os.environ.get("JSA_MIPS_FLAGS", "0"))   # A/B switches of include/jsa_mips.h (0 = product path)
        engine.debug_config(8 | dbg, False)   # flag 8: CUDA events around every full-shard scan launch

        def step():
            for q in q_sets:
                out = index.search(q, args.k, replicated=repl)
            return out

        def sync_all():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()

        def rank_max(x):
            if world == 1:
                return x
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        for _ in range(max(args.warmup, 3)):
            step()
        sync_all()
        engine.scan_times_ms()          # drop warm-up launches
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        t_wall0 = time.perf_counter()
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        sync_all()
        t_wall1 = time.perf_counter()
        ms = rank_max(ev0.elapsed_time(ev1))
        scan_ms = engine.scan_times_ms()
        p2p = bool(getattr(index, "_xchg", None))
        # own kernels per step: the local search + (push, wait+merge) with the peer exchange, or the merge after NCCL
        p2p_q = bool(getattr(index, "_xchg_q", None))
        launches = sps * (engine.last_launch_count() + (((2 if p2p else 1) + (2 if p2p_q else 0)) if world > 1 else 0))
        clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

        # ---- sustained load: the same step back to back for >= sustained_seconds (the timed region above is a burst
        #      of a fraction of a second; a training loop sits at the 1 kW power cap) ----
        sustained = None
        if args.sustained_seconds > 0:
            n_sus = max(args.steps, int(args.sustained_seconds * 1e3 / max(1e-3, ms / args.steps)) + 1)
            if world > 1:
                t = torch.tensor([n_sus], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                n_sus = int(t.item())
            sampler2 = ClockSampler(local_rank)
            if rank == 0:
                sampler2.start()
            sync_all()
            engine.scan_times_ms()
            w0 = time.perf_counter()
            ev0.record()
            for it in range(n_sus):
                step()
                if it % 64 == 63:
                    engine.scan_times_ms()          # keep only the tail of the run (at most 256 launches are recorded)
            ev1.record()
            sync_all()
            w1 = time.perf_counter()
            sus_ms = rank_max(ev0.elapsed_time(ev1))
            sus_scan = engine.scan_times_ms()
            sus_clocks = sampler2.stop(w0 + 0.5 * (w1 - w0), w1) if rank == 0 else None   # second half: settled clocks
            sustained = {"steps": n_sus, "seconds": sus_ms * 1e-3, "value": args.batch * sps * n_sus / (sus_ms * 1e-3),
                         "ms_per_step": sus_ms / n_sus, "avg_launch_ms": (sum(sus_scan) / len(sus_scan)) if sus_scan else None,
                         "clocks": sus_clocks}

        # ---- end to end through the public API with HOST buffers (H2D of the queries + D2H of the result) ----
        res_s = torch.empty(per, args.k, dtype=torch.float32).pin_memory()
        res_i = torch.empty(per, args.k, dtype=torch.int64).pin_memory()

        graphed = None
        engine.debug_config(dbg, False)   # no event records inside the captured graph
        if world > 1 and index.equal_batch:
            # a ~1 ms distributed search is sensitive to ~150 us of Python/launch overhead per step: replay a CUDA
            # graph of the same public search (collectives included); fall back to the eager call if capture fails
            okf = torch.ones(1, device=dev)
            try:
                graphed = index.make_graphed_search(per, args.k, replicated=repl)
            except Exception as ex:  # noqa: BLE001
                okf.zero_()
                sys.stderr.write(f"[rank {rank}] CUDA-graph capture unavailable, eager e2e path: {ex}\n")
            dist.all_reduce(okf, op=dist.ReduceOp.MIN)
            if okf.item() == 0:
                if graphed is not None:
                    graphed.release()
                graphed = None

        def e2e_step():
            for qh in q_hosts:
                if world == 1:
                    engine.search_host(qh, args.k, out=(res_s, res_i))          # C ABI: mips_search_host
                else:
                    if graphed is not None:
                        s, i = graphed(qh)
                    else:
                        s, i = index.search(qh.to(dev, non_blocking=True), args.k, replicated=repl)
                    res_s.copy_(s, non_blocking=True); res_i.copy_(i, non_blocking=True)
                    torch.cuda.current_stream().synchronize()

        graphed_used = graphed is not None
        for _ in range(3):
            e2e_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        sync_all()
        e2e_s = rank_max(time.perf_counter() - t0)
        if graphed is not None:
            graphed.release()          # graphs that captured NCCL kernels must die before the communicator
            graphed = None

        # ---- the reference-facing call itself: search_knn -> (docs, scores) as nested Python lists ----
        api = None
        if not args.no_api_e2e and not repl:      # search_knn is the per-rank-queries call of the reference
            if not isinstance(index.doc_map, PooledDocMap):
                index.doc_map = PooledDocMap(n_loc)
                index.refresh_passages()
            q_dev = q_sets[0]
            for _ in range(2):
                index.search_knn(q_dev, args.k)
            sync_all()
            n_api = max(3, min(args.steps, 20))
            t0 = time.perf_counter()
            for _ in range(n_api):
                docs, scores = index.search_knn(q_dev, args.k)
            sync_all()
            api_s = rank_max(time.perf_counter() - t0)
            t0 = time.perf_counter()
            for _ in range(n_api):
                index.search(q_dev, args.k)
            sync_all()
            dev_s = rank_max(time.perf_counter() - t0)
            assert len(docs) == q_dev.shape[0] and (not docs or (len(docs[0]) == args.k and isinstance(scores[0][0], float)))
            api = {"value": args.batch * n_api / api_s, "unit": UNIT, "ms_per_search": 1e3 * api_s / n_api,
                   "host_tail_ms": 1e3 * (api_s - dev_s) / n_api,
                   "what": "B200Index.search_knn (reference signature, src/index.py:123-158): device search + D2H + passage "
                           "dicts for the k winners of this rank's queries + fp16-rounded score lists",
                   "passages": getattr(index, "last_passage_path", None)}

        parity = None if args.no_parity else parity_checks(eng, index, args, world, rank, dev, tdtype, q_sets, repl)

        if rank == 0:
            peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
            peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
            n_launch = max(1, len(scan_ms))
            scan_avg_ms = sum(scan_ms) / n_launch
            launches_per_step = max(1, len(scan_ms) // max(1, args.steps))
            if args.batch >= 256:
                # several query blocks per launch: the scan is tensor-core bound (BASELINE.md crossover B* ~ 215)
                bound, unit = "tensor", "TFLOP/s"
                per_launch = 2.0 * args.batch * sps * n_loc * args.dim / launches_per_step   # flops per launch (average)
                scale = 1e12
                if "bf16_tflops_sustained" in peaks:
                    peak, peak_src = float(peaks["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
                else:
                    peak, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained)"
                algo_key = "algorithmic_flops_per_launch"
            else:
                bound, unit = "hbm", "GB/s"
                per_launch = float(n_loc * args.dim * 2)                                      # index bytes, read once per launch
                scale = 1e9
                if "hbm_gbs" in peaks:
                    peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
                else:
                    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
                algo_key = "algorithmic_bytes_per_launch"
            achieved = per_launch / (scan_avg_ms * 1e-3) / scale if scan_ms else None
            traffic, traffic_src = None, None
            tpath = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tpath):
                tj = json.load(open(tpath))
                ent = tj.get("per_launch", {}).get(f"rows{n_loc}_dim{args.dim}_b{args.batch}_k{args.k}_{args.dtype}")
                if ent:
                    traffic, traffic_src = ent["dram_bytes"], ent["source"]
            roof = {"bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                    "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                    "kernel": "mips::mips_scan_pair_kernel (full-shard pass, CTA pairs)" if args.batch > 128 and not (dbg & (128 | 32))
                              else "mips::mips_scan_kernel (full-shard pass)",
                    "peak_source": peak_src, algo_key: per_launch, "avg_launch_ms": scan_avg_ms,
                    "launches_timed": len(scan_ms), "launches_per_step": launches_per_step}
            if sustained is not None:
                sa = per_launch / (sustained["avg_launch_ms"] * 1e-3) / scale if sustained["avg_launch_ms"] else None
                roof["sustained"] = {"achieved": sa, "frac": (sa / peak) if sa else None, "unit": unit,
                                     "avg_launch_ms": sustained["avg_launch_ms"], "seconds": sustained["seconds"],
                                     "steps": sustained["steps"], "value": sustained["value"], "value_unit": UNIT,
                                     "sm_mhz": (sustained["clocks"] or {}).get("sm_mhz"),
                                     "reasons": (sustained["clocks"] or {}).get("reasons")}
            line = {
                "metric": METRIC, "value": args.batch * sps * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f16" if args.dtype == "fp16" else "bf16", "data": "synthetic",
                "config": workload(args, world),       # identical to the reference arm's config object
                "exchange": (("nvlink peer stores (queries and candidates) + wait-and-merge kernel" if p2p
                              else "nccl all-gather + merge kernel") if world > 1 else None),
                "e2e": {"value": args.batch * sps * args.steps / e2e_s, "unit": UNIT,
                        "path": "mips_search_host (C ABI, host buffers)" if world == 1 else
                                ("B200Index.make_graphed_search (CUDA-graph replay of the public distributed search)"
                                 if graphed_used else "B200Index.search (public distributed API, eager)"),
                        "h2d_bytes_per_step": int(args.batch * args.dim * 4 * sps * (world if repl else 1)),
                        "d2h_bytes_per_step": int(args.batch * args.k * 12 * sps * (world if repl else 1))},
                "api_e2e": api,
                "gpu_launches": launches * args.steps,
                "clocks": clocks,
                "roofline": roof,
                "parity": parity,
            }
            if world == 1 and not args.no_cpu_baseline and first:
                cb = cpu_reference_rate(args, reps=3)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "search_knn")}
            else:
                line["cpu_baseline"] = None
            print(json.dumps(line), flush=True)

    configs = [(args.batch, args.k, args.searches_per_step)]
    if args.sweep:
        configs = []
        for item in args.sweep.split(","):
            f = [int(x) for x in item.split(":")]
            configs.append((f[0], f[1] if len(f) > 1 else args.k, f[2] if len(f) > 2 else 1))
    import copy
    for ci, (b_, k_, sps_) in enumerate(configs):
        a = copy.copy(args)
        a.batch, a.k, a.searches_per_step = b_, k_, sps_
        measure(a, ci == 0)
    if world > 1:
        # the result line is out; a teardown problem must never stall the caller
        guard = threading.Timer(30.0, os._exit, (0,))
        guard.daemon = True
        guard.start()
        torch.cuda.synchronize()
        index.close_exchange()         # collective: peer-mapped exchange buffers are unmapped before anyone frees
        dist.barrier()
        dist.destroy_process_group()
        guard.cancel()


def main():
    args = parse_args()
    if os.environ.get("JSA_BENCH_WATCHDOG_S"):
        # debugging aid: dump every thread's Python stack and exit if the run is still going after that many seconds
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["JSA_BENCH_WATCHDOG_S"]), exit=True)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
