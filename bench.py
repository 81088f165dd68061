#!/usr/bin/env python
"""bench.py — exact top-k MIPS throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" is one search of a batch of 64 queries (top-100) over the whole 33M x 768 fp16 index.
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
    return ap.parse_args()


def workload(args, world):
    return {"workload": f"{args.rows/1e6:g}M x {args.dim} {args.dtype} index, batch {args.batch}, top-{args.k} exact inner product",
            "rows": args.rows, "dim": args.dim, "batch": args.batch, "k": args.k,
            "sharding": "single shard" if world == 1 else f"round-robin rows over {world} ranks",
            "l2": "index pass (rows*dim*2 bytes per rank) exceeds the 126 MB L2; no flush needed",
            "queries_per_rank": args.batch // world}


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
    return {"value": qps_full, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n_s} of {args.rows} rows (fp16 [dim, n] layout), batch {args.batch}, top-{args.k}; "
                      f"median of {reps} passes, {t*1e3:.1f} ms each; scaled linearly to the full index",
            "seconds_per_sample_pass": t}


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
            "gpu_launches": 0}
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
def run_b200(args):
    import importlib
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
    index.init_embeddings([None] * 0, dim=args.dim)   # doc_map is not exercised by the tensor-level path
    index._store = torch.empty(n_loc, args.dim, dtype=tdtype, device=dev)
    index._set_sharding("round_robin")
    index.equal_batch = True   # every rank contributes batch/N queries
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    for a in range(0, n_loc, 1 << 20):
        c = torch.randn(min(1 << 20, n_loc - a), args.dim, generator=g, device=dev)
        index._store[a:a + c.shape[0]] = torch.nn.functional.normalize(c, dim=1).to(tdtype)
    del c
    q_all = torch.nn.functional.normalize(torch.randn(args.batch, args.dim, device=dev,
                                                      generator=torch.Generator(device=dev).manual_seed(4321)), dim=1)
    per = args.batch // world
    q_mine = q_all[rank * per:(rank + 1) * per].contiguous() if world > 1 else q_all
    q_host = q_mine.cpu().pin_memory()

    engine = index._get_engine()
    dbg = int(os.environ.get("JSA_MIPS_FLAGS", "0"))   # A/B switches of include/jsa_mips.h (0 = product path)
    engine.debug_config(8 | dbg, False)   # flag 8: CUDA events around every full-shard scan launch

    def step():
        return index.search(q_mine, args.k)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

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
        out = step()
    ev1.record()
    sync_all()
    t_wall1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    scan_ms = engine.scan_times_ms()
    p2p = bool(getattr(index, "_xchg", None))
    # own kernels per step: the local search + (push, wait+merge) with the peer exchange, or the merge after NCCL
    p2p_q = bool(getattr(index, "_xchg_q", None))
    launches = engine.last_launch_count() + (((2 if p2p else 1) + (2 if p2p_q else 0)) if world > 1 else 0)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # ---- end to end through the public API with HOST buffers (H2D of the queries + D2H of the result) ----
    res_s = torch.empty(q_mine.shape[0], args.k, dtype=torch.float32).pin_memory()
    res_i = torch.empty(q_mine.shape[0], args.k, dtype=torch.int64).pin_memory()

    graphed = None
    engine.debug_config(dbg, False)   # no event records inside the captured graph
    if world > 1:
        # a ~1 ms distributed search is sensitive to ~150 us of Python/launch overhead per step: replay a CUDA
        # graph of the same public search (collectives included); fall back to the eager call if capture fails
        ok = torch.ones(1, device=dev)
        try:
            graphed = index.make_graphed_search(q_mine.shape[0], args.k)
        except Exception as ex:  # noqa: BLE001
            ok.zero_()
            sys.stderr.write(f"[rank {rank}] CUDA-graph capture unavailable, eager e2e path: {ex}\n")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            if graphed is not None:
                graphed.release()
            graphed = None

    def e2e_step():
        if world == 1:
            engine.search_host(q_host, args.k, out=(res_s, res_i))          # C ABI: mips_search_host
        else:
            if graphed is not None:
                s, i = graphed(q_host)
            else:
                s, i = index.search(q_host.to(dev, non_blocking=True), args.k)
            res_s.copy_(s, non_blocking=True); res_i.copy_(i, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    for _ in range(3):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    sync_all()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
        n_launch = max(1, len(scan_ms))
        scan_avg_ms = sum(scan_ms) / n_launch
        launches_per_step = max(1, len(scan_ms) // max(1, args.steps))
        if args.batch >= 256:
            # several query blocks per launch: the scan is tensor-core bound (BASELINE.md crossover B* ~ 215)
            bound, unit = "tensor", "TFLOP/s"
            algo = 2.0 * args.batch * n_loc * args.dim / launches_per_step            # flops per launch (average)
            achieved = algo / (scan_avg_ms * 1e-3) / 1e12 if scan_ms else None
            if "bf16_tflops_sustained" in peaks:
                peak, peak_src = float(peaks["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
            else:
                peak, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained)"
            algo_key = "algorithmic_flops_per_launch"
        else:
            bound, unit = "hbm", "GB/s"
            algo = float(n_loc * args.dim * 2)                                         # index bytes, read once per launch
            achieved = algo / (scan_avg_ms * 1e-3) / 1e9 if scan_ms else None
            algo_key = "algorithmic_bytes_per_launch"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and args.batch == 64 and args.rows == 33_000_000:
            traffic = json.load(open(tpath)).get(f"n{world}")
        line = {
            "metric": METRIC, "value": args.batch * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f16" if args.dtype == "fp16" else "bf16", "data": "synthetic",
            "config": dict(workload(args, world), **({"exchange": "nvlink peer stores (queries and candidates) + wait-and-merge kernel" if p2p
                                                      else "nccl all-gather + merge kernel"} if world > 1 else {})),
            "e2e": {"value": args.batch * args.steps / e2e_s, "unit": UNIT,
                    "path": "mips_search_host (C ABI, host buffers)" if world == 1 else
                            ("B200Index.make_graphed_search (CUDA-graph replay of the public distributed search)"
                             if graphed is not None else "B200Index.search (public distributed API)"),
                    "h2d_bytes_per_step": int(q_host.numel() * 4 * world),
                    "d2h_bytes_per_step": int((res_s.numel() * 4 + res_i.numel() * 8) * world)},
            "gpu_launches": launches * args.steps,
            "clocks": clocks,
            "roofline": {"bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "kernel": "mips::mips_scan_kernel (full-shard pass)", "peak_source": peak_src,
                         algo_key: algo, "avg_launch_ms": scan_avg_ms, "launches_timed": len(scan_ms),
                         "launches_per_step": launches_per_step},
        }
        if world == 1 and not args.no_cpu_baseline:
            del index._store
            cb = cpu_reference_rate(args, reps=3)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        # the result line is out; a teardown problem must never stall the caller
        guard = threading.Timer(30.0, os._exit, (0,))
        guard.daemon = True
        guard.start()
        if graphed is not None:
            graphed.release()          # graphs that captured NCCL kernels must die before the communicator
        torch.cuda.synchronize()
        index.close_exchange()         # collective: peer-mapped exchange buffers are unmapped before anyone frees
        dist.barrier()
        dist.destroy_process_group()
        guard.cancel()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
