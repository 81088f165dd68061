"""Probe: can a distributed search (NCCL all-gathers inside) be captured in a CUDA graph here?  Run under
torchrun with a tight timeout; prints a marker per stage so a hang can be located."""
import faulthandler, importlib, os, sys, signal
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
faulthandler.register(signal.SIGTERM, all_threads=True)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
mode = os.environ.get("PROBE_MODE", "thread_local")
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def say(*a):
    print(f"[r{rank}]", *a, flush=True)
eng = importlib.import_module("jsa-rag_b200")
n, d, k, b = 400_000, 768, 100, 32
g = torch.Generator(device=dev).manual_seed(7)
e = torch.nn.functional.normalize(torch.randn(n, d, generator=g, device=dev), dim=1).half()
q = torch.nn.functional.normalize(torch.randn(b, d, generator=g, device=dev), dim=1)
index = eng.B200Index(); index._store = e[rank::world].contiguous(); index._set_sharding("round_robin"); index.equal_batch = True
s0, i0 = index.search(q, k); torch.cuda.synchronize(); say("eager ok")
static_q = q.clone()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3): index.search(static_q, k)
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize(); dist.barrier(); say("warm-up ok")
graph = torch.cuda.CUDAGraph()
kw = {} if mode == "global" else {"capture_error_mode": mode}
with torch.cuda.graph(graph, **kw):
    gs, gi = index.search(static_q, k)
say("capture ok")
graph.replay(); torch.cuda.synchronize(); say("replay ok", bool(torch.equal(gi, i0)), bool(torch.equal(gs, s0)))
static_q.copy_(torch.roll(q, 1, 0)); graph.replay(); torch.cuda.synchronize()
s1, i1 = index.search(torch.roll(q, 1, 0), k); torch.cuda.synchronize()
say("second replay ok", bool(torch.equal(gi, i1)))
dist.barrier(); say("barrier ok")
if os.environ.get("PROBE_DEL", "1") == "1":
    del graph, gs, gi
    torch.cuda.synchronize(); say("graph deleted")
dist.destroy_process_group(); say("done")
