"""One search per shard size for an ncu pass that reads the scan kernel's DRAM bytes:
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \\
      -k regex:mips_scan --csv --log-file gpurun_out/traffic.csv python tools/gpu_traffic.py
TRAFFIC_CASES = "rows:batch:k:dtype,..." (default: the per-rank shards of the 33M index at 1/2/4/8 GPUs, batch 64 and 512)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jsa_rag_b200 as eng

dev = torch.device("cuda:0")
cases = os.environ.get("TRAFFIC_CASES", "33000000:64:100:fp16,16500000:64:100:fp16,8250000:64:100:fp16,4125000:64:100:fp16,"
                                        "33000000:512:100:fp16,33000000:1024:100:fp16")
store = None
for case in cases.split(","):
    rows, b, k, dt = case.split(":")
    rows, b, k = int(rows), int(b), int(k)
    dtype = torch.bfloat16 if dt == "bf16" else torch.float16
    if store is None or store.shape[0] < rows or store.dtype != dtype:
        store = None
        g = torch.Generator(device=dev).manual_seed(1)
        store = torch.empty(rows, 768, dtype=dtype, device=dev)
        for s in range(0, rows, 1 << 20):
            c = torch.randn(min(1 << 20, rows - s), 768, generator=g, device=dev)
            store[s:s + c.shape[0]] = torch.nn.functional.normalize(c, dim=1).to(dtype)
    m = eng.MipsEngine(768, dtype, dev)
    m.bind(store[:rows])
    q = torch.nn.functional.normalize(torch.randn(b, 768, device=dev), dim=1)
    m.search(q, k)
    torch.cuda.synchronize()
    print(f"case rows={rows} batch={b} k={k} dtype={dt} launches={m.last_launch_count()}", flush=True)
    m.close()
