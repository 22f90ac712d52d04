"""Allocator behaviour of the eager end-to-end step (diagnostic): python tools/diag_alloc.py [workload] [heads]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

wlname = sys.argv[1] if len(sys.argv) > 1 else "mg"
heads = len(sys.argv) > 2 and sys.argv[2] == "heads"
wl = dict(bench.WORKLOADS[wlname])
g, cams, cts = bench.host_inputs(wl, 0, 1)
Gh, Ch, Th = bench.to_device(g, cams, cts, torch, pinned=True)
step = bench.make_e2e("ours", wl, torch, None, heads=heads)
st = {}
prev = torch.cuda.memory_stats().get("num_device_alloc", 0)
for blk in range(10):
    t0 = time.perf_counter()
    for _ in range(20):
        step(Gh, Ch, Th, st)
    step.flush()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20 * 1e3
    ms = torch.cuda.memory_stats()
    print(f"block {blk}: {dt:.3f} ms/step  allocated {torch.cuda.memory_allocated() / 1e6:.1f} MB  reserved {torch.cuda.memory_reserved() / 1e6:.1f} MB  "
          f"cudaMalloc +{ms.get('num_device_alloc', 0) - prev}  active blocks {ms.get('active.all.current')}  inactive split {ms.get('inactive_split.all.current')}")
    prev = ms.get("num_device_alloc", 0)
# largest live blocks
snap = torch.cuda.memory_snapshot()
sizes = {}
for seg in snap:
    for b in seg["blocks"]:
        if b["state"] != "inactive":
            sizes[b["size"]] = sizes.get(b["size"], 0) + 1
print("live blocks by size:", sorted(sizes.items(), key=lambda kv: -kv[0] * kv[1])[:12])
