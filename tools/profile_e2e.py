"""Host-side profile of bench.py's end-to-end step (diagnostic; run on the GPU box):  python tools/profile_e2e.py [workload]"""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

wlname = sys.argv[1] if len(sys.argv) > 1 else "mg"
wl = dict(bench.WORKLOADS[wlname])
g, cams, cts = bench.host_inputs(wl, 0, 1)
Gh, Ch, Th = bench.to_device(g, cams, cts, torch, pinned=True)
step = bench.make_e2e("ours", wl, torch, None)
st = {}
for _ in range(10):
    step(Gh, Ch, Th, st)
step.flush()
torch.cuda.synchronize()
per = []
for _ in range(30):
    t0 = time.perf_counter()
    step(Gh, Ch, Th, st)
    per.append(round((time.perf_counter() - t0) * 1e3, 3))
step.flush()
torch.cuda.synchronize()
print("host ms per step:", per)
ms0 = torch.cuda.memory_stats()
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step(Gh, Ch, Th, st)
step.flush()
torch.cuda.synchronize()
pr.disable()
ms1 = torch.cuda.memory_stats()
print("cudaMalloc during 50 steps:", ms1.get("num_device_alloc", 0) - ms0.get("num_device_alloc", 0), "cudaFree:", ms1.get("num_device_free", 0) - ms0.get("num_device_free", 0))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
print(s.getvalue()[:6000])
