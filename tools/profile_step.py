"""Host/GPU time breakdown of one bench step (diagnostic; run on the GPU box)."""
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

impl_name = sys.argv[1] if len(sys.argv) > 1 else "ours"
wlname = sys.argv[2] if len(sys.argv) > 2 else "c3"
wl = dict(bench.WORKLOADS[wlname])
g, cams, cts = bench.host_inputs(wl, 0, 1)
G, C, T = bench.to_device(g, cams, cts, torch)
impl = bench.Impl(wl["F"], wl["depth"]) if impl_name == "ours" else bench.RefImpl(wl["F"], wl["depth"])
flat, acc = bench.make_packed(wl["P"], wl["F"], 4, torch)
for _ in range(3):
    bench.run_step(impl, G, C, T, flat, acc)
torch.cuda.synchronize()


def seg():
    t = {"fwd": 0.0, "bwd": 0.0, "acc": 0.0}
    flat.zero_()
    for cam, ct in zip(C, T):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = impl.fwd(G, cam)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        grads = impl.bwd(G, cam, out, ct)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        gd = dict(zip(bench.GRAD_ORDER, grads))
        for k, v in acc.items():
            v.add_(gd[k].reshape(v.shape))
        torch.cuda.synchronize(); t3 = time.perf_counter()
        t["fwd"] += t1 - t0; t["bwd"] += t2 - t1; t["acc"] += t3 - t2
    return t


for _ in range(2):
    print(impl_name, wlname, {k: round(v * 1e3, 3) for k, v in seg().items()}, "ms per step (synchronised segments)")
ms0 = torch.cuda.memory_stats()
per = []
for _ in range(30):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    bench.run_step(impl, G, C, T, flat, acc)
    torch.cuda.synchronize(); per.append(round((time.perf_counter() - t0) * 1e3, 2))
ms1 = torch.cuda.memory_stats()
print("30 consecutive steps (ms):", per)
print("cudaMalloc calls during them:", ms1.get("num_device_alloc", 0) - ms0.get("num_device_alloc", 0),
      "cudaFree:", ms1.get("num_device_free", 0) - ms0.get("num_device_free", 0),
      "reserved MB:", ms1["reserved_bytes.all.current"] / 1e6, "allocated MB:", ms1["allocated_bytes.all.current"] / 1e6)
# host-only cost: how long does the Python/driver side take when the GPU is not waited for
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    bench.run_step(impl, G, C, T, flat, acc)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("5 steps: host returned after %.2f ms, GPU drained after %.2f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    bench.run_step(impl, G, C, T, flat, acc)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
print(s.getvalue()[:6000])
