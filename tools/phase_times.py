"""Phase timing of one multi-view step (diagnostic): forward of all views / backward of all views / accumulation."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from manigaussian_b200 import rasterizer as R
from manigaussian_b200 import GaussianRasterizationSettings as S

wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"])
g, cams, cts = bench.host_inputs(wl, 0, 1)
G, C, T = bench.to_device(g, cams, cts, torch)
F, depth = wl["F"], wl["depth"]
views = [S(c["H"], c["W"], c["tanfovx"], c["tanfovy"], c["bg"], 1.0, c["viewmatrix"], c["projmatrix"], 1, c["campos"], False, False, F > 0) for c in C]
flat, acc = bench.make_packed(wl["P"], F, 4, torch)
sync = torch.cuda.synchronize


def step(timed):
    t = [time.perf_counter()]
    outs, sts = R.rasterize_views_raw(views, G["means3D"], G["empty"], G["feature"], G["opacities"], G["scales"], G["rotations"], 1.0,
                                      G["empty"], G["shs"], 1, F > 0, return_depth=depth)
    t.append(time.perf_counter())
    if timed: sync()
    t.append(time.perf_counter())
    grads = R.rasterize_views_backward_raw(views, outs, sts, [x["dL_dcolor"] for x in T], [x["dL_dfeature"] for x in T] if F else None,
                                           G["means3D"], G["empty"], G["feature"], G["scales"], G["rotations"], 1.0, G["empty"], G["shs"], 1, F > 0)
    t.append(time.perf_counter())
    if timed: sync()
    t.append(time.perf_counter())
    flat.zero_()
    for o, g9 in zip(outs, grads):
        gd = dict(zip(bench.GRAD_ORDER, g9))
        for k, v in acc.items():
            v.add_(gd[k].reshape(v.shape))
    if timed: sync()
    t.append(time.perf_counter())
    return [round((b - a) * 1e3, 3) for a, b in zip(t, t[1:])]


for _ in range(10):
    step(False)
sync()
for _ in range(3):
    print("host fwd issue, fwd drain, host bwd issue, bwd drain, accumulate (ms):", step(True))
sync(); t0 = time.perf_counter()
for _ in range(20):
    step(False)
sync(); print("untimed steady state ms/step: %.3f" % ((time.perf_counter() - t0) / 20 * 1e3))
