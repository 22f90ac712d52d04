"""Work distribution of the blend CTAs on a bench workload: tile list lengths, contributing depth per 8x4 block
(max n_contrib = how far the backward walks), and what the longest CTA means for the kernel's critical path.
Run on a GPU box: python tools/tile_stats.py [c3|c5]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from manigaussian_b200 import rasterizer as R
from manigaussian_b200 import GaussianRasterizationSettings as S

wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"])
g, cams, cts = bench.host_inputs(wl, 0, 1)
G, C, T = bench.to_device(g, cams, cts, torch)
F = wl["F"]
views = [S(c["H"], c["W"], c["tanfovx"], c["tanfovy"], c["bg"], 1.0, c["viewmatrix"], c["projmatrix"], bench.SH_DEGREE, c["campos"],
           False, False, F > 0) for c in C]
outs, sts = R.rasterize_views_raw(views, G["means3D"], G["empty"], G["feature"], G["opacities"], G["scales"], G["rotations"], 1.0,
                                  G["empty"], G["shs"], bench.SH_DEGREE, F > 0)
torch.cuda.synchronize()
res = []
for v, (o, c) in enumerate(zip(outs, C)):
    H, W = c["H"], c["W"]
    N = H * W
    gx, gy = (W + 15) // 16, (H + 15) // 16
    Tn = gx * gy
    al = lambda x: (x + 127) // 128 * 128
    buf = o[6]  # imgBuffer: final_T [N] f32, n_contrib [N] u32, ranges [T] uint2, each 128-byte aligned
    base = buf.data_ptr()
    off0 = (-base) % 128
    raw = buf.cpu().numpy()
    o1 = off0 + al(4 * N)
    o2 = o1 + al(4 * N)
    ncontrib = raw[o1:o1 + 4 * N].view(np.uint32).reshape(H, W)
    ranges = raw[o2:o2 + 8 * Tn].view(np.uint32).reshape(Tn, 2)
    lens = (ranges[:, 1] - ranges[:, 0]).astype(np.int64)
    # per 8x4 block: max n_contrib (1-based position in the tile list of the last contributor)
    blk = ncontrib.reshape(gy, 4, 4, gx, 2, 8).transpose(0, 3, 1, 4, 2, 5).reshape(Tn, 8, 32).max(axis=2)
    depth_frac = blk / np.maximum(lens[:, None], 1)
    res.append(dict(view=v, tiles=int(Tn), R=int(lens.sum()), len_mean=float(lens.mean()), len_max=int(lens.max()),
                    len_p50=float(np.percentile(lens, 50)), len_p90=float(np.percentile(lens, 90)), len_p99=float(np.percentile(lens, 99)),
                    fwd_walk_mean=float(blk.mean()), fwd_walk_max=int(blk.max()),
                    max_over_mean_len=float(lens.max() / max(lens.mean(), 1)), max_over_mean_walk=float(blk.max() / max(blk.mean(), 1)),
                    walked_fraction_of_list=float(blk.sum() / max(8 * lens.sum(), 1)),
                    top_tiles=sorted(lens.tolist(), reverse=True)[:12]))
print(json.dumps(res, indent=1))
