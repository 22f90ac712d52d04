"""View-parallel correctness: gradients of `render_views(..., sync_gradients=True)` with the views sharded
over the ranks == gradients of all views rendered on one GPU (rel-L2 <= 1e-5; fp32 summation order differs).
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_view_parallel.py"""
import datetime
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    from manigaussian_b200 import GaussianRasterizationSettings as S
    from manigaussian_b200 import scenes
    from manigaussian_b200.gaussian_renderer import render_views
    from manigaussian_b200.parallel import shard_views
    P, F, W, H = 30000, 32, 96, 96
    V = 2 * world
    g = scenes.make_gaussians(P, F=F, sh_degree=1, seed=7)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    cams = [scenes.make_camera(W, H, v, V) for v in range(V)]
    views = [S(H, W, c["tanfovx"], c["tanfovy"], torch.zeros(3, device="cuda"), 1.0, t(c["viewmatrix"]), t(c["projmatrix"]), 1,
               t(c["campos"]), False, False, True) for c in cams]
    cts = [scenes.make_cotangents(W, H, F, seed=50 + v) for v in range(V)]

    def run(ids, sync):
        L = {k: t(g[k]).requires_grad_(True) for k in ("means3D", "rotations", "scales", "opacities", "shs", "feature")}
        o = render_views(views, L["means3D"], L["rotations"], L["scales"], L["opacities"], features_color=L["shs"],
                         features_language=L["feature"], view_ids=ids, sync_gradients=sync)
        loss = sum((o["render"][i] * t(cts[v]["dL_dcolor"])).sum() + (o["render_embed"][i] * t(cts[v]["dL_dfeature"])).sum()
                   for i, v in enumerate(ids))
        loss.backward()
        return {k: v.grad.clone() for k, v in L.items()}

    mine = run(shard_views(V, rank, world), True)
    full = run(list(range(V)), None)
    worst = 0.0
    for k in mine:
        a, b = mine[k].double(), full[k].double()
        worst = max(worst, float((a - b).norm() / b.norm()))
    res = torch.tensor([worst], device="cuda")
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    ok = float(res.item()) < 1e-5
    if rank == 0:
        print(f"view-parallel x{world}: max rel-L2 of all-reduced vs single-GPU gradients = {float(res.item()):.3e} ->", "PASS" if ok else "FAIL")
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
