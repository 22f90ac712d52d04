"""N > 1 on real GPUs (skipped on single-GPU boxes; the host logic is covered on CPU by test_multigpu_cpu.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_view_parallel_gradients_match_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tools", "check_view_parallel.py")], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "PASS" in r.stdout


def test_ops_run_on_the_tensors_device_not_the_current_one():
    """Tensors on cuda:1 while cuda:0 is the current device: every op must launch on (and allocate from) the tensors' device.
    The fused pre-ops and the rasterizer are compared with the same call made with cuda:1 current."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    from manigaussian_b200 import GaussianRasterizationSettings, GaussianRasterizer
    from manigaussian_b200.gaussian_params import activate_gaussians
    inp = util.make_inputs(P=3000, W=48, H=32, F=32, seed=5)
    g, cam = inp["g"], inp["cam"]
    dev = torch.device("cuda", 1)

    def run():
        t = {k: torch.tensor(np.asarray(g[k], np.float32), device=dev).requires_grad_(True)
             for k in ("means3D", "rotations", "scales", "opacities", "shs", "feature")}
        means, rot, scales, opac, nf = activate_gaussians(t["means3D"], t["rotations"], t["scales"].log(),
                                                          torch.logit(t["opacities"].clamp(1e-3, 1 - 1e-3)), features=t["feature"])
        s = GaussianRasterizationSettings(cam["H"], cam["W"], cam["tanfovx"], cam["tanfovy"],
                                          torch.tensor(np.asarray(inp["bg"], np.float32), device=dev), 1.0,
                                          torch.tensor(np.asarray(cam["viewmatrix"], np.float32), device=dev),
                                          torch.tensor(np.asarray(cam["projmatrix"], np.float32), device=dev), 1,
                                          torch.tensor(np.asarray(cam["campos"], np.float32), device=dev), False, False, True)
        img, emb, radii = GaussianRasterizer(s)(means3D=means, means2D=torch.zeros_like(means), opacities=opac, shs=t["shs"],
                                                language_feature_precomp=nf, scales=scales, rotations=rot)
        assert img.device == dev and emb.device == dev and radii.device == dev
        (img.sum() + emb.square().sum()).backward()
        torch.cuda.synchronize(dev)
        return [img.detach().cpu().numpy(), emb.detach().cpu().numpy()] + [t[k].grad.cpu().numpy() for k in sorted(t)]

    with torch.cuda.device(0):
        a = run()
    with torch.cuda.device(1):
        b = run()
    for x, y in zip(a, b):
        assert np.isfinite(x).all()
        assert util.rel_l2(x, y) < 1e-6
