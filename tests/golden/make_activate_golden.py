"""Generates tests/golden/activate_*.npz with the REAL PyTorch operators the reference applies
(agents/manigaussian_bc/models_embed.py:85-88,245-252,297-304; gaussian_renderer/__init__.py:66-68), on CPU with
autograd.  Run in the build container:  python tests/golden/make_activate_golden.py
The fixtures pin oracle/activate_oracle.py (tests/test_activate_cpu.py) and, through it, the CUDA kernels."""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def make(name, P, F, seed, deformed, edge=False):
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    xyz, xyz_maps = rn(P, 3), 0.05 * rn(P, 3)
    rot_maps, scale_maps, opacity_maps = rn(P, 4), -3.5 + 0.8 * rn(P, 3), -2 + rn(P, 1)
    feature_maps = rn(P, F) if F else None
    if edge:  # zero-norm rows and a scale exactly at the clamp
        rot_maps[0] = 0
        if F:
            feature_maps[1] = 0
        scale_maps[2] = float(np.log(0.05))
        scale_maps[3] = 5.0
    leaves = [xyz_maps, rot_maps, scale_maps, opacity_maps] + ([feature_maps] if F else [])
    for t in leaves:
        t.requires_grad_(True)
    # --- the reference's expressions, verbatim in meaning
    scales = torch.clamp_max(torch.exp(scale_maps), 0.05)
    means = xyz + xyz_maps
    rots = torch.nn.functional.normalize(rot_maps, dim=-1)
    opac = torch.sigmoid(opacity_maps)
    feat = feature_maps / (feature_maps.norm(dim=-1, keepdim=True) + 1e-12) if F else None
    rec = dict(P=P, F=F, xyz=xyz, xyz_maps=xyz_maps, rot_maps=rot_maps, scale_maps=scale_maps, opacity_maps=opacity_maps)
    outs = dict(means=means, rot=rots, scales=scales, opac=opac)
    if F:
        rec["feature_maps"] = feature_maps
        outs["feature"] = feat
    cot = {k: rn(*v.shape) for k, v in outs.items()}
    loss = sum((cot[k] * v).sum() for k, v in outs.items())
    if deformed:
        next_xyz, next_rot = 0.01 * rn(P, 3), 0.05 * rn(P, 4)
        next_xyz.requires_grad_(True)
        next_rot.requires_grad_(True)
        n_means = means.detach() + next_xyz
        n_rot = torch.nn.functional.normalize(rots.detach() + next_rot, dim=-1)
        n_feat = feature_maps.detach() / (feature_maps.detach().norm(dim=-1, keepdim=True) + 1e-12) if F else None
        ncot = dict(means=rn(P, 3), rot=rn(P, 4))
        loss = loss + (ncot["means"] * n_means).sum() + (ncot["rot"] * n_rot).sum()
        rec.update(next_xyz=next_xyz, next_rot=next_rot, next_out_means=n_means, next_out_rot=n_rot,
                   next_cot_means=ncot["means"], next_cot_rot=ncot["rot"])
        if F:
            rec["next_out_feature"] = n_feat
    loss.backward()
    for k, v in outs.items():
        rec["out_" + k] = v
        rec["cot_" + k] = cot[k]
    rec.update(grad_xyz_maps=xyz_maps.grad, grad_rot_maps=rot_maps.grad, grad_scale_maps=scale_maps.grad,
               grad_opacity_maps=opacity_maps.grad)
    if F:
        rec["grad_feature_maps"] = feature_maps.grad
    if deformed:
        rec.update(grad_next_xyz=next_xyz.grad, grad_next_rot=next_rot.grad)
    np.savez_compressed(os.path.join(HERE, f"activate_{name}.npz"),
                        **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in rec.items()})


if __name__ == "__main__":
    torch.set_num_threads(1)
    make("f32", P=257, F=32, seed=11, deformed=True)
    make("f3_edge", P=64, F=3, seed=12, deformed=False, edge=True)
    make("f0", P=100, F=0, seed=13, deformed=True)
    print("wrote", [f for f in os.listdir(HERE) if f.startswith("activate_")])
