"""Generates tests/golden/loss_heads.npz by running the REFERENCE's own loss functions in the build container:
l2_loss and cosine_loss are lifted out of agents/manigaussian_bc/loss.py with `ast` (the module imports einops & co. at
the top; only the two function definitions are executed) and differentiated with torch autograd on the CPU, applied as
NeuralRenderer.forward does (neural_rendering.py:300-318: images channel-last, one view).  Nothing of the reference is
copied into this repository.   Run:  python tests/golden/make_loss_golden.py"""
import ast
import os

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference/agents/manigaussian_bc/loss.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_losses():
    tree = ast.parse(open(REF).read())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("l2_loss", "cosine_loss")]
    ns = {"torch": torch, "F": F}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "loss.py", "exec"), ns)
    return ns["l2_loss"], ns["cosine_loss"]


def main():
    l2_loss, cosine_loss = load_reference_losses()
    rng = np.random.default_rng(11)
    out = {}
    for name, (Fch, H, W) in {"f32": (32, 24, 40), "f3": (3, 17, 9)}.items():
        render = torch.tensor(rng.uniform(-0.2, 1.2, (3, H, W)).astype(np.float32), requires_grad=True)
        gt = torch.tensor(rng.uniform(0, 1, (3, H, W)).astype(np.float32))
        embed = rng.normal(0, 1, (Fch, H, W)).astype(np.float32)
        embed[:, 0, 0] = 0.0          # a pixel with a zero rendered embedding (norm clamp branch)
        embed[:, 1, 1] *= 1e-6
        embed = torch.tensor(embed, requires_grad=True)
        gt_e = rng.normal(0, 1, (Fch, H, W)).astype(np.float32)
        gt_e[:, 2, 2] = 0.0           # a pixel with a zero target embedding
        gt_e = torch.tensor(gt_e)
        # the reference compares channel-last batches [B,H,W,C] (neural_rendering.py:283, :311-312)
        l_rgb = l2_loss(render.permute(1, 2, 0)[None], gt.permute(1, 2, 0)[None])
        l_emb = cosine_loss(embed.permute(1, 2, 0)[None], gt_e.permute(1, 2, 0)[None])
        (g_r,) = torch.autograd.grad(l_rgb, render)
        (g_e,) = torch.autograd.grad(l_emb, embed)
        for k, v in dict(render=render, gt=gt, embed=embed, gt_embed=gt_e, loss_rgb=l_rgb, loss_embed=l_emb, d_render=g_r, d_embed=g_e).items():
            out[f"{name}_{k}"] = v.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "loss_heads.npz"), **out)
    print("wrote loss_heads.npz", {k: v.shape for k, v in out.items() if "loss" in k})


if __name__ == "__main__":
    main()
