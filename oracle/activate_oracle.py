"""CPU oracle (numpy, fp32) of ManiGaussian's per-Gaussian pre-ops -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.  It restates, operator by operator,
the PyTorch expressions of
  agents/manigaussian_bc/models_embed.py:245-252   exp + clamp_max(0.05), xyz + xyz_maps, F.normalize, sigmoid
  agents/manigaussian_bc/models_embed.py:297-304   xyz.detach() + next_xyz, F.normalize(rot.detach() + next_rot)
  agents/manigaussian_bc/gaussian_renderer/__init__.py:66-68   feature / (feature.norm(dim=-1, keepdim=True) + 1e-12)
and their autograd derivatives (ATen: exp_backward g*y; clamp_max mask x <= max; sigmoid_backward g*(1-y)*y;
norm_backward g*x/n with 0 at n == 0; div_backward).  Pinned: tests/golden/activate_*.npz hold outputs of the real
torch operators (CPU, autograd) produced by tests/golden/make_activate_golden.py; tests/test_activate_cpu.py
checks this file against them (<= 2e-6 rel-L2; row reductions run in a different order than ATen's).
"""
import numpy as np

f32 = np.float32
ROT_EPS = f32(1e-12)
FEAT_EPS = f32(1e-12)


def _pre(x, d):
    x = np.asarray(x, f32)
    return x if d is None else (x + np.asarray(d, f32)).astype(f32)


def forward(means, rot, scales, opac, feature=None, d_means=None, d_rot=None, d_scales=None,
            scale_mode=1, scale_max=0.05, opacity_mode=1, rot_normalize=True, feature_normalize=True):
    """Returns dict(means, rot, scales, opac, feature) of fp32 arrays."""
    out = {}
    out["means"] = _pre(means, d_means)
    s = _pre(scales, d_scales)
    out["scales"] = np.minimum(np.exp(s), f32(scale_max)).astype(f32) if scale_mode == 1 else s
    q = _pre(rot, d_rot)
    if rot_normalize:
        n = np.sqrt((q * q).sum(-1, keepdims=True, dtype=f32)).astype(f32)
        q = (q / np.maximum(n, ROT_EPS)).astype(f32)
    out["rot"] = q
    o = np.asarray(opac, f32)
    out["opac"] = (f32(1) / (f32(1) + np.exp(-o))).astype(f32) if opacity_mode == 1 else o
    if feature is not None:
        x = np.asarray(feature, f32)
        if feature_normalize:
            n = np.sqrt((x * x).sum(-1, keepdims=True, dtype=f32)).astype(f32)
            x = (x / (n + FEAT_EPS)).astype(f32)
        out["feature"] = x
    else:
        out["feature"] = None
    return out


def backward(means, rot, scales, opac, feature, d_means, d_rot, d_scales, g,
             scale_mode=1, scale_max=0.05, opacity_mode=1, rot_normalize=True, feature_normalize=True):
    """g: dict of gradients w.r.t. the activated arrays (keys as forward's).  Returns dict of gradients w.r.t. the raw
    arrays; the gradient w.r.t. an offset equals the gradient w.r.t. the array it is added to."""
    out = {}
    out["means"] = np.asarray(g["means"], f32)
    gs = np.asarray(g["scales"], f32)
    if scale_mode == 1:
        e = np.exp(_pre(scales, d_scales)).astype(f32)
        gs = np.where(e <= f32(scale_max), gs * e, f32(0)).astype(f32)
    out["scales"] = gs
    gq = np.asarray(g["rot"], f32)
    if rot_normalize:
        q = _pre(rot, d_rot)
        n = np.sqrt((q * q).sum(-1, keepdims=True, dtype=f32)).astype(f32)
        d = np.maximum(n, ROT_EPS)
        dot = (gq * q).sum(-1, keepdims=True, dtype=f32)
        with np.errstate(divide="ignore", invalid="ignore"):
            k = np.where((n >= ROT_EPS) & (n > 0), dot / (d * d * n), f32(0)).astype(f32)
        gq = (gq / d - q * k).astype(f32)
    out["rot"] = gq
    go = np.asarray(g["opac"], f32)
    if opacity_mode == 1:
        o = np.asarray(opac, f32)
        y = (f32(1) / (f32(1) + np.exp(-o))).astype(f32)
        go = ((go.reshape(y.shape) * (f32(1) - y)) * y).astype(f32)
    out["opac"] = go
    if feature is not None and g.get("feature") is not None:
        x = np.asarray(feature, f32)
        gf = np.asarray(g["feature"], f32)
        if feature_normalize:
            n = np.sqrt((x * x).sum(-1, keepdims=True, dtype=f32)).astype(f32)
            d = n + FEAT_EPS
            dot = (gf * x).sum(-1, keepdims=True, dtype=f32)
            with np.errstate(divide="ignore", invalid="ignore"):
                k = np.where(n > 0, dot / (d * d * n), f32(0)).astype(f32)
            gf = (gf / d - x * k).astype(f32)
        out["feature"] = gf
    else:
        out["feature"] = None
    return out
