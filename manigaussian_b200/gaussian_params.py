"""View-independent per-Gaussian pre-ops of the rasterizer: activations, deformation offsets, feature normalisation.

Host mirror of the elementwise PyTorch chain ManiGaussian applies to the regressed Gaussian maps before each
render -- agents/manigaussian_bc/models_embed.py:245-252 (current frame), :297-304 (next frame, through the
deformation field's offsets) and agents/manigaussian_bc/gaussian_renderer/__init__.py:66-68 (feature
normalisation) -- as ONE fused sm_100a kernel per direction behind `mgs_activate` / `mgs_activate_backward`
(include/mgs_rasterizer.h).  The activated arrays are what `GaussianRasterizer` / `render_views` consume; the
backward receives the per-Gaussian gradients summed over all views and returns gradients w.r.t. the raw maps and
the offsets (d_means / d_rot / d_scales, the "gradients through the deformation-field offsets" of the dyna path).

PyTorch carries device memory and autograd plumbing only.
"""
import torch

from . import _binding as _b
from .rasterizer import _prep, _ptr, _stream

SCALE_IDENTITY, SCALE_EXP_CLAMP = 0, 1
OPACITY_IDENTITY, OPACITY_SIGMOID = 0, 1


def _modes(scale_activation, opacity_activation):
    sm = {None: SCALE_IDENTITY, "identity": SCALE_IDENTITY, "exp_clamp": SCALE_EXP_CLAMP}
    om = {None: OPACITY_IDENTITY, "identity": OPACITY_IDENTITY, "sigmoid": OPACITY_SIGMOID}
    if scale_activation not in sm:
        raise ValueError(f"unknown scale_activation {scale_activation!r}")
    if opacity_activation not in om:
        raise ValueError(f"unknown opacity_activation {opacity_activation!r}")
    return sm[scale_activation], om[opacity_activation]


def _check(means, rot, scales, opac, feature, d_means, d_rot, d_scales):
    P = means.shape[0]
    if means.dim() != 2 or means.shape[1] != 3:
        raise ValueError("means must have dimensions (num_points, 3)")
    for name, t, shape in (("rotations", rot, (P, 4)), ("scales", scales, (P, 3)), ("d_means", d_means, (P, 3)),
                           ("d_rot", d_rot, (P, 4)), ("d_scales", d_scales, (P, 3))):
        if t is not None and tuple(t.shape) != shape:
            raise ValueError(f"{name} must have dimensions {shape}")
    if opac.numel() != P:
        raise ValueError("opacity must have num_points elements")
    if feature is not None and (feature.dim() != 2 or feature.shape[0] != P):
        raise ValueError("features must have dimensions (num_points, F)")
    return P


def activate_gaussians_raw(means, rot, scales, opac, feature, d_means, d_rot, d_scales, scale_mode, scale_max,
                           opacity_mode, rot_normalize, feature_normalize):
    """One launch of mgs_activate; returns (means, rot, scales, opac, feature) activated (feature None when absent)."""
    dev = means.device
    if dev.type != "cuda":
        raise RuntimeError("manigaussian_b200 runs on CUDA tensors only (no CPU path)")
    means, rot, scales, opac = (_prep(t, dev) for t in (means, rot, scales, opac))
    feature, d_means, d_rot, d_scales = (_prep(t, dev) for t in (feature, d_means, d_rot, d_scales))
    P = _check(means, rot, scales, opac, feature, d_means, d_rot, d_scales)
    F = 0 if feature is None else int(feature.shape[1])
    o_means, o_rot, o_scales, o_opac = (torch.empty_like(t) for t in (means, rot, scales, opac))
    o_feat = torch.empty_like(feature) if F > 0 else None
    with torch.cuda.device(dev):  # the tensors' device need not be the current one (train_device=rank without set_device)
        _b.check(_b.lib().mgs_activate(P, F, _ptr(means), _ptr(d_means), _ptr(rot), _ptr(d_rot), _ptr(scales), _ptr(d_scales),
                                       _ptr(opac), _ptr(feature), scale_mode, float(scale_max), opacity_mode,
                                       int(bool(rot_normalize)), int(bool(feature_normalize)),
                                       _ptr(o_means), _ptr(o_rot), _ptr(o_scales), _ptr(o_opac), _ptr(o_feat), _stream(dev)),
                  "mgs_activate")
    return o_means, o_rot, o_scales, o_opac, o_feat


def activate_gaussians_backward_raw(means, rot, scales, opac, feature, d_means, d_rot, d_scales, scale_mode, scale_max,
                                    opacity_mode, rot_normalize, feature_normalize,
                                    g_means, g_rot, g_scales, g_opac, g_feature, want):
    """One launch of mgs_activate_backward.  g_* are gradients w.r.t. the activated arrays (None skips a field);
    `want` is an 8-tuple of bools (means, d_means, rot, d_rot, scales, d_scales, opac, feature) selecting outputs."""
    dev = means.device
    P = means.shape[0]
    F = 0 if feature is None else int(feature.shape[1])
    g_means, g_rot, g_scales, g_opac, g_feature = (_prep(t, dev) for t in (g_means, g_rot, g_scales, g_opac, g_feature))
    like = (means, means, rot, rot, scales, scales, opac, feature)
    have = (g_means, g_means, g_rot, g_rot, g_scales, g_scales, g_opac, g_feature)
    outs = [torch.empty_like(l) if (w and h is not None and l is not None) else None for w, l, h in zip(want, like, have)]
    with torch.cuda.device(dev):
        _b.check(_b.lib().mgs_activate_backward(P, F, _ptr(means), _ptr(d_means), _ptr(rot), _ptr(d_rot), _ptr(scales),
                                                _ptr(d_scales), _ptr(opac), _ptr(feature), scale_mode, float(scale_max),
                                                opacity_mode, int(bool(rot_normalize)), int(bool(feature_normalize)),
                                                _ptr(g_means), _ptr(g_rot), _ptr(g_scales), _ptr(g_opac), _ptr(g_feature),
                                                *[_ptr(o) for o in outs], _stream(dev)),
                  "mgs_activate_backward")
    return outs


def normalize_features_raw(feature):
    """feature / (||feature|| + 1e-12) per row (gaussian_renderer/__init__.py:66-68): mgs_activate with only the feature field."""
    if feature.device.type != "cuda":
        raise RuntimeError("manigaussian_b200 runs on CUDA tensors only (no CPU path)")
    P, F = feature.shape
    out = torch.empty_like(feature)
    if P and F:
        with torch.cuda.device(feature.device):
            _b.check(_b.lib().mgs_activate(P, F, None, None, None, None, None, None, None, _ptr(feature), 0, 0.0, 0, 0, 1,
                                           None, None, None, None, _ptr(out), _stream(feature.device)), "mgs_activate")
    return out


def normalize_features_backward_raw(feature, g):
    P, F = feature.shape
    g = _prep(g, feature.device)
    out = torch.empty_like(feature)
    if P and F:
        with torch.cuda.device(feature.device):
            _b.check(_b.lib().mgs_activate_backward(P, F, None, None, None, None, None, None, None, _ptr(feature), 0, 0.0, 0, 0, 1,
                                                    None, None, None, None, _ptr(g), None, None, None, None, None, None, None,
                                                    _ptr(out), _stream(feature.device)), "mgs_activate_backward")
    return out


class _ActivateGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, rot, scales, opac, feature, d_means, d_rot, d_scales, cfg):
        scale_mode, scale_max, opacity_mode, rot_normalize, feature_normalize = cfg
        outs = activate_gaussians_raw(means, rot, scales, opac, feature, d_means, d_rot, d_scales, scale_mode, scale_max,
                                      opacity_mode, rot_normalize, feature_normalize)
        ctx.cfg = cfg
        ctx.has = tuple(t is not None for t in (feature, d_means, d_rot, d_scales))
        ctx.opac_shape = opac.shape
        dev = means.device
        ctx.save_for_backward(*[(_prep(t, dev) if t is not None else torch.empty(0, device=dev))
                                for t in (means, rot, scales, opac, feature, d_means, d_rot, d_scales)])
        o_means, o_rot, o_scales, o_opac, o_feat = outs
        o_opac = o_opac.view(opac.shape)
        if o_feat is None:
            o_feat = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(o_feat)
        return o_means, o_rot, o_scales, o_opac, o_feat

    @staticmethod
    def backward(ctx, g_means, g_rot, g_scales, g_opac, g_feat):
        saved = list(ctx.saved_tensors)
        means, rot, scales, opac = saved[:4]
        feature, d_means, d_rot, d_scales = (t if h else None for t, h in zip(saved[4:], ctx.has))
        need = ctx.needs_input_grad  # means rot scales opac feature d_means d_rot d_scales cfg
        want = (need[0], need[5] and d_means is not None, need[1], need[6] and d_rot is not None,
                need[2], need[7] and d_scales is not None, need[3], need[4] and feature is not None)
        if feature is None:
            g_feat = None
        outs = activate_gaussians_backward_raw(means, rot, scales, opac, feature, d_means, d_rot, d_scales, *ctx.cfg,
                                               g_means, g_rot, g_scales, g_opac, g_feat, want)
        dm, ddm, dr, ddr, ds, dds, do, df = outs
        if do is not None:
            do = do.view(ctx.opac_shape)
        return dm, dr, ds, do, df, ddm, ddr, dds, None


def activate_gaussians(means, rotations, scales, opacity, features=None, d_means=None, d_rotations=None, d_scales=None,
                       scale_activation="exp_clamp", scale_max=0.05, opacity_activation="sigmoid",
                       normalize_rotation=True, normalize_feature=True):
    """Fused, differentiable pre-ops of one Gaussian cloud.

    Current frame (models_embed.py:245-252): `activate_gaussians(xyz, rot_maps, scale_maps, opacity_maps, feature_maps,
    d_means=xyz_maps)`.  Next frame (:297-304): `activate_gaussians(xyz_act.detach(), rot_act.detach(), scale_act.detach(),
    opacity_act.detach(), feature_maps.detach(), d_means=next_xyz_maps, d_rotations=next_rot_maps, scale_activation=None,
    opacity_activation=None)`.  Returns (means, rotations, scales, opacity, features) ready for the rasterizer;
    `features` is None when no features were given."""
    sm, om = _modes(scale_activation, opacity_activation)
    cfg = (sm, float(scale_max), om, bool(normalize_rotation), bool(normalize_feature))
    o = _ActivateGaussians.apply(means, rotations, scales, opacity, features, d_means, d_rotations, d_scales, cfg)
    return o[0], o[1], o[2], o[3], (o[4] if features is not None else None)
