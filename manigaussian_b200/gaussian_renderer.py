"""`render()` and its multi-view form `render_views()` -- the immediate caller of the rasterizer.

`render` keeps the signature, semantics and return dictionary of ManiGaussian's
agents/manigaussian_bc/gaussian_renderer/__init__.py:17-94, so NeuralRenderer.pts2render
(agents/manigaussian_bc/neural_rendering.py:383-402) runs on it unchanged.  Differences are only in what it avoids:
the background colour is uploaded once per (device, colour) instead of every call (:25), tan(fov/2) and the image size
come from host scalars when the camera dictionary was built by `manigaussian_b200.cameras` (no `.item()` syncs, :35-40),
the feature normalisation (:66-68) is the fused sm_100a kernel behind `mgs_activate`, and no dummy [P,3] zero feature
tensor is made when features are absent (:70-71).

`render_views` renders V cameras of ONE Gaussian cloud in a single autograd node: the projection/binning chains of all
views are enqueued on per-view streams by one C call that never synchronises with the host, every view's backward SUMS its
per-Gaussian gradients in registers into one packed buffer (`parallel.PackedGradients`), and -- for the view-parallel
multi-GPU mode -- that buffer is all-reduced with ONE collective before the gradients are handed back to autograd.
The reference has no counterpart (bs == 1 is asserted, neural_rendering.py:386; one view per call).
"""
import math

import torch

from . import gaussian_params as _gp
from .cameras import CameraBatch
from .parallel import PackedGradients
from .rasterizer import (GaussianRasterizationSettings, GaussianRasterizer, _prep, rasterize_views_backward_raw,
                         rasterize_views_raw)

_BG_CACHE = {}
MIN_DENOMINATOR = 1e-12  # gaussian_renderer/__init__.py:67; compiled into the kernel (activate.cu FEAT_EPS)


def _background(bg_color, device):
    if torch.is_tensor(bg_color):
        return bg_color.to(device=device, dtype=torch.float32)
    key = (str(device), tuple(float(c) for c in bg_color))
    t = _BG_CACHE.get(key)
    if t is None:
        t = torch.tensor(key[1], dtype=torch.float32, device=device)
        _BG_CACHE[key] = t
    return t


class _NormalizeFeatures(torch.autograd.Function):
    """feature / (||feature|| + 1e-12) along the last dimension, fused forward and backward kernels."""

    @staticmethod
    def forward(ctx, feature):
        dev = feature.device
        f = _prep(feature, dev)
        ctx.save_for_backward(f)
        return _gp.normalize_features_raw(f)

    @staticmethod
    def backward(ctx, g):
        (f,) = ctx.saved_tensors
        return _gp.normalize_features_backward_raw(f, g)


def normalize_features(features_language):
    return _NormalizeFeatures.apply(features_language)


def _camera_scalars(nv, idx):
    host = nv.get("_host")
    if host is not None:
        return host["tanfovx"][idx], host["tanfovy"][idx], int(host["height"]), int(host["width"])
    # reference-built dictionary: the values live on the device and reading them synchronises, as in the reference
    return (math.tan(float(nv["FovX"][idx]) * 0.5), math.tan(float(nv["FovY"][idx]) * 0.5),
            int(nv["height"][idx]), int(nv["width"][idx]))


def render(data, idx, pts_xyz, rotations, scales, opacity, bg_color, pts_rgb=None, features_color=None,
           features_language=None, return_depth=False):
    """Render view `idx` of data['novel_view'].  Same arguments and result as the reference's render(); with
    return_depth=True the dictionary also holds "depth" [H,W] (view-space z, alpha-blended)."""
    device = pts_xyz.device
    nv = data["novel_view"]
    bg = _background(bg_color, device)
    # zero tensor whose gradient receives the screen-space mean gradients (reference :28-32)
    screenspace_points = torch.zeros_like(pts_xyz, dtype=torch.float32, requires_grad=True, device=device) + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    tanfovx, tanfovy, H, W = _camera_scalars(nv, idx)
    raster_settings = GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=tanfovx, tanfovy=tanfovy, bg=bg, scale_modifier=1.0,
        viewmatrix=nv["world_view_transform"][idx], projmatrix=nv["full_proj_transform"][idx],
        sh_degree=3 if features_color is None else 1, campos=nv["camera_center"][idx], prefiltered=False, debug=False,
        include_feature=(features_language is not None))
    rasterizer = GaussianRasterizer(raster_settings=raster_settings, return_depth=return_depth)
    shs, colors_precomp = None, None
    if features_color is not None:
        shs = features_color
    else:
        assert pts_rgb is not None
        colors_precomp = pts_rgb
    feat = normalize_features(features_language) if features_language is not None else None
    out = rasterizer(means3D=pts_xyz, means2D=screenspace_points, shs=shs, colors_precomp=colors_precomp,
                     language_feature_precomp=feat, opacities=opacity, scales=scales, rotations=rotations, cov3D_precomp=None)
    ret = {"render": out[0], "render_embed": out[1], "viewspace_points": screenspace_points, "radii": out[2]}
    if return_depth:
        ret["depth"] = out[3]
    return ret


class _RasterizeViews(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, feature, opacities, scales, rotations, views, return_depth, group):
        dev = means3D.device
        V, P = len(views), means3D.shape[0]
        H, W = int(views[0].image_height), int(views[0].image_width)
        if any((int(s.image_height), int(s.image_width)) != (H, W) for s in views):
            raise ValueError("render_views needs one image size for all views")
        include = feature is not None and feature.numel() > 0
        F = int(feature.shape[1]) if include else 0
        means3D, sh, colors_precomp, feature = (_prep(t, dev) for t in (means3D, sh, colors_precomp, feature))
        opacities, scales, rotations = (_prep(t, dev) for t in (opacities, scales, rotations))
        opts = dict(dtype=torch.float32, device=dev)
        color = torch.empty((V, 3, H, W), **opts)
        feat_img = torch.empty((V, F, H, W), **opts) if include else None
        depth = torch.empty((V, H, W), **opts) if return_depth else None
        degree = views[0].sh_degree
        outs, streams = rasterize_views_raw(views, means3D, colors_precomp, feature, opacities, scales, rotations,
                                            views[0].scale_modifier, None, sh, degree, include, return_depth=return_depth,
                                            debug=views[0].debug, out_buffers=(color, feat_img, depth))
        main = torch.cuda.current_stream(dev)
        for st in streams:
            main.wait_stream(st)
        radii = torch.stack([o[3] for o in outs])
        ctx.views, ctx.streams, ctx.group = views, streams, group
        ctx.state = [(o[0], o[3], o[4], o[5], o[6]) for o in outs]   # R, radii, geometry / binning / image state per view
        ctx.include, ctx.F, ctx.return_depth, ctx.degree = include, F, return_depth, degree
        ctx.opac_shape = opacities.shape
        ctx.save_for_backward(means3D, sh if sh is not None else torch.empty(0, device=dev),
                              colors_precomp if colors_precomp is not None else torch.empty(0, device=dev),
                              feature if include else torch.empty(0, device=dev), scales, rotations)
        ctx.mark_non_differentiable(radii)
        empty = torch.empty(0, device=dev)
        return color, (feat_img if include else empty), radii, (depth if return_depth else empty)

    @staticmethod
    def backward(ctx, g_color, g_feat, _g_radii, g_depth):
        means3D, sh, colors_precomp, feature, scales, rotations = ctx.saved_tensors
        dev = means3D.device
        views, V, P = ctx.views, len(ctx.views), means3D.shape[0]
        H, W = int(views[0].image_height), int(views[0].image_width)
        M = sh.shape[1] if sh.numel() else 0
        use_colors = colors_precomp.numel() > 0
        if g_color is None:
            g_color = torch.zeros((V, 3, H, W), dtype=torch.float32, device=dev)
        if ctx.include and g_feat is None:
            g_feat = torch.zeros((V, ctx.F, H, W), dtype=torch.float32, device=dev)
        g_color = _prep(g_color, dev)
        g_feat = _prep(g_feat, dev) if ctx.include else None
        g_depth = _prep(g_depth, dev) if (ctx.return_depth and g_depth is not None and g_depth.numel()) else None
        pk = PackedGradients(P, ctx.F, M, dev, colors=use_colors, zero=False)  # every row is written by the backward
        grp = None if ctx.group in (None, True) else ctx.group
        m2d = torch.empty((V, P, 3), dtype=torch.float32, device=dev)
        outs = [(R, None, None, radii, geom, binb, img) for (R, radii, geom, binb, img) in ctx.state]
        rasterize_views_backward_raw(views, outs, ctx.streams, g_color, g_feat, means3D, colors_precomp if use_colors else None,
                                     feature if ctx.include else None, scales, rotations, views[0].scale_modifier, None,
                                     sh if M else None, ctx.degree, ctx.include, grads_depth=g_depth, debug=views[0].debug,
                                     accumulate_into=pk.views, means2D_per_view=m2d,
                                     after_blend=(lambda: pk.all_reduce_begin(grp)) if ctx.group is not None else None)
        if ctx.group is not None:
            # view-parallel multi-GPU: the one exchange of the path, a SUM all-reduce of the packed buffer; the feature field
            # (final after the blend stage) is already travelling while the per-Gaussian kernel produced the other fields
            pk.all_reduce_finish(grp)
        v = pk.views
        return (v["dL_dmeans3D"], m2d, v.get("dL_dsh").view(P, M, 3) if M else None, v.get("dL_dcolors"),
                v.get("dL_dfeature") if ctx.include else None, v["dL_dopacity"].view(ctx.opac_shape),
                v["dL_dscales"], v["dL_drotations"], None, None, None)


class _RasterizeViewsLoss(torch.autograd.Function):
    """render_views with the loss heads fused into the forward blend's epilogue: returns the per-view
    L2 colour loss and cosine embedding loss of NeuralRenderer.forward (neural_rendering.py:300-318, loss.py:12-23); the
    images come back too, but gradients flow through the two losses only."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, feature, opacities, scales, rotations, views, target_rgb, target_embed, group):
        dev = means3D.device
        V, P = len(views), means3D.shape[0]
        H, W = int(views[0].image_height), int(views[0].image_width)
        if any((int(s.image_height), int(s.image_width)) != (H, W) for s in views):
            raise ValueError("render_views needs one image size for all views")
        include = feature is not None and feature.numel() > 0
        F = int(feature.shape[1]) if include else 0
        means3D, sh, colors_precomp, feature = (_prep(t, dev) for t in (means3D, sh, colors_precomp, feature))
        opacities, scales, rotations = (_prep(t, dev) for t in (opacities, scales, rotations))
        target_rgb = _prep(target_rgb, dev)
        target_embed = _prep(target_embed, dev) if (include and target_embed is not None) else None
        if tuple(target_rgb.shape) != (V, 3, H, W) or (target_embed is not None and tuple(target_embed.shape) != (V, F, H, W)):
            raise ValueError("targets must be [V,3,H,W] and [V,F,H,W]")
        opts = dict(dtype=torch.float32, device=dev)
        color = torch.empty((V, 3, H, W), **opts)
        feat_img = torch.empty((V, F, H, W), **opts) if include else None
        cot_rgb = torch.empty((V, 3, H, W), **opts)
        cot_embed = torch.empty((V, F, H, W), **opts) if target_embed is not None else (torch.zeros((V, F, H, W), **opts) if include else None)
        loss_acc = torch.empty((V, 2), **opts)
        degree = views[0].sh_degree
        outs, streams = rasterize_views_raw(views, means3D, colors_precomp, feature, opacities, scales, rotations,
                                            views[0].scale_modifier, None, sh, degree, include, debug=views[0].debug,
                                            out_buffers=(color, feat_img, None),
                                            loss_heads=(target_rgb, target_embed, cot_rgb, cot_embed, loss_acc))
        N = float(H * W)
        loss_rgb = loss_acc[:, 0] / (3.0 * N)
        loss_embed = 1.0 - loss_acc[:, 1] / N if target_embed is not None else torch.zeros((V,), **opts)
        radii = torch.stack([o[3] for o in outs])
        ctx.views, ctx.streams, ctx.group = views, streams, group
        # capacity, radii and the three state buffers per view -- never the images: they are views of this node's outputs, and an
        # output kept on ctx is a reference cycle (output -> grad_fn -> ctx -> output) that leaks the whole step
        ctx.state = [(o[0], o[3], o[4], o[5], o[6]) for o in outs]
        ctx.include, ctx.F, ctx.degree = include, F, degree
        ctx.opac_shape = opacities.shape
        ctx.cots = (cot_rgb, cot_embed)
        ctx.save_for_backward(means3D, sh if sh is not None else torch.empty(0, device=dev),
                              colors_precomp if colors_precomp is not None else torch.empty(0, device=dev),
                              feature if include else torch.empty(0, device=dev), scales, rotations)
        # ONE call: a second mark_non_differentiable replaces the first
        ctx.mark_non_differentiable(*([radii, color] + ([feat_img] if include else [])))
        empty = torch.empty(0, device=dev)
        return loss_rgb, loss_embed, color, (feat_img if include else empty), radii

    @staticmethod
    def backward(ctx, g_rgb, g_embed, _gc, _gf, _gr):
        means3D, sh, colors_precomp, feature, scales, rotations = ctx.saved_tensors
        dev = means3D.device
        views, V, P = ctx.views, len(ctx.views), means3D.shape[0]
        M = sh.shape[1] if sh.numel() else 0
        use_colors = colors_precomp.numel() > 0
        z = torch.zeros((V,), dtype=torch.float32, device=dev)
        scale = torch.stack([g_rgb if g_rgb is not None else z, g_embed if g_embed is not None else z], 1).contiguous()
        cot_rgb, cot_embed = ctx.cots
        pk = PackedGradients(P, ctx.F, M, dev, colors=use_colors, zero=False)
        grp = None if ctx.group in (None, True) else ctx.group
        m2d = torch.empty((V, P, 3), dtype=torch.float32, device=dev)
        outs = [(R, None, None, radii, geom, binb, img) for (R, radii, geom, binb, img) in ctx.state]
        rasterize_views_backward_raw(views, outs, ctx.streams, cot_rgb, cot_embed if ctx.include else None, means3D,
                                     colors_precomp if use_colors else None, feature if ctx.include else None, scales, rotations,
                                     views[0].scale_modifier, None, sh if M else None, ctx.degree, ctx.include, debug=views[0].debug,
                                     accumulate_into=pk.views, means2D_per_view=m2d, cot_scale=scale,
                                     after_blend=(lambda: pk.all_reduce_begin(grp)) if ctx.group is not None else None)
        if ctx.group is not None:
            pk.all_reduce_finish(grp)
        v = pk.views
        return (v["dL_dmeans3D"], m2d, v.get("dL_dsh").view(P, M, 3) if M else None, v.get("dL_dcolors"),
                v.get("dL_dfeature") if ctx.include else None, v["dL_dopacity"].view(ctx.opac_shape),
                v["dL_dscales"], v["dL_drotations"], None, None, None, None)


def render_views(cameras, pts_xyz, rotations, scales, opacity, bg_color=(0.0, 0.0, 0.0), pts_rgb=None, features_color=None,
                 features_language=None, view_ids=None, return_depth=False, sync_gradients=None, normalize_feature=True,
                 targets=None):
    """Render several cameras of one Gaussian cloud in one autograd node.

    cameras: `cameras.CameraBatch` (or a sequence of GaussianRasterizationSettings); view_ids selects a subset (e.g. this
    rank's shard, `parallel.shard_views`).  Gaussian arguments as `render()`.  sync_gradients: None (single GPU), True
    (all-reduce over the default process group) or a process group -- the per-Gaussian gradients of this rank's views are
    then summed over ranks inside the backward with one collective.  normalize_feature=False skips the reference's
    per-render feature normalisation (for callers that already hold unit features, e.g. `activate_gaussians` output).
    targets: {"rgb": [V,3,H,W], "embed": [V,F,H,W] (optional)} fuses ManiGaussian's loss heads into the forward blend's
    epilogue: the result then also holds "loss_rgb" and "loss_embed" ([V] each: l2_loss and cosine_loss of
    agents/manigaussian_bc/loss.py per view; combine them as neural_rendering.py:300-318 does) through which -- and only
    through which -- gradients flow; "render"/"render_embed" are returned detached.
    Returns {"render" [V,3,H,W], "render_embed" [V,F,H,W] | None, "depth" [V,H,W] (if asked), "viewspace_points" [V,P,3]
    (its .grad holds each view's screen-space mean gradients), "radii" [V,P] int32}."""
    device = pts_xyz.device
    include = features_language is not None
    sh_degree = 3 if features_color is None else 1
    if isinstance(cameras, CameraBatch):
        ids = list(range(len(cameras))) if view_ids is None else list(view_ids)
        bg = _background(bg_color, device)
        views = tuple(cameras.settings(i, bg, sh_degree, include) for i in ids)
    else:
        views = tuple(cameras if view_ids is None else [cameras[i] for i in view_ids])
    if not views:
        raise ValueError("render_views needs at least one view")
    if features_color is None and pts_rgb is None:
        raise ValueError("Please provide excatly one of either SHs or precomputed colors!")
    feat = None
    if include:
        feat = normalize_features(features_language) if normalize_feature else features_language
    V, P = len(views), pts_xyz.shape[0]
    screenspace_points = torch.zeros((V, P, 3), dtype=torch.float32, device=device, requires_grad=True)
    if targets is not None:
        if return_depth:
            raise ValueError("targets and return_depth cannot be combined")
        loss_rgb, loss_embed, color, feat_img, radii = _RasterizeViewsLoss.apply(
            pts_xyz, screenspace_points, features_color, pts_rgb if features_color is None else None, feat, opacity, scales,
            rotations, views, targets["rgb"], targets.get("embed"), sync_gradients)
        return {"render": color, "render_embed": feat_img if include else None, "viewspace_points": screenspace_points,
                "radii": radii, "loss_rgb": loss_rgb, "loss_embed": loss_embed}
    if V == 1 and sync_gradients is None:
        # a single view needs neither side streams nor the packed accumulation buffer: the single-view operator is leaner
        out = GaussianRasterizer(views[0], return_depth=return_depth)(
            means3D=pts_xyz, means2D=screenspace_points[0], shs=features_color,
            colors_precomp=pts_rgb if features_color is None else None, language_feature_precomp=feat, opacities=opacity,
            scales=scales, rotations=rotations, cov3D_precomp=None)
        color, feat_img, radii = out[0].unsqueeze(0), out[1].unsqueeze(0), out[2].unsqueeze(0)
        depth = out[3].unsqueeze(0) if return_depth else None
    else:
        color, feat_img, radii, depth = _RasterizeViews.apply(
            pts_xyz, screenspace_points, features_color, pts_rgb if features_color is None else None, feat, opacity, scales,
            rotations, views, bool(return_depth), sync_gradients)
    ret = {"render": color, "render_embed": feat_img if include else None, "viewspace_points": screenspace_points,
           "radii": radii}
    if return_depth:
        ret["depth"] = depth
    return ret
