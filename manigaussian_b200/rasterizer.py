"""Host-side mirror of the reference's Python operator for the rasterizer hot path.

Same names, argument order, return values and error behaviour as
DGR/diff_gaussian_rasterization/__init__.py (GaussianRasterizationSettings :166-179,
GaussianRasterizer :181-233, _RasterizeGaussians :46-164) and the pybind functions of DGR/ext.cpp:14-18
(rasterize_gaussians / rasterize_gaussians_backward / mark_visible, DGR/rasterize_points.cu:36-247), so that
ManiGaussian's agents/manigaussian_bc/gaussian_renderer/__init__.py:14-94 runs unchanged.  PyTorch is used for
device memory, streams and autograd plumbing only; all computation happens in the sm_100a library behind
include/mgs_rasterizer.h.

Supersets of the reference: the feature width F is read from the tensor at run time (0..32) instead of being
compiled in (config.h:16), and `return_depth=True` adds a view-space depth plane.
"""
from typing import NamedTuple

import os

import torch
import torch.nn as nn

from . import _binding as _b


def _ptr(t):
    """Device pointer of a tensor; None / empty tensor -> NULL (the reference's `data_ptr() == nullptr`)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _prep(t, device):
    """float32, contiguous, on `device` (the reference calls .contiguous() on every input, rasterize_points.cu:104-125)."""
    if t is None:
        return None
    if t.numel() == 0:
        return t
    if t.dtype != torch.float32 or t.device != device or not t.is_contiguous():
        t = t.to(device=device, dtype=torch.float32).contiguous()
    return t


class _Alloc:
    """State-buffer allocation backed by torch's caching allocator (replaces resizeFunctional, rasterize_points.cu:27-33).
    One process-wide C callback serves every call; the `user` pointer selects the _Alloc instance."""
    _live = {}
    _next = [1]

    def __init__(self, device):
        self.device = device
        self.tensor = None
        self.error = None  # exception raised inside the C callback (ctypes would swallow it): re-raised by the caller
        self.key = _Alloc._next[0]
        _Alloc._next[0] += 1
        _Alloc._live[self.key] = self

    def release(self):
        _Alloc._live.pop(self.key, None)
        if self.tensor is None:
            self.tensor = torch.empty(0, dtype=torch.uint8, device=self.device)
        return self.tensor

    def alloc(self, nbytes):
        # Round large requests up to a geometric bucket (<= 25 % slack): the binning state's size follows the
        # data-dependent instance count, and near-equal sizes must hit the same cached block or the caching
        # allocator keeps falling back to cudaMalloc for many iterations.
        n = max(int(nbytes), 1)
        if n > (1 << 20):
            q = 1 << max(20, n.bit_length() - 3)
            n = (n + q - 1) // q * q
        self.tensor = torch.empty(n, dtype=torch.uint8, device=self.device)
        return self.tensor.data_ptr()


def _alloc_trampoline(user, nbytes):
    a = _Alloc._live.get(int(user or 0))
    if a is None:
        return 0
    try:
        return a.alloc(nbytes)
    except BaseException as ex:  # e.g. torch.cuda.OutOfMemoryError: hand it to the caller instead of losing it in ctypes
        a.error = ex
        return 0


def _checked(allocs, rc, what):
    """_b.check that re-raises an exception stashed by an allocator callback and always releases the allocators"""
    err = next((a.error for a in allocs if a.error is not None), None)
    if err is not None:
        for a in allocs:
            a.release()
        raise err
    try:
        return _b.check(rc, what)
    except Exception:
        for a in allocs:
            a.release()
        raise


_ALLOC_CB = _b.ALLOC_FN(_alloc_trampoline)


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def rasterize_gaussians_raw(bg, means3D, colors, language_feature, opacity, scales, rotations, scale_modifier,
                            cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width,
                            sh, degree, campos, prefiltered, debug, include_feature, return_depth=False):
    """Same contract as the reference's `_C.rasterize_gaussians` (DGR/rasterize_points.cu:36-128).

    Returns (num_rendered, color [3,H,W], feature [F,H,W] or [1], radii [P] int32, geomBuffer, binningBuffer,
    imgBuffer) and, when return_depth, an extra trailing depth [H,W] tensor.
    """
    if means3D.ndimension() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("manigaussian_b200 runs on CUDA tensors only (there is no CPU path)")
    L = _b.lib()
    dev = means3D.device
    P, H, W = means3D.size(0), int(image_height), int(image_width)
    means3D, colors, language_feature, opacity = (_prep(x, dev) for x in (means3D, colors, language_feature, opacity))
    scales, rotations, cov3D_precomp, sh = (_prep(x, dev) for x in (scales, rotations, cov3D_precomp, sh))
    bg, viewmatrix, projmatrix, campos = (_prep(x, dev) for x in (bg, viewmatrix, projmatrix, campos))
    F = 0
    if include_feature and language_feature is not None and language_feature.numel() > 0:
        F = language_feature.size(1)
    M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
    out_color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
    out_feature = torch.empty((F, H, W), dtype=torch.float32, device=dev) if include_feature else \
        torch.zeros((1,), dtype=torch.float32, device=dev)
    out_depth = torch.empty((H, W), dtype=torch.float32, device=dev) if return_depth else None
    radii = torch.empty((P,), dtype=torch.int32, device=dev)
    ga, ba, ia = _Alloc(dev), _Alloc(dev), _Alloc(dev)
    rendered = 0
    if P != 0:
        with torch.cuda.device(dev):
            rendered = _checked((ga, ba, ia), L.mgs_forward(
                _ALLOC_CB, ga.key, _ALLOC_CB, ba.key, _ALLOC_CB, ia.key,
                P, int(degree), M, F,
                _ptr(bg), W, H,
                _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(language_feature) if F else None,
                _ptr(opacity), _ptr(scales), float(scale_modifier), _ptr(rotations),
                _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos),
                float(tan_fovx), float(tan_fovy), int(bool(prefiltered)),
                _ptr(out_color), _ptr(out_feature) if F else None, _ptr(out_depth), _ptr(radii),
                int(bool(debug)), _stream(dev)), "mgs_forward")
    else:
        out_color.zero_()
        if include_feature:
            out_feature.zero_()
        if out_depth is not None:
            out_depth.zero_()
    ret = (rendered, out_color, out_feature, radii, ga.release(), ba.release(), ia.release())
    return ret + (out_depth,) if return_depth else ret


def rasterize_gaussians_backward_raw(bg, means3D, radii, colors, language_feature, scales, rotations, scale_modifier,
                                     cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color,
                                     dL_dout_language_feature, sh, degree, campos, geomBuffer, R, binningBuffer,
                                     imageBuffer, debug, include_feature, dL_dout_depth=None, accumulate_into=None):
    """Same contract as the reference's `_C.rasterize_gaussians_backward` (DGR/rasterize_points.cu:131-225):
    returns (dL_dmeans2D, dL_dcolors, dL_dlanguage_feature, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh,
    dL_dscales, dL_drotations).

    accumulate_into: optional dict of preallocated fp32 tensors keyed like manigaussian_b200.parallel.FIELDS
    (dL_dmeans3D, dL_dmeans2D, dL_dscales, dL_drotations, dL_dopacity, dL_dsh, dL_dfeature[, dL_dcolors]); the gradients of
    this view are then ADDED to what they hold (C ABI `accumulate=1`; calls sharing buffers must be issued on one stream)
    and the same tensors are returned.  For several views of one cloud prefer rasterize_views_backward_raw."""
    L = _b.lib()
    dev = means3D.device
    P = means3D.size(0)
    H, W = dL_dout_color.size(1), dL_dout_color.size(2)
    means3D, colors, language_feature = (_prep(x, dev) for x in (means3D, colors, language_feature))
    scales, rotations, cov3D_precomp, sh = (_prep(x, dev) for x in (scales, rotations, cov3D_precomp, sh))
    bg, viewmatrix, projmatrix, campos = (_prep(x, dev) for x in (bg, viewmatrix, projmatrix, campos))
    dL_dout_color = _prep(dL_dout_color, dev)
    dL_dout_depth = _prep(dL_dout_depth, dev)
    F = 0
    if include_feature and language_feature is not None and language_feature.numel() > 0:
        F = language_feature.size(1)
        dL_dout_language_feature = _prep(dL_dout_language_feature, dev)
    M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
    opts = dict(dtype=torch.float32, device=dev)
    acc = accumulate_into
    if acc is not None:
        need = ["dL_dmeans3D", "dL_dmeans2D", "dL_dopacity"]
        need += ["dL_dscales", "dL_drotations"] if (scales is not None and scales.numel() != 0) else []
        need += ["dL_dcolors"] if (colors is not None and colors.numel() != 0) else []
        missing = [k for k in need if acc.get(k) is None]
        if missing:
            raise ValueError(f"accumulate_into lacks {missing}: gradients of inputs of this render would be dropped")
        dL_dmeans3D, dL_dmeans2D, dL_dopacity = acc["dL_dmeans3D"], acc["dL_dmeans2D"], acc["dL_dopacity"]
        dL_dscales, dL_drotations = acc.get("dL_dscales"), acc.get("dL_drotations")
        dL_dsh = acc.get("dL_dsh") if M else torch.empty((P, 0, 3), **opts)
        dL_dfeature = acc.get("dL_dfeature") if F else torch.zeros((1,), **opts)
        dL_dcolors, dL_dcov3D = acc.get("dL_dcolors"), None
    else:
        dL_dmeans3D = torch.empty((P, 3), **opts)
        dL_dmeans2D = torch.empty((P, 3), **opts)
        dL_dcolors = torch.empty((P, 3), **opts)
        dL_dfeature = torch.empty((P, F), **opts) if F else torch.zeros((1,), **opts)
        dL_dopacity = torch.empty((P, 1), **opts)
        dL_dcov3D = torch.empty((P, 6), **opts)
        dL_dsh = torch.empty((P, M, 3), **opts)
        dL_dscales = torch.empty((P, 3), **opts)
        dL_drotations = torch.empty((P, 4), **opts)
    if P != 0:
        scratch = torch.empty((L.mgs_backward_scratch_bytes(P),), dtype=torch.uint8, device=dev)
        has_sr = scales is not None and scales.numel() != 0
        with torch.cuda.device(dev):
            _b.check(L.mgs_backward(
                P, int(degree), M, F, int(R),
                _ptr(bg), W, H,
                _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(language_feature) if F else None,
                _ptr(scales), float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp),
                _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos), float(tan_fovx), float(tan_fovy),
                _ptr(radii), _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer),
                _ptr(dL_dout_color), _ptr(dL_dout_language_feature) if F else None, _ptr(dL_dout_depth),
                _ptr(dL_dmeans2D), None, _ptr(dL_dopacity), _ptr(dL_dcolors), _ptr(dL_dfeature) if F else None,
                _ptr(dL_dmeans3D), _ptr(dL_dcov3D), _ptr(dL_dsh) if M else None,
                _ptr(dL_dscales) if has_sr else None, _ptr(dL_drotations) if has_sr else None,
                _ptr(scratch), int(acc is not None), int(bool(debug)), _stream(dev)), "mgs_backward")
        if not has_sr and acc is None:
            dL_dscales.zero_()
            dL_drotations.zero_()
    return (dL_dmeans2D, dL_dcolors, dL_dfeature, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)


def mark_visible_raw(means3D, viewmatrix, projmatrix):
    """Same contract as the reference's `_C.mark_visible` (DGR/rasterize_points.cu:227-247)."""
    L = _b.lib()
    dev = means3D.device
    P = means3D.size(0)
    means3D, viewmatrix, projmatrix = (_prep(x, dev) for x in (means3D, viewmatrix, projmatrix))
    present = torch.zeros((P,), dtype=torch.uint8, device=dev)
    if P != 0:
        with torch.cuda.device(dev):
            _b.check(L.mgs_mark_visible(P, _ptr(means3D), _ptr(viewmatrix), _ptr(projmatrix), _ptr(present), _stream(dev)),
                     "mgs_mark_visible")
    return present.bool()


def state_array(which, name, state, a0, a1=0):
    """Device address of a named array inside an opaque state buffer (tests only)."""
    import ctypes
    out = ctypes.c_void_p()
    _b.check(_b.lib().mgs_state_array(which.encode(), name.encode(), state.data_ptr(), int(a0), int(a1), ctypes.byref(out)),
             "mgs_state_array")
    return out.value


def cpu_deep_copy_tuple(input_tuple):
    return tuple(item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, language_feature_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, raster_settings, return_depth=False):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, language_feature_precomp, opacities, scales,
                                     rotations, cov3Ds_precomp, raster_settings, return_depth)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, language_feature_precomp, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings, return_depth=False):
        s = raster_settings
        args = (s.bg, means3D, colors_precomp, language_feature_precomp, opacities, scales, rotations, s.scale_modifier,
                cov3Ds_precomp, s.viewmatrix, s.projmatrix, s.tanfovx, s.tanfovy, s.image_height, s.image_width, sh,
                s.sh_degree, s.campos, s.prefiltered, s.debug, s.include_feature)
        ctx.async_view = None
        if s.debug:
            cpu_args = cpu_deep_copy_tuple(args)  # copy before they can be corrupted
            try:
                out = rasterize_gaussians_raw(*args, return_depth=return_depth)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        elif means3D.is_cuda and means3D.size(0) > 0:
            # no host synchronisation: the binning state is sized from the instance counts of earlier calls (the reference reads
            # the count back on every call, rasterizer_impl.cu:284); the operator itself never exposes the count
            outs, streams = rasterize_views_raw([s], means3D, colors_precomp, language_feature_precomp, opacities, scales, rotations,
                                                s.scale_modifier, cov3Ds_precomp, sh, s.sh_degree, s.include_feature,
                                                return_depth=return_depth)
            out = outs[0]
            # only the capacity and the streams: the images and radii in `outs` are OUTPUTS of this node, and an output kept on
            # ctx is a reference cycle (output -> grad_fn -> ctx -> output) that no collector frees -- the state buffers and the
            # radii come back through saved_tensors in the backward
            ctx.async_view = (int(out[0]), streams)
        else:
            out = rasterize_gaussians_raw(*args, return_depth=return_depth)
        num_rendered, color, language_feature, radii, geomBuffer, binningBuffer, imgBuffer = out[:7]
        depth = out[7] if return_depth else None
        ctx.raster_settings = s
        ctx.num_rendered = num_rendered
        ctx.return_depth = return_depth
        ctx.save_for_backward(colors_precomp, language_feature_precomp, means3D, scales, rotations, cov3Ds_precomp, radii,
                              sh, geomBuffer, binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        if return_depth:
            return color, language_feature, radii, depth
        return color, language_feature, radii

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_language_feature, _radii=None, grad_out_depth=None):
        s = ctx.raster_settings
        (colors_precomp, language_feature_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
         binningBuffer, imgBuffer) = ctx.saved_tensors
        if grad_out_color is None:
            grad_out_color = torch.zeros((3, s.image_height, s.image_width), dtype=torch.float32, device=means3D.device)
        if s.include_feature and grad_out_language_feature is None:
            grad_out_language_feature = torch.zeros((language_feature_precomp.size(1), s.image_height, s.image_width),
                                                    dtype=torch.float32, device=means3D.device)
        args = (s.bg, means3D, radii, colors_precomp, language_feature_precomp, scales, rotations, s.scale_modifier,
                cov3Ds_precomp, s.viewmatrix, s.projmatrix, s.tanfovx, s.tanfovy, grad_out_color,
                grad_out_language_feature, sh, s.sh_degree, s.campos, geomBuffer, ctx.num_rendered, binningBuffer,
                imgBuffer, s.debug, s.include_feature)
        depth_grad = grad_out_depth if ctx.return_depth else None
        if ctx.async_view is not None:
            cap, streams = ctx.async_view
            outs = [ViewOut((cap, None, None, radii, geomBuffer, binningBuffer, imgBuffer))]
            grads = rasterize_views_backward_raw(
                [s], outs, streams, [grad_out_color], [grad_out_language_feature] if s.include_feature else None, means3D,
                colors_precomp, language_feature_precomp, scales, rotations, s.scale_modifier, cov3Ds_precomp, sh, s.sh_degree,
                s.include_feature, grads_depth=[depth_grad] if depth_grad is not None else None)
        elif s.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                grads = rasterize_gaussians_backward_raw(*args, dL_dout_depth=depth_grad)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            grads = rasterize_gaussians_backward_raw(*args, dL_dout_depth=depth_grad)
        (grad_means2D, grad_colors_precomp, grad_language_feature_precomp, grad_opacities, grad_means3D,
         grad_cov3Ds_precomp, grad_sh, grad_scales, grad_rotations) = grads

        def _or_none(g, inp):
            return g if (inp is not None and inp.numel() != 0) else None

        return (grad_means3D, grad_means2D, _or_none(grad_sh, sh), _or_none(grad_colors_precomp, colors_precomp),
                _or_none(grad_language_feature_precomp, language_feature_precomp) if s.include_feature else None,
                grad_opacities, _or_none(grad_scales, scales), _or_none(grad_rotations, rotations),
                _or_none(grad_cov3Ds_precomp, cov3Ds_precomp), None, None)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    include_feature: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings, return_depth=False):
        super().__init__()
        self.raster_settings = raster_settings
        self.return_depth = return_depth

    def markVisible(self, positions):
        # Mark visible points (based on frustum culling for camera) with a boolean
        with torch.no_grad():
            s = self.raster_settings
            return mark_visible_raw(positions, s.viewmatrix, s.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, language_feature_precomp=None,
                scales=None, rotations=None, cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        empty = torch.Tensor([])
        shs = empty if shs is None else shs
        colors_precomp = empty if colors_precomp is None else colors_precomp
        language_feature_precomp = empty if language_feature_precomp is None else language_feature_precomp
        scales = empty if scales is None else scales
        rotations = empty if rotations is None else rotations
        cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, language_feature_precomp, opacities, scales,
                                   rotations, cov3D_precomp, self.raster_settings, self.return_depth)


# ----------------------------------------------------------------------------------------------------------------------
# Multi-view batches (no counterpart in the reference, which renders one view per call and
# blocks the host on every view's instance count, rasterizer_impl.cu:284).  All views share one Gaussian cloud.  One C
# call (mgs_forward_views / mgs_backward_views) enqueues every view on its own CUDA stream, forks from and joins back
# into the caller's current stream, and NEVER synchronises with the host: the per-view binning state is sized from the
# instance counts observed on earlier calls (plus slack), the real count stays on the device, and {count, overflow flag}
# come back asynchronously through pinned memory to be looked at on a LATER call.  A step built from these calls
# allocates through torch's caching allocator only and can be captured in a CUDA graph.
import warnings

_VIEW_STREAMS = {}
_CAPACITY = {}        # (device index, P, W, H, slot) -> instance capacity to use for the next call
_PENDING_STATUS = []  # [(key, pinned int32[2], cuda event, capacity used)] of forwards whose status has not been read yet
_SIZES = {}
_STATUS_RING = {"buf": None, "next": 0}   # pinned {count, overflow} slots, handed out round-robin (pin_memory() per call would
_STATUS_SLOTS = 4096                       # cost a cudaHostAlloc and a device synchronisation each time)


def _status_slots(n):
    """n consecutive pinned int32[2] slots.  A slot is reused after _STATUS_SLOTS later views: by then its forward finished."""
    if _STATUS_RING["buf"] is None:
        _STATUS_RING["buf"] = torch.zeros((_STATUS_SLOTS, 2), dtype=torch.int32).pin_memory()
    i = _STATUS_RING["next"]
    if i + n > _STATUS_SLOTS:
        i = 0
    _STATUS_RING["next"] = i + n
    return _STATUS_RING["buf"][i:i + n]


def _view_streams(device, n):
    if os.environ.get("MGS_ONE_STREAM"):  # measurement hook: every view on the caller's stream (no overlap between views)
        return [torch.cuda.current_stream(device)] * n
    key = (device.index if device.index is not None else torch.cuda.current_device())
    pool = _VIEW_STREAMS.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


def _round_capacity(n):
    """geometric buckets (<= 12.5 % slack) so that slowly varying counts keep hitting the same cached allocation"""
    n = max(int(n), 4096)
    q = 1 << max(10, n.bit_length() - 4)
    return (n + q - 1) // q * q


def _poll_status(block=False):
    """Fold the instance counts of finished forwards into the capacity table (never waits unless block=True)."""
    keep = []
    for key, st, ev, cap in _PENDING_STATUS:
        if block:
            ev.synchronize()
        if not ev.query():
            keep.append((key, st, ev, cap))
            continue
        R, over = int(st[0]), int(st[1])
        want = _round_capacity(R * 1.25 + 4096)
        if over:
            warnings.warn(f"manigaussian_b200: a view produced {R} tile instances but its binning state held {cap}; the farthest "
                          f"{R - cap} instances were dropped from that render (capacity raised for the next call)")
            want = _round_capacity(R * 1.5 + 4096)
        cur = _CAPACITY.get(key, 0)
        # grow at once; shrink only when the state is more than twice too large (stable sizes keep hitting the same cached
        # allocation, and a render with few instances must not starve the next one)
        if want > cur or want < cur // 2:
            _CAPACITY[key] = want
    _PENDING_STATUS[:] = keep


def reset_capacity_estimates():
    """Forget the instance-count history (tests; or after a drastic scene change to force re-calibration)."""
    _poll_status(block=True)
    _CAPACITY.clear()


class ViewOut(tuple):
    """Result of one view of rasterize_views_raw: (capacity, color, feature, radii, geomBuffer, binningBuffer, imgBuffer
    [, depth]) -- the layout of rasterize_gaussians_raw with the binning CAPACITY in place of the instance count, which is
    what the backward needs.  `.num_rendered()` waits for the forward and returns the true count; `.status` is the pinned
    {count, overflow} pair."""
    status = None
    event = None

    def num_rendered(self):
        self.event.synchronize()
        return int(self.status[0])

    def overflowed(self):
        self.event.synchronize()
        return bool(self.status[1])


def _state_sizes(L, P, W, H):
    key = (P, W, H)
    s = _SIZES.get(key)
    if s is None:
        s = (int(L.mgs_geometry_state_bytes(P)), int(L.mgs_image_state_bytes(W, H)), int(L.mgs_backward_scratch_bytes(P)))
        _SIZES[key] = s
    return s


def _calibrate(L, views, keys, dev, P, degree, M, means3D, sh, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, debug):
    """First call for a (cloud size, image size): one projection + scan per view with a host read of the instance count
    (the reference does this on EVERY call), only to seed the capacity table."""
    counts = torch.zeros(len(views), dtype=torch.int32).pin_memory()
    st = torch.cuda.current_stream(dev)
    keep = []
    for v, s in enumerate(views):
        if keys[v] in _CAPACITY:
            continue
        H, W = int(s.image_height), int(s.image_width)
        ga, ia = _Alloc(dev), _Alloc(dev)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        vm, pm, cp = (_prep(x, dev) for x in (s.viewmatrix, s.projmatrix, s.campos))
        _checked((ga, ia), L.mgs_forward_begin(_ALLOC_CB, ga.key, _ALLOC_CB, ia.key, P, int(degree), M, W, H, _ptr(means3D), _ptr(sh), _ptr(colors),
                                     _ptr(opacity), _ptr(scales), float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp), _ptr(vm),
                                     _ptr(pm), _ptr(cp), float(s.tanfovx), float(s.tanfovy), _ptr(radii), counts.data_ptr() + 4 * v,
                                     int(bool(debug)), st.cuda_stream), "mgs_forward_begin")
        keep.append((ga.release(), ia.release(), radii, vm, pm, cp))
    st.synchronize()
    for v in range(len(views)):
        if keys[v] not in _CAPACITY:
            _CAPACITY[keys[v]] = _round_capacity(int(counts[v]) * 1.25 + 4096)


def rasterize_views_raw(views, means3D, colors, language_feature, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                        sh, degree, include_feature, return_depth=False, debug=False, out_buffers=None, capacities=None,
                        loss_heads=None):
    """Forward of V views of one Gaussian cloud.  `views` is a sequence of GaussianRasterizationSettings (bg, viewmatrix,
    projmatrix, tanfov*, image size, campos are read per view).  Returns (outs, streams): outs[v] is a ViewOut; all work
    is joined into the caller's current stream before the call returns (`streams` are the per-view streams used).
    `out_buffers` = (color [V,3,H,W], feature [V,F,H,W] or None, depth [V,H,W] or None): render straight into slices of
    caller-owned batch tensors instead of per-view allocations.  `capacities` (per-view instance capacities) overrides
    the history-based sizing of the binning state.  `loss_heads` = (target_rgb [V,3,H,W], target_embed [V,F,H,W] or None,
    cot_rgb [V,3,H,W], cot_embed [V,F,H,W] or None, loss_acc [V,2]): the forward blend's epilogue also evaluates the L2 colour
    head and the cosine embedding head against the targets (C ABI mgs_view.target_*), leaving the loss sums in loss_acc and
    the cotangent planes ready for rasterize_views_backward_raw."""
    L = _b.lib()
    dev = means3D.device
    if not means3D.is_cuda:
        raise RuntimeError("manigaussian_b200 runs on CUDA tensors only (there is no CPU path)")
    P = means3D.size(0)
    if P == 0:
        raise RuntimeError("rasterize_views_raw needs at least one Gaussian")
    means3D, colors, language_feature, opacity = (_prep(x, dev) for x in (means3D, colors, language_feature, opacity))
    scales, rotations, cov3D_precomp, sh = (_prep(x, dev) for x in (scales, rotations, cov3D_precomp, sh))
    F = language_feature.size(1) if (include_feature and language_feature is not None and language_feature.numel() > 0) else 0
    M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
    V = len(views)
    main = torch.cuda.current_stream(dev)
    streams = _view_streams(dev, V) if V > 1 else [main]  # a single view needs no side stream
    di = dev.index if dev.index is not None else torch.cuda.current_device()
    keys = [(di, P, int(s.image_width), int(s.image_height), v) for v, s in enumerate(views)]
    capturing = torch.cuda.is_current_stream_capturing()  # inside a CUDA-graph capture: no event queries, no calibration
    with torch.cuda.device(dev):
        if not capturing:
            _poll_status()
        if capacities is None and any(k not in _CAPACITY for k in keys):
            if capturing:
                raise RuntimeError("render the views once before capturing them in a CUDA graph (the binning capacity is learned "
                                   "from an eager call) or pass `capacities`")
            _calibrate(L, views, keys, dev, P, degree, M, means3D, sh, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                       debug)
        arr = (_b.View * V)()
        status = _status_slots(V)
        keep, outs = [], []
        for v, s in enumerate(views):
            H, W = int(s.image_height), int(s.image_width)
            cap = int(capacities[v]) if capacities is not None else _CAPACITY[keys[v]]
            gbytes, ibytes, _ = _state_sizes(L, P, W, H)
            geom = torch.empty(gbytes, dtype=torch.uint8, device=dev)
            img = torch.empty(ibytes, dtype=torch.uint8, device=dev)
            binb = torch.empty(int(L.mgs_binning_state_bytes(cap)), dtype=torch.uint8, device=dev)
            radii = torch.empty((P,), dtype=torch.int32, device=dev)
            bg, vm, pm, cp = (_prep(x, dev) for x in (s.bg, s.viewmatrix, s.projmatrix, s.campos))
            if out_buffers is not None:
                out_color = out_buffers[0][v]
                out_feature = out_buffers[1][v] if (include_feature and F) else torch.zeros((1,), dtype=torch.float32, device=dev)
                out_depth = out_buffers[2][v] if return_depth else None
            else:
                out_color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
                out_feature = torch.empty((F, H, W), dtype=torch.float32, device=dev) if include_feature else \
                    torch.zeros((1,), dtype=torch.float32, device=dev)
                out_depth = torch.empty((H, W), dtype=torch.float32, device=dev) if return_depth else None
            w = arr[v]
            w.viewmatrix, w.projmatrix, w.cam_pos, w.background = _ptr(vm), _ptr(pm), _ptr(cp), _ptr(bg)
            w.tan_fovx, w.tan_fovy, w.width, w.height = float(s.tanfovx), float(s.tanfovy), W, H
            w.geometry_state, w.binning_state, w.image_state, w.binning_capacity = geom.data_ptr(), binb.data_ptr(), img.data_ptr(), cap
            w.out_color, w.out_feature, w.out_depth, w.radii = _ptr(out_color), (_ptr(out_feature) if F else None), _ptr(out_depth), _ptr(radii)
            w.status = status.data_ptr() + 8 * v
            w.stream = streams[v].cuda_stream
            if loss_heads is not None:
                t_rgb, t_emb, c_rgb, c_emb, l_acc = loss_heads
                w.target_color, w.cot_color, w.loss_acc = _ptr(t_rgb[v]), _ptr(c_rgb[v]), _ptr(l_acc[v])
                if t_emb is not None and F:
                    w.target_feature, w.cot_feature = _ptr(t_emb[v]), _ptr(c_emb[v])
            keep.append((bg, vm, pm, cp))
            ret = (cap, out_color, out_feature, radii, geom, binb, img) + ((out_depth,) if return_depth else ())
            outs.append(ViewOut(ret))
        _b.check(L.mgs_forward_views(V, arr, P, int(degree), M, F, _ptr(means3D), _ptr(sh), _ptr(colors),
                                     _ptr(language_feature) if F else None, _ptr(opacity), _ptr(scales), float(scale_modifier),
                                     _ptr(rotations), _ptr(cov3D_precomp), 0, int(bool(debug)), main.cuda_stream), "mgs_forward_views")
        ev = torch.cuda.Event()
        ev.record(main)
        for v, o in enumerate(outs):
            o.status, o.event = status[v], ev
            if capacities is None and not capturing:
                _PENDING_STATUS.append((keys[v], status[v], ev, o[0]))
    return outs, streams


def rasterize_views_backward_raw(views, outs, streams, grads_color, grads_feature, means3D, colors, language_feature, scales,
                                 rotations, scale_modifier, cov3D_precomp, sh, degree, include_feature, grads_depth=None,
                                 debug=False, accumulate_into=None, means2D_per_view=None, accumulate=False, after_blend=None,
                                 cot_scale=None):
    """Backward of the V views rendered by rasterize_views_raw: every view's blend backward on its stream, then ONE
    per-Gaussian chain-rule kernel that sums over the views (mgs_backward_views).  Returns the 9-tuple of
    rasterize_gaussians_backward_raw holding the SUMS over the views (dL_dmeans2D: `means2D_per_view` [V,P,3] when given,
    else the sum).  `accumulate_into`: dict of preallocated fp32 tensors keyed like manigaussian_b200.parallel.FIELDS that
    receive the sums (e.g. the views of a PackedGradients buffer -- the all-reduce message); rows are overwritten unless
    accumulate=True.  `after_blend()` is called between the blend stage and the per-Gaussian stage, when dL_dfeature is final
    (multi-GPU: start its all-reduce there; it overlaps the per-Gaussian kernel).  `cot_scale` [V,2] (device): upstream
    gradients of the fused loss heads' two scalars per view, multiplied onto the colour / feature cotangent planes as they
    are loaded.  The caller's current stream holds the result on return."""
    L = _b.lib()
    dev = means3D.device
    P = means3D.size(0)
    means3D, colors, language_feature = (_prep(x, dev) for x in (means3D, colors, language_feature))
    scales, rotations, cov3D_precomp, sh = (_prep(x, dev) for x in (scales, rotations, cov3D_precomp, sh))
    F = language_feature.size(1) if (include_feature and language_feature is not None and language_feature.numel() > 0) else 0
    M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
    V = len(views)
    main = torch.cuda.current_stream(dev)
    opts = dict(dtype=torch.float32, device=dev)
    acc = accumulate_into or {}
    has_sr = scales is not None and scales.numel() != 0
    use_colors = colors is not None and colors.numel() != 0

    def out(name, shape, need=True):
        t = acc.get(name)
        if t is None and need:
            if accumulate_into is not None and name not in ("dL_dmeans2D", "dL_dcov3D"):
                raise ValueError(f"accumulate_into lacks '{name}', the gradient of an input of this render")
            t = torch.empty(shape, **opts)
        return t

    dL_dmeans3D = out("dL_dmeans3D", (P, 3))
    dL_dopacity = out("dL_dopacity", (P, 1))
    dL_dscales = out("dL_dscales", (P, 3), has_sr)
    dL_drotations = out("dL_drotations", (P, 4), has_sr)
    dL_dsh = out("dL_dsh", (P, M, 3), M > 0)
    dL_dfeature = out("dL_dfeature", (P, F), F > 0)
    dL_dcolors = out("dL_dcolors", (P, 3), use_colors)
    dL_dcov3D = out("dL_dcov3D", (P, 6), cov3D_precomp is not None and cov3D_precomp.numel() != 0)
    shared = means2D_per_view is None
    m2d = out("dL_dmeans2D", (P, 3)) if shared else means2D_per_view
    with torch.cuda.device(dev):
        arr = (_b.View * V)()
        keep = []
        for v, s in enumerate(views):
            o = outs[v]
            H, W = int(s.image_height), int(s.image_width)
            bg, vm, pm, cp = (_prep(x, dev) for x in (s.bg, s.viewmatrix, s.projmatrix, s.campos))
            gc = _prep(grads_color[v], dev)
            gf = _prep(grads_feature[v], dev) if (F and grads_feature is not None) else None
            gd = _prep(grads_depth[v], dev) if grads_depth is not None else None
            scratch = torch.empty((_state_sizes(L, P, W, H)[2],), dtype=torch.uint8, device=dev)
            w = arr[v]
            w.viewmatrix, w.projmatrix, w.cam_pos, w.background = _ptr(vm), _ptr(pm), _ptr(cp), _ptr(bg)
            w.tan_fovx, w.tan_fovy, w.width, w.height = float(s.tanfovx), float(s.tanfovy), W, H
            w.geometry_state, w.binning_state, w.image_state, w.binning_capacity = o[4].data_ptr(), o[5].data_ptr(), o[6].data_ptr(), int(o[0])
            w.radii = _ptr(o[3])
            w.dL_dpix, w.dL_dpix_F, w.dL_dpix_depth = _ptr(gc), _ptr(gf), _ptr(gd)
            w.blend_scratch = scratch.data_ptr()
            w.dL_dmean2D = _ptr(m2d) if shared else _ptr(m2d[v])
            w.stream = streams[v].cuda_stream
            if cot_scale is not None:
                w.cot_scale = cot_scale.data_ptr() + 8 * v
            keep.append((bg, vm, pm, cp, gc, gf, gd, scratch))
        for stages in ((1, 2) if after_blend is not None else (3,)):
            _b.check(L.mgs_backward_views(V, arr, P, int(degree), M, F, _ptr(means3D), _ptr(sh) if M else None, _ptr(colors),
                                          _ptr(language_feature) if F else None, _ptr(scales), float(scale_modifier), _ptr(rotations),
                                          _ptr(cov3D_precomp), _ptr(dL_dmeans3D), _ptr(dL_dopacity), _ptr(dL_dcolors),
                                          _ptr(dL_dfeature) if F else None, _ptr(dL_dcov3D), _ptr(dL_dsh) if M else None,
                                          _ptr(dL_dscales) if has_sr else None, _ptr(dL_drotations) if has_sr else None,
                                          int(shared), int(bool(accumulate)), stages, int(bool(debug)), main.cuda_stream),
                     "mgs_backward_views")
            if stages == 1:
                after_blend()
    return (m2d, dL_dcolors, dL_dfeature, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)
