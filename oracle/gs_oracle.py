"""TEST INFRASTRUCTURE -- ctypes/numpy front end of the CPU oracle (oracle/gs_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package (manigaussian_b200/) never does.

`forward()` / `backward()` chain the oracle's stage functions in the order of
CudaRasterizer::Rasterizer::forward / ::backward (DGR/cuda_rasterizer/rasterizer_impl.cu:198-355,
:359-463) and return every intermediate the reference keeps in its state buffers
(rasterizer_impl.h:29-65), so tests can compare stage by stage.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libgs_oracle.so")
    src = os.path.join(_HERE, "gs_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libgs_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.gso_scan.restype = C.c_uint32
        _LIB.gso_get_higher_msb.restype = C.c_uint32
        _LIB.gso_max_threads.restype = C.c_int
    return _LIB


def _p(a):
    """numpy array (or None) -> void* ."""
    if a is None:
        return C.c_void_p(0)
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def max_threads():
    return lib().gso_max_threads()


def set_threads(n):
    lib().gso_set_threads(int(n))


def get_higher_msb(n):
    return int(lib().gso_get_higher_msb(C.c_uint32(n)))


def forward(means3D, opacities, viewmatrix, projmatrix, campos, W, H, tanfovx, tanfovy, bg,
            scales=None, rotations=None, cov3D_precomp=None, shs=None, sh_degree=0,
            colors_precomp=None, feature=None, scale_modifier=1.0, stages_only=False):
    """Full forward pass.  `feature` is [P,F] or None (include_feature=False).  Returns a dict of numpy arrays."""
    L = lib()
    means3D = _f32(means3D)
    P = means3D.shape[0]
    opacities = _f32(opacities).reshape(-1)
    scales, rotations, cov3D_precomp = _f32(scales), _f32(rotations), _f32(cov3D_precomp)
    shs, colors_precomp, feature = _f32(shs), _f32(colors_precomp), _f32(feature)
    viewmatrix, projmatrix = _f32(viewmatrix).reshape(-1), _f32(projmatrix).reshape(-1)
    campos, bg = _f32(campos).reshape(-1), _f32(bg).reshape(-1)
    M = 0 if shs is None else shs.shape[1]
    F = 0 if feature is None else feature.shape[1]
    N = W * H
    gx, gy = (W + 15) // 16, (H + 15) // 16
    o = dict(P=P, W=W, H=H, F=F, M=M)
    o["radii"] = np.zeros(P, np.int32)
    o["means2D"] = np.zeros((P, 2), np.float32)
    o["depths"] = np.zeros(P, np.float32)
    o["cov3D"] = np.zeros((P, 6), np.float32)
    o["rgb"] = np.zeros((P, 3), np.float32)
    o["conic_opacity"] = np.zeros((P, 4), np.float32)
    o["clamped"] = np.zeros((P, 3), np.uint8)
    o["tiles_touched"] = np.zeros(P, np.uint32)
    L.gso_preprocess(C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(means3D), _p(scales), C.c_float(scale_modifier),
                     _p(rotations), _p(opacities), _p(shs), _p(cov3D_precomp), _p(colors_precomp), _p(viewmatrix),
                     _p(projmatrix), _p(campos), C.c_int(W), C.c_int(H), C.c_float(tanfovx), C.c_float(tanfovy),
                     _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["cov3D"]), _p(o["rgb"]),
                     _p(o["conic_opacity"]), _p(o["clamped"]), _p(o["tiles_touched"]))
    o["point_offsets"] = np.zeros(P, np.uint32)
    R = int(L.gso_scan(C.c_int(P), _p(o["tiles_touched"]), _p(o["point_offsets"]))) if P else 0
    o["num_rendered"] = R
    ku, vu = np.zeros(R, np.uint64), np.zeros(R, np.uint32)
    L.gso_duplicate_with_keys(C.c_int(P), _p(o["means2D"]), _p(o["depths"]), _p(o["point_offsets"]), _p(o["radii"]),
                              C.c_int(W), C.c_int(H), _p(ku), _p(vu))
    o["keys_unsorted"], o["values_unsorted"] = ku, vu
    ks, vs = np.zeros(R, np.uint64), np.zeros(R, np.uint32)
    bit = get_higher_msb(gx * gy)
    L.gso_sort_pairs(C.c_uint32(R), _p(ku), _p(vu), _p(ks), _p(vs), C.c_int(32 + bit))
    o["point_list_keys"], o["point_list"] = ks, vs
    o["ranges"] = np.zeros((gx * gy, 2), np.uint32)
    L.gso_identify_tile_ranges(C.c_uint32(R), _p(ks), C.c_int(gx * gy), _p(o["ranges"]))
    if stages_only:
        return o
    colors = colors_precomp if colors_precomp is not None else o["rgb"]
    o["final_T"] = np.zeros(N, np.float32)
    o["n_contrib"] = np.zeros(N, np.uint32)
    o["out_color"] = np.zeros((3, H, W), np.float32)
    o["out_feature"] = np.zeros((F, H, W), np.float32) if F else np.zeros((1,), np.float32)
    L.gso_render_forward(C.c_int(W), C.c_int(H), C.c_int(F), _p(o["ranges"]), _p(vs), _p(o["means2D"]), _p(colors),
                         _p(feature), _p(o["conic_opacity"]), _p(bg), _p(o["final_T"]), _p(o["n_contrib"]),
                         _p(o["out_color"]), _p(o["out_feature"]))
    return o


def backward(fw, dL_dcolor, dL_dfeature, means3D, viewmatrix, projmatrix, campos, tanfovx, tanfovy, bg,
             scales=None, rotations=None, cov3D_precomp=None, shs=None, sh_degree=0,
             colors_precomp=None, feature=None, scale_modifier=1.0):
    """Full backward pass given `fw` = forward()'s dict.  Returns the 9 gradients of
    RasterizeGaussiansBackwardCUDA (DGR/rasterize_points.cu:131-225) plus dL_dconic."""
    L = lib()
    P, W, H, F, M = fw["P"], fw["W"], fw["H"], fw["F"], fw["M"]
    means3D = _f32(means3D)
    scales, rotations, cov3D_precomp = _f32(scales), _f32(rotations), _f32(cov3D_precomp)
    shs, colors_precomp, feature = _f32(shs), _f32(colors_precomp), _f32(feature)
    viewmatrix, projmatrix = _f32(viewmatrix).reshape(-1), _f32(projmatrix).reshape(-1)
    campos, bg = _f32(campos).reshape(-1), _f32(bg).reshape(-1)
    dL_dcolor = _f32(dL_dcolor)
    dL_dfeature = _f32(dL_dfeature) if F else None
    g = dict(
        dL_dmeans2D=np.zeros((P, 3), np.float32), dL_dconic=np.zeros((P, 4), np.float32),
        dL_dopacity=np.zeros((P, 1), np.float32), dL_dcolors=np.zeros((P, 3), np.float32),
        dL_dfeature=np.zeros((P, F), np.float32) if F else np.zeros((1,), np.float32),
        dL_dmeans3D=np.zeros((P, 3), np.float32), dL_dcov3D=np.zeros((P, 6), np.float32),
        dL_dsh=np.zeros((P, M, 3), np.float32), dL_dscales=np.zeros((P, 3), np.float32),
        dL_drotations=np.zeros((P, 4), np.float32))
    colors = colors_precomp if colors_precomp is not None else fw["rgb"]
    L.gso_render_backward(C.c_int(P), C.c_int(W), C.c_int(H), C.c_int(F), _p(fw["ranges"]), _p(fw["point_list"]), _p(bg),
                          _p(fw["means2D"]), _p(fw["conic_opacity"]), _p(colors), _p(feature), _p(fw["final_T"]),
                          _p(fw["n_contrib"]), _p(dL_dcolor), _p(dL_dfeature), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]),
                          _p(g["dL_dopacity"]), _p(g["dL_dcolors"]), _p(g["dL_dfeature"]))
    cov3D = cov3D_precomp if cov3D_precomp is not None else fw["cov3D"]
    L.gso_preprocess_backward(C.c_int(P), C.c_int(sh_degree), C.c_int(M), _p(means3D), _p(fw["radii"]), _p(shs),
                              _p(fw["clamped"]), _p(scales), _p(rotations), C.c_float(scale_modifier), _p(cov3D),
                              _p(viewmatrix), _p(projmatrix), C.c_int(W), C.c_int(H), C.c_float(tanfovx),
                              C.c_float(tanfovy), _p(campos), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]),
                              _p(g["dL_dcolors"]), _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]), _p(g["dL_dsh"]),
                              _p(g["dL_dscales"]), _p(g["dL_drotations"]))
    return g


def mark_visible(means3D, viewmatrix, projmatrix):
    means3D = _f32(means3D)
    P = means3D.shape[0]
    out = np.zeros(P, np.uint8)
    lib().gso_mark_visible(C.c_int(P), _p(means3D), _p(_f32(viewmatrix).reshape(-1)), _p(_f32(projmatrix).reshape(-1)), _p(out))
    return out.astype(bool)
