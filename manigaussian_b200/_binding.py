"""ctypes binding of the C ABI in include/mgs_rasterizer.h.

The CUDA library is the product path: importing this module raises if it cannot be built or loaded.
There is no CPU or PyTorch fallback anywhere in this package.
"""
import ctypes as C
import os

from . import build as _build

_c_float_p = C.c_void_p  # device pointers are passed as integers
ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)



class View(C.Structure):
    """mirror of `mgs_view` (include/mgs_rasterizer.h)"""
    _fields_ = [("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("cam_pos", C.c_void_p), ("background", C.c_void_p),
                ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("width", C.c_int), ("height", C.c_int),
                ("geometry_state", C.c_void_p), ("binning_state", C.c_void_p), ("image_state", C.c_void_p),
                ("binning_capacity", C.c_int),
                ("out_color", C.c_void_p), ("out_feature", C.c_void_p), ("out_depth", C.c_void_p), ("radii", C.c_void_p),
                ("status", C.c_void_p),
                ("dL_dpix", C.c_void_p), ("dL_dpix_F", C.c_void_p), ("dL_dpix_depth", C.c_void_p),
                ("blend_scratch", C.c_void_p), ("dL_dmean2D", C.c_void_p), ("stream", C.c_void_p),
                ("target_color", C.c_void_p), ("target_feature", C.c_void_p), ("cot_color", C.c_void_p),
                ("cot_feature", C.c_void_p), ("loss_acc", C.c_void_p), ("cot_scale", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()
    if not os.path.exists(path):
        raise ImportError(f"manigaussian_b200: CUDA library missing at {path}")
    L = C.CDLL(path)
    L.mgs_abi_version.restype = C.c_int
    L.mgs_last_error.restype = C.c_char_p
    for name in ("mgs_geometry_state_bytes", "mgs_binning_state_bytes", "mgs_backward_scratch_bytes"):
        getattr(L, name).restype = C.c_size_t
        getattr(L, name).argtypes = [C.c_int]
    L.mgs_image_state_bytes.restype = C.c_size_t
    L.mgs_image_state_bytes.argtypes = [C.c_int, C.c_int]
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    L.mgs_forward.restype = C.c_int
    L.mgs_forward.argtypes = [ALLOC_FN, vp, ALLOC_FN, vp, ALLOC_FN, vp,
                              i, i, i, i,            # P D M F
                              vp, i, i,              # background, width, height
                              vp, vp, vp, vp,        # means3D shs colors feature
                              vp, vp, f, vp,         # opacities scales scale_modifier rotations
                              vp, vp, vp, vp,        # cov3D_precomp viewmatrix projmatrix cam_pos
                              f, f, i,               # tan_fovx tan_fovy prefiltered
                              vp, vp, vp, vp,        # out_color out_feature out_depth radii
                              i, vp]                 # debug stream
    L.mgs_forward_begin.restype = C.c_int
    L.mgs_forward_begin.argtypes = [ALLOC_FN, vp, ALLOC_FN, vp, i, i, i, i, i,
                                    vp, vp, vp, vp, vp, f, vp, vp, vp, vp, vp, f, f, vp, vp, i, vp]
    L.mgs_forward_finish.restype = C.c_int
    L.mgs_forward_finish.argtypes = [ALLOC_FN, vp, vp, vp, vp, i, i, i, i, vp, vp, vp, i, vp, vp, vp, i, i, vp]
    L.mgs_backward.restype = C.c_int
    L.mgs_backward.argtypes = [i, i, i, i, i,        # P D M F R
                               vp, i, i,             # background width height
                               vp, vp, vp, vp,       # means3D shs colors feature
                               vp, f, vp, vp,        # scales scale_modifier rotations cov3D_precomp
                               vp, vp, vp, f, f,     # viewmatrix projmatrix campos tan_fovx tan_fovy
                               vp, vp, vp, vp,       # radii geom binning image
                               vp, vp, vp,           # dL_dpix dL_dpix_F dL_dpix_depth
                               vp, vp, vp, vp, vp,   # dL_dmean2D dL_dconic dL_dopacity dL_dcolor dL_dfeature
                               vp, vp, vp, vp, vp,   # dL_dmean3D dL_dcov3D dL_dsh dL_dscale dL_drot
                               vp, i, i, vp]         # scratch accumulate debug stream
    L.mgs_forward_views.restype = C.c_int
    L.mgs_forward_views.argtypes = [i, C.POINTER(View), i, i, i, i,     # V views P D M F
                                    vp, vp, vp, vp,                     # means3D shs colors feature
                                    vp, vp, f, vp, vp,                  # opacities scales scale_modifier rotations cov3D_precomp
                                    i, i, vp]                           # prefiltered debug join_stream
    L.mgs_backward_views.restype = C.c_int
    L.mgs_backward_views.argtypes = [i, C.POINTER(View), i, i, i, i,    # V views P D M F
                                     vp, vp, vp, vp,                    # means3D shs colors feature
                                     vp, f, vp, vp,                     # scales scale_modifier rotations cov3D_precomp
                                     vp, vp, vp, vp, vp, vp, vp, vp,    # dL: mean3D opacity color feature cov3D sh scale rot
                                     i, i, i, i, vp]                    # shared_mean2D accumulate stages debug join_stream
    L.mgs_loss_heads.restype = C.c_int
    L.mgs_loss_heads.argtypes = [i, i, i, vp, vp, vp, vp, vp, vp, vp, vp]
    L.mgs_mark_visible.restype = C.c_int
    L.mgs_mark_visible.argtypes = [i, vp, vp, vp, vp, vp]
    L.mgs_state_array.restype = C.c_int
    L.mgs_state_array.argtypes = [C.c_char_p, C.c_char_p, vp, i, i, C.POINTER(C.c_void_p)]
    L.mgs_profile_enable.argtypes = [i]
    L.mgs_profile_num_stages.restype = C.c_int
    L.mgs_profile_stage_name.restype = C.c_char_p
    L.mgs_profile_stage_name.argtypes = [i]
    L.mgs_profile_read.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.mgs_activate.restype = C.c_int
    L.mgs_activate.argtypes = [i, i, vp, vp, vp, vp, vp, vp, vp, vp,   # P F means d_means rot d_rot scales d_scales opac feature
                               i, f, i, i, i,                           # scale_mode scale_max opacity_mode rot_normalize feature_normalize
                               vp, vp, vp, vp, vp, vp]                  # out_means out_rot out_scales out_opac out_feature stream
    L.mgs_activate_backward.restype = C.c_int
    L.mgs_activate_backward.argtypes = [i, i, vp, vp, vp, vp, vp, vp, vp, vp, i, f, i, i, i,
                                        vp, vp, vp, vp, vp,             # g_means g_rot g_scales g_opac g_feature
                                        vp, vp, vp, vp, vp, vp, vp, vp, # dL: means d_means rot d_rot scales d_scales opac feature
                                        vp]
    if L.mgs_abi_version() != 200:
        raise ImportError("manigaussian_b200: ABI version mismatch")
    _lib = L
    return L


EXPORTED_SYMBOLS = (
    "mgs_abi_version", "mgs_last_error", "mgs_geometry_state_bytes", "mgs_image_state_bytes",
    "mgs_binning_state_bytes", "mgs_backward_scratch_bytes", "mgs_forward", "mgs_forward_begin", "mgs_forward_finish",
    "mgs_backward", "mgs_forward_views", "mgs_backward_views", "mgs_loss_heads",
    "mgs_activate", "mgs_activate_backward",
    "mgs_mark_visible", "mgs_state_array", "mgs_profile_enable", "mgs_profile_num_stages",
    "mgs_profile_stage_name", "mgs_profile_read",
)


def profile_enable(on=True):
    lib().mgs_profile_enable(int(bool(on)))


def profile_read():
    """{stage: (total_ms, launches)} since the last read."""
    L = lib()
    n = L.mgs_profile_num_stages()
    ms, cnt = (C.c_float * n)(), (C.c_int * n)()
    L.mgs_profile_read(ms, cnt)
    return {L.mgs_profile_stage_name(k).decode(): (float(ms[k]), int(cnt[k])) for k in range(n)}


def last_error():
    return lib().mgs_last_error().decode()


def check(rc, what):
    if rc < 0:
        raise RuntimeError(f"{what} failed ({rc}): {last_error()}")
    return rc
