"""Synthetic cameras and Gaussian clouds for tests and bench.py.

Cameras follow ManiGaussian's construction (agents/manigaussian_bc/neural_rendering.py:205-248 with
graphics_utils.py:17-52: getWorld2View2, getProjectionMatrix(K), focal2fov): matrices are stored transposed
(row-vector convention), full_proj = world_view @ projection, camera centre = inverse(world_view)[3,:3].
Everything is generated with numpy on the CPU from a seed, so the CPU oracle and the GPU see identical bits.
"""
import math

import numpy as np

SCENE_BOUNDS = np.array([-0.3, -0.5, 0.6, 0.7, 0.5, 1.6], np.float32)  # conf/config.yaml:21 of the reference
SCENE_CENTER = np.array([0.2, 0.0, 1.1], np.float32)
ZNEAR, ZFAR = 0.1, 4.0  # conf/method/ManiGaussian_BC.yaml:107-108


def _projection_matrix(znear, zfar, K, h, w):
    """OpenGL-style projection from intrinsics (restates graphics_utils.py:31-48)."""
    near_fx, near_fy = znear / K[0, 0], znear / K[1, 1]
    left, right = -(w - K[0, 2]) * near_fx, K[0, 2] * near_fx
    bottom, top = (K[1, 2] - h) * near_fy, K[1, 2] * near_fy
    P = np.zeros((4, 4), np.float32)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def make_camera(W, H, view=0, num_views=1, fov_deg=40.0, radius=1.6, height=0.4, center=SCENE_CENTER):
    """Camera `view` of `num_views` on a circle around the scene centre, looking at it (x right, y down, z forward)."""
    az = 2.0 * math.pi * view / max(num_views, 1)
    eye = np.array([center[0] + radius * math.cos(az), center[1] + radius * math.sin(az), center[2] + height], np.float64)
    fwd = np.asarray(center, np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, eye
    w2c = np.linalg.inv(c2w)
    fx = fy = W / (2.0 * math.tan(math.radians(fov_deg) / 2.0))
    K = np.array([[fx, 0, W / 2.0], [0, fy, H / 2.0], [0, 0, 1]], np.float64)
    world_view = np.float32(w2c).T  # transposed storage
    proj = _projection_matrix(ZNEAR, ZFAR, K, H, W).T
    full = (world_view @ proj).astype(np.float32)
    campos = np.linalg.inv(world_view.astype(np.float64))[3, :3].astype(np.float32)
    fovx, fovy = 2 * math.atan(W / (2 * fx)), 2 * math.atan(H / (2 * fy))
    return dict(W=W, H=H, viewmatrix=np.ascontiguousarray(world_view), projmatrix=np.ascontiguousarray(full),
                campos=campos, tanfovx=math.tan(fovx * 0.5), tanfovy=math.tan(fovy * 0.5))


def make_gaussians(P, F=0, sh_degree=1, seed=0, scale0=None, precomp_colors=False):
    """Seeded Gaussian cloud in ManiGaussian's scene box with its activations' ranges (models_embed.py:245-252)."""
    rng = np.random.default_rng(seed)
    lo, hi = SCENE_BOUNDS[:3], SCENE_BOUNDS[3:]
    means = (lo + (hi - lo) * rng.random((P, 3))).astype(np.float32)
    s0 = scale0 if scale0 is not None else 0.02 * math.sqrt(16384.0 / max(P, 1))
    scales = np.minimum(np.exp(rng.normal(math.log(s0), 0.5, (P, 3))), 0.05).astype(np.float32)
    q = rng.normal(0, 1, (P, 4))
    rotations = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    opacities = (1.0 / (1.0 + np.exp(-rng.normal(-2.0, 1.0, (P, 1))))).astype(np.float32)
    M = (sh_degree + 1) ** 2
    shs = rng.normal(0, 0.1, (P, M, 3))
    shs[:, 0, :] = rng.normal(0, 1.0, (P, 3))
    out = dict(means3D=means, scales=scales, rotations=rotations, opacities=opacities, sh_degree=sh_degree,
               shs=shs.astype(np.float32), colors_precomp=None, feature=None)
    if precomp_colors:
        out["colors_precomp"] = rng.random((P, 3)).astype(np.float32)
        out["shs"] = None
    if F > 0:
        f = rng.normal(0, 1, (P, F))
        out["feature"] = (f / (np.linalg.norm(f, axis=1, keepdims=True) + 1e-12)).astype(np.float32)
    return out


def make_cotangents(W, H, F=0, seed=0, depth=False):
    rng = np.random.default_rng(seed + 7919)
    out = dict(dL_dcolor=rng.normal(0, 1, (3, H, W)).astype(np.float32),
               dL_dfeature=rng.normal(0, 1, (F, H, W)).astype(np.float32) if F > 0 else None)
    out["dL_ddepth"] = rng.normal(0, 1, (H, W)).astype(np.float32) if depth else None
    return out


def alg_bytes_per_view(P, R, N, M, F, depth=False, precomp_colors=False):
    """Algorithmic (compulsory) HBM bytes of one fwd+bwd view, SURVEY.md 8(d):
    P*(212 + 36*M + 12*F') + 20*R + N*(8*(3+F') + 16), F' = F + depth."""
    Fp = F + (1 if depth else 0)
    m_term = 36 if precomp_colors else 36 * M
    return P * (212 + m_term + 12 * Fp) + 20 * R + N * (8 * (3 + Fp) + 16)
