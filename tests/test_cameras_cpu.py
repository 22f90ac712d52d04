"""manigaussian_b200.cameras against fixtures produced by the reference's own get_novel_calib / graphics_utils
(tests/golden/make_camera_golden.py).  CPU only: the builder is host code; the device path is one packed copy."""
import math
import os

import numpy as np
import pytest
import torch

from manigaussian_b200 import cameras

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "cameras.npz"))
TOL = 2e-6


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("name", ["square128", "wide"])
def test_matches_reference_calibration(name):
    W, H = (int(x) for x in GOLD[f"{name}_WH"])
    cb = cameras.build_cameras(torch.from_numpy(GOLD[f"{name}_intr"]), torch.from_numpy(GOLD[f"{name}_extr"]), W, H,
                               znear=0.1, zfar=4.0, trans=GOLD[f"{name}_trans"], scale=float(GOLD[f"{name}_scale"]))
    B = len(cb)
    assert B == GOLD[f"{name}_intr"].shape[0] and cb.width == W and cb.height == H
    np.testing.assert_allclose(cb.FovX, GOLD[f"{name}_FovX"], rtol=1e-6)
    np.testing.assert_allclose(cb.FovY, GOLD[f"{name}_FovY"], rtol=1e-6)
    assert np.array_equal(GOLD[f"{name}_width"], [W] * B) and np.array_equal(GOLD[f"{name}_height"], [H] * B)
    for b in range(B):
        assert rel(cb.world_view_transform[b].numpy(), GOLD[f"{name}_world_view_transform"][b]) < TOL
        assert rel(cb.full_proj_transform[b].numpy(), GOLD[f"{name}_full_proj_transform"][b]) < TOL
        assert rel(cb.camera_center[b].numpy(), GOLD[f"{name}_camera_center"][b]) < TOL
        assert cb.tanfovx[b] == pytest.approx(math.tan(0.5 * float(GOLD[f"{name}_FovX"][b])), rel=1e-6)
    nv = cb.as_novel_view()
    assert set(nv) >= {"FovX", "FovY", "width", "height", "world_view_transform", "full_proj_transform", "camera_center"}
    s = cb.settings(1, bg=torch.zeros(3), sh_degree=1, include_feature=True)
    assert s.image_width == W and s.image_height == H and s.viewmatrix.shape == (4, 4) and s.viewmatrix.is_contiguous()


def test_single_camera_and_bad_shapes():
    cb = cameras.build_cameras(GOLD["square128_intr"][0], GOLD["square128_extr"][0], 128, 128)
    assert len(cb) == 1
    with pytest.raises(ValueError):
        cameras.build_cameras(np.zeros((2, 3, 3)), np.zeros((3, 4, 4)), 8, 8)
