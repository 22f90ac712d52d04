"""oracle/activate_oracle.py against fixtures made with the real PyTorch operators of the reference
(tests/golden/make_activate_golden.py; models_embed.py:245-252,297-304, gaussian_renderer/__init__.py:66-68)."""
import os

import numpy as np
import pytest

from oracle import activate_oracle as ao

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 2e-6


def rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / d if d > 0 else np.linalg.norm(a - b)


@pytest.mark.parametrize("name", ["f32", "f3_edge", "f0"])
def test_oracle_matches_torch_golden(name):
    z = np.load(os.path.join(GOLD, f"activate_{name}.npz"))
    F = int(z["F"])
    feat = z["feature_maps"] if F else None
    fw = ao.forward(z["xyz"], z["rot_maps"], z["scale_maps"], z["opacity_maps"], feat, d_means=z["xyz_maps"])
    for k in ("means", "rot", "scales", "opac") + (("feature",) if F else ()):
        assert rel(fw[k], z["out_" + k]) < TOL, k
    g = {k: z["cot_" + k] for k in ("means", "rot", "scales", "opac")}
    g["feature"] = z["cot_feature"] if F else None
    bw = ao.backward(z["xyz"], z["rot_maps"], z["scale_maps"], z["opacity_maps"], feat, z["xyz_maps"], None, None, g)
    assert rel(bw["means"], z["grad_xyz_maps"]) < TOL
    assert rel(bw["rot"], z["grad_rot_maps"]) < TOL
    assert rel(bw["scales"], z["grad_scale_maps"]) < TOL
    assert rel(bw["opac"], z["grad_opacity_maps"]) < TOL
    if F:
        assert rel(bw["feature"], z["grad_feature_maps"]) < TOL
    if "next_xyz" in z.files:
        # next frame: activated values detached, offsets added, identity scale/opacity (models_embed.py:297-304)
        nf = ao.forward(fw["means"], fw["rot"], fw["scales"], fw["opac"], feat, d_means=z["next_xyz"], d_rot=z["next_rot"],
                        scale_mode=0, opacity_mode=0)
        assert rel(nf["means"], z["next_out_means"]) < TOL
        assert rel(nf["rot"], z["next_out_rot"]) < TOL
        assert np.array_equal(nf["scales"], fw["scales"]) and np.array_equal(nf["opac"], fw["opac"])
        if F:
            assert rel(nf["feature"], z["next_out_feature"]) < TOL
        zero = lambda a: np.zeros_like(a)
        gn = dict(means=z["next_cot_means"], rot=z["next_cot_rot"], scales=zero(fw["scales"]), opac=zero(fw["opac"]), feature=None)
        nb = ao.backward(fw["means"], fw["rot"], fw["scales"], fw["opac"], None, z["next_xyz"], z["next_rot"], None, gn,
                         scale_mode=0, opacity_mode=0)
        assert rel(nb["means"], z["grad_next_xyz"]) < TOL
        assert rel(nb["rot"], z["grad_next_rot"]) < TOL


def test_edge_rows_are_finite_where_torch_is():
    z = np.load(os.path.join(GOLD, "activate_f3_edge.npz"))
    fw = ao.forward(z["xyz"], z["rot_maps"], z["scale_maps"], z["opacity_maps"], z["feature_maps"], d_means=z["xyz_maps"])
    assert np.all(fw["rot"][0] == 0) and np.all(fw["feature"][1] == 0)   # zero-norm rows stay zero, no NaN
    assert fw["scales"][3].max() == np.float32(0.05)                     # clamp
    assert np.all(z["grad_scale_maps"][3] == 0)                          # no gradient through the clamped scale
