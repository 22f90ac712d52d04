"""The numpy oracle of the loss heads (oracle/loss_oracle.py) against fixtures produced by the REFERENCE's own l2_loss /
cosine_loss under torch autograd (tests/golden/make_loss_golden.py)."""
import os

import numpy as np
import pytest

from oracle import loss_oracle as lo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("name", ["f32", "f3"])
def test_loss_oracle_against_reference_golden(name):
    z = np.load(os.path.join(GOLD, "loss_heads.npz"))
    l, g = lo.l2_head(z[f"{name}_render"], z[f"{name}_gt"])
    assert abs(l - float(z[f"{name}_loss_rgb"])) <= 1e-6 * abs(float(z[f"{name}_loss_rgb"]))
    assert _rel(g, z[f"{name}_d_render"]) < 1e-6
    l, g = lo.cosine_head(z[f"{name}_embed"], z[f"{name}_gt_embed"])
    assert abs(l - float(z[f"{name}_loss_embed"])) <= 1e-6
    assert _rel(g, z[f"{name}_d_embed"]) < 1e-5
