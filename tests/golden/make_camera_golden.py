"""Generates tests/golden/cameras.npz by running the REFERENCE's own camera code in the build container:
NeuralRenderer.get_novel_calib (agents/manigaussian_bc/neural_rendering.py:205-248) is lifted out of the source file with
`ast` at run time (its module cannot be imported here: rlbench, visdom, ... are absent) and executed against the real
helpers of agents/manigaussian_bc/graphics_utils.py.  Nothing of the reference is copied into this repository.
Run:  python tests/golden/make_camera_golden.py"""
import ast
import importlib.util
import os
import types

import numpy as np
import torch

REF = "/root/reference/agents/manigaussian_bc"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_calib():
    spec = importlib.util.spec_from_file_location("ref_graphics_utils", os.path.join(REF, "graphics_utils.py"))
    gu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gu)
    tree = ast.parse(open(os.path.join(REF, "neural_rendering.py")).read())
    fn = next(n for c in tree.body if isinstance(c, ast.ClassDef) and c.name == "NeuralRenderer"
              for n in c.body if isinstance(n, ast.FunctionDef) and n.name == "get_novel_calib")
    ns = {"np": np, "torch": torch, "getWorld2View2": gu.getWorld2View2, "getProjectionMatrix": gu.getProjectionMatrix,
          "focal2fov": gu.focal2fov}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "neural_rendering.py", "exec"), ns)
    return ns["get_novel_calib"]


def look_at(eye, target, up=(0, 0, 1)):
    """camera-to-world pose, OpenCV axes (x right, y down, z forward)"""
    eye, target, up = (np.asarray(a, np.float64) for a in (eye, target, up))
    z = target - eye
    z /= np.linalg.norm(z)
    x = np.cross(z, up)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x, y, z, eye
    return m


if __name__ == "__main__":
    calib = load_reference_calib()
    rng = np.random.default_rng(5)
    out = {}
    for name, (W, H, trans, scale) in {"square128": (128, 128, [0, 0, 0], 1.0), "wide": (160, 96, [0.1, -0.2, 0.05], 1.5)}.items():
        B = 5
        intr = np.zeros((B, 3, 3), np.float32)
        extr = np.zeros((B, 4, 4), np.float32)
        for b in range(B):
            f = W / (2 * np.tan(np.deg2rad(20 + 5 * b)))
            intr[b] = [[f, 0, W / 2 + rng.normal(0, 2)], [0, f * (1 + 0.02 * b), H / 2 + rng.normal(0, 2)], [0, 0, 1]]
            az = 2 * np.pi * b / B
            extr[b] = look_at([0.2 + 1.6 * np.cos(az), 1.6 * np.sin(az), 1.5], [0.2, 0, 1.1])
        self = types.SimpleNamespace(W=W, H=H, znear=0.1, zfar=4.0, trans=trans, scale=scale)
        # K is handed over as float64 holding the float32 values: under the reference's NumPy 1.x, `znear / K[0, 0]` promotes
        # a float32 scalar to float64; under this container's NumPy 2 it would stay float32 and getProjectionMatrix's
        # `P[0, 0] = ...` then fails (np.float32 is not assignable to a torch element).  Same numbers as the reference's env.
        nv = calib(self, {"intr": torch.from_numpy(intr.astype(np.float64)), "extr": torch.from_numpy(extr)})
        out.update({f"{name}_intr": intr, f"{name}_extr": extr, f"{name}_WH": np.array([W, H]),
                    f"{name}_trans": np.array(trans, np.float64), f"{name}_scale": np.array(scale)})
        for k, v in nv.items():
            out[f"{name}_{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "cameras.npz"), **out)
    print("wrote cameras.npz:", sorted(out)[:8], "...")
