"""step-by-step probe of the library under ncu (diagnostic)"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
def say(*a):
    print("PROBE", *a, flush=True)
import torch
say("torch ok", float(torch.ones(4, device="cuda").sum()))
L = ctypes.CDLL("manigaussian_b200/lib/libmgs_rasterizer.so")
say("dlopen ok", L.mgs_abi_version())
P = 1000
m = torch.rand(P, 3, device="cuda"); vm = torch.eye(4, device="cuda").reshape(-1).contiguous(); pr = torch.zeros(P, dtype=torch.uint8, device="cuda")
L.mgs_mark_visible.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 5
say("mark_visible rc", L.mgs_mark_visible(P, m.data_ptr(), vm.data_ptr(), vm.data_ptr(), pr.data_ptr(), None)); torch.cuda.synchronize(); say("sync ok")
sys.path.insert(0, "tests")
import util
say("util imported")
inp = util.make_inputs(P=300, W=32, H=32, F=3, seed=1)
say("inputs made")
from manigaussian_b200 import _binding as b
Lb = b.lib()
say("binding lib ok")
say("geom bytes", Lb.mgs_geometry_state_bytes(300))
say("img bytes", Lb.mgs_image_state_bytes(32, 32))
say("bin bytes", Lb.mgs_binning_state_bytes(1000))
fw, bw = util.run_ours(inp, debug=True)
say("run_ours ok", fw["num_rendered"])
