"""Per-tensor relative-L2 report: ours vs oracle, ours vs compiled reference, reference vs oracle (GPU box)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from test_parity_gpu import CASES  # noqa: E402

cases = dict(CASES)
cases["c3_like_100k"] = dict(P=100000, W=256, H=256, F=32, seed=9)
rep = {}
for name, kw in cases.items():
    inp = util.make_inputs(**kw)
    o_fw, o_bw = util.run_ours(inp)
    c_fw, c_bw = util.run_oracle(inp)
    r_fw, r_bw = util.run_reference(inp)
    F = inp["F"]
    row = {}
    for tag, (afw, abw), (bfw, bbw) in (("ours_vs_oracle", (o_fw, o_bw), (c_fw, c_bw)), ("ours_vs_ref", (o_fw, o_bw), (r_fw, r_bw)),
                                          ("ref_vs_oracle", (r_fw, r_bw), (c_fw, c_bw))):
        if afw is None or bfw is None:
            continue
        d = {"out_color": util.rel_l2(afw["out_color"], bfw["out_color"]), "final_T": util.rel_l2(afw["final_T"], bfw["final_T"]),
             "n_contrib_mismatch": int((afw["n_contrib"] != bfw["n_contrib"]).sum()),
             "radii_mismatch": int((afw["radii"] != bfw["radii"]).sum()),
             "keys_equal": bool(afw["num_rendered"] == bfw["num_rendered"] and np.array_equal(afw["point_list_keys"], bfw["point_list_keys"])),
             "point_list_equal": bool(afw["num_rendered"] == bfw["num_rendered"] and np.array_equal(afw["point_list"], bfw["point_list"]))}
        if F:
            d["out_feature"] = util.rel_l2(afw["out_feature"], bfw["out_feature"])
        for k in abw:
            if k in bbw and not (k == "dL_dfeature" and not F):
                d[k] = util.rel_l2(abw[k], bbw[k])
        row[tag] = d
    rep[name] = row
    print(name, json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w"), indent=1)
