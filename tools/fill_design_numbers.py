"""Regenerate the results table of DESIGN.md section 5 (between the BENCH_TABLE markers) from the bench JSON lines
committed under profiles/."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROWS = (("c3: 500k, 4×256², RGB+32 feat (headline)", "r1_bench_ours_c3.json", "r1_bench_ref_c3.json"),
        ("c4: c3 + deformation offsets, 4 current + 4 next-frame views, grads to raw maps and Δμ/Δr/Δs", "r1_bench_ours_c4.json", "r1_bench_ref_c4.json"),
        ("c2: 200k, 256², RGB+depth", "r1_bench_ours_c2.json", "r1_bench_ref_c2.json"),
        ("c1: 50k, 128², RGB", "r1_bench_ours_c1.json", "r1_bench_ref_c1.json"),
        ("ManiGaussian's real call: 16k, 128², F=3", "r1_bench_ours_mg.json", "r1_bench_ref_mg.json"),
        ("c3 on 2 GPUs (view-parallel + all-reduce, 4 views per GPU)", "r1_scale_ours_2.json", "r1_scale_ref_2.json"),
        ("c3 on 4 GPUs", "r1_scale_ours_4.json", "r1_scale_ref_4.json"))


def load(name):
    try:
        return json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
    except Exception:
        return None


def main():
    lines = ["| workload | ours Gaussians/s (ms/step) | reference (ms/step) | ratio | e2e ours / reference |", "|---|---|---|---|---|"]
    for label, fo, fr in ROWS:
        o, r = load(fo), load(fr)
        if not o or not r:
            continue
        eo, er = o.get("e2e", {}).get("value"), r.get("e2e", {}).get("value")
        # later e2e-only runs (packed H2D copy, single-view fast path) supersede the e2e column where they exist
        o2, r2 = load(fo.replace("r1_bench_", "r1_e2e_")), load(fr.replace("r1_bench_", "r1_e2e_"))
        if fo.startswith("r1_bench_") and o2 and r2:
            eo, er = o2["e2e"]["value"], r2["e2e"]["value"]
        e2e = f"{eo:.3g} / {er:.3g} = {eo / er:.1f}×" if eo and er else "—"
        lines.append(f"| {label} | {o['value']:.3g} ({o['ms_per_step']:.2f}) | {r['value']:.3g} ({r['ms_per_step']:.2f}) | "
                     f"{o['value'] / r['value']:.1f}× | {e2e} |")
    p = os.path.join(ROOT, "DESIGN.md")
    s = open(p).read()
    a, b = s.index("<!-- BENCH_TABLE_BEGIN -->"), s.index("<!-- BENCH_TABLE_END -->")
    s = s[:a] + "<!-- BENCH_TABLE_BEGIN -->\n" + "\n".join(lines) + "\n" + s[b:]
    open(p, "w").write(s)
    print("\n".join(lines))


if __name__ == "__main__":
    sys.exit(main())
