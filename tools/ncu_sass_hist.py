"""Opcode-level summary of `ncu --page source --csv` (SASS view): executed warp-instructions and stall samples per opcode,
plus the hottest address ranges.  usage: ncu -i X.ncu-rep --page source --csv > f.csv; python tools/ncu_sass_hist.py f.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
kern, hdr, per = None, None, collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        kern = r[1][:60]
        per[kern] = []
        continue
    if r[0] == "Address":
        hdr = r
        continue
    if hdr is None or kern is None:
        continue
    d = dict(zip(hdr, r))
    try:
        per[kern].append((d["Source"].strip(), float(d["Instructions Executed"]), float(d["# Samples"]), d))
    except Exception:
        pass
for k, items in per.items():
    ti, ts = sum(i[1] for i in items), sum(i[2] for i in items)
    print("=====", k, "inst %.4g samples %d sass lines %d" % (ti, ts, len(items)))
    ops = collections.defaultdict(lambda: [0.0, 0.0])
    for src, ie, s, _ in items:
        op = src.split()[0] if not src.startswith("@") else src.split()[1]
        op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "HMMA", "MUFU", "RED", "LDGSTS")) and "." in op else "")
        ops[op][0] += ie
        ops[op][1] += s
    for op, (ie, s) in sorted(ops.items(), key=lambda x: -x[1][0])[:28]:
        print("  %-14s inst %5.1f%%  samples %5.1f%%" % (op, 100 * ie / ti, 100 * s / max(ts, 1)))
    # hottest 64-instruction windows by samples
    W = 64
    wins = []
    for i in range(0, len(items), W):
        seg = items[i:i + W]
        wins.append((sum(x[2] for x in seg), sum(x[1] for x in seg), i))
    print("  hottest windows (sass index: samples%, inst%):", " ".join("%d:%.1f/%.1f" % (i, 100 * s / max(ts, 1), 100 * ie / ti) for s, ie, i in sorted(wins, reverse=True)[:12]))
    st = collections.defaultdict(float)
    for _, _, _, d in items:
        for kk, v in d.items():
            if kk.startswith("stall_") and "Not Issued" not in kk:
                try:
                    st[kk] += float(v)
                except Exception:
                    pass
    tot = sum(st.values())
    print("  stall samples:", " ".join("%s=%.1f%%" % (a[6:], 100 * b / max(tot, 1)) for a, b in sorted(st.items(), key=lambda x: -x[1])[:10]))
