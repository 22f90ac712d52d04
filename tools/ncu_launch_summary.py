"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    if r[iu] == "us":
        v *= 1e3
    elif r[iu] == "ms":
        v *= 1e6
    name = r[ik].split("(")[0][-70:]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"# {len(rows) - 1} launches, {tot / 1e6:.3f} ms total (cold-cache, serialised: compare SHARES, not absolutes)")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * t / tot:6.2f}%  {t / 1e3 / n:10.1f} us/launch  x{n:<4d} {name}")
