"""Instruction evidence per kernel from the built library: python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "manigaussian_b200", "lib", "libmgs_rasterizer.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
blocks = re.split(r"\n\s*Function : \S+\n", "\n" + sass)[1:]
pats = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("LDGSTS", r"\bLDGSTS"), ("HMMA", r"\bHMMA\.1688\.F32\.TF32"), ("REDG", r"\bREDG\.E\.ADD\.F32\b(?!x)"),
        ("REDGx4", r"\bREDG\.E\.ADD\.F32x4"), ("LDS128", r"\bLDS\.128"), ("STS128", r"\bSTS\.128")]
print("# cuobjdump -sass manigaussian_b200/lib/libmgs_rasterizer.so (sm_100a cubins): instruction evidence per hand-written kernel")
print("# UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier ops, LDGSTS = cp.async, HMMA.1688.F32.TF32 = mma.sync m16n8k8 TF32,")
print("# REDG.E.ADD.F32 = red.global.add.f32, REDGx4 = red.global.add.v4.f32 (REDG.E.ADD.F32x4), LDS.128/STS.128 = 128-bit shared-memory accesses")
print("%-58s %6s " % ("kernel", "instrs") + " ".join("%6s" % p[0] for p in pats))
for name, body in zip(names, blocks):
    short = re.sub(r"\(mgs::.*|\(int,.*|\(const.*|\(unsigned.*", "", name.replace("mgs::", "").replace("(bool)", "").replace("(int)", ""))[:58]
    if "cub::" in short and "DeviceScanKernel" not in short:
        continue
    n = len(re.findall(r"^\s+/\*[0-9a-f]{4,6}\*/", body, re.M))
    print("%-58s %6d " % (short, n) + " ".join("%6d" % len(re.findall(p[1], body)) for p in pats))
