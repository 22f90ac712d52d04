"""Who keeps the per-step tensors alive?  python tools/diag_leak.py [workload] [heads]"""
import collections
import gc
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "mg"])
heads = len(sys.argv) > 2 and sys.argv[2] == "heads"
g, cams, cts = bench.host_inputs(wl, 0, 1)
Gh, Ch, Th = bench.to_device(g, cams, cts, torch, pinned=True)
step = bench.make_e2e("ours", wl, torch, None, heads=heads)
st = {}
for _ in range(8):
    step(Gh, Ch, Th, st)
step.flush()
torch.cuda.synchronize()
gc.collect()
objs = [o for o in gc.get_objects() if isinstance(o, torch.Tensor) and o.is_cuda]
by = collections.Counter((tuple(o.shape), str(o.dtype).replace("torch.", ""), o.requires_grad, o.grad_fn is not None) for o in objs)
print("live CUDA tensor objects (python-visible):", len(objs))
for k, n in by.most_common(25):
    print("  ", n, k)
# python-level owners of the tensors that exist many times
def owners(t, depth=0, seen=None):
    seen = seen or set()
    out = []
    for r in gc.get_referrers(t):
        if id(r) in seen or r is objs or r is by:
            continue
        seen.add(id(r))
        name = type(r).__name__
        if name in ("frame",):
            continue
        desc = name
        if isinstance(r, dict):
            desc += " keys=" + ",".join(str(k)[:20] for k in list(r.keys())[:8])
        elif isinstance(r, (tuple, list)):
            desc += f" len={len(r)} of " + ",".join(type(x).__name__ for x in list(r)[:8])
        out.append(desc)
        if depth < 2 and isinstance(r, (tuple, list, dict)) and name not in ("module",):
            out += ["  <- " + x for x in owners(r, depth + 1, seen)]
    return out
common = [k for k, n in by.items() if n >= 6]
for k in common[:6]:
    t = next(o for o in objs if (tuple(o.shape), str(o.dtype).replace("torch.", ""), o.requires_grad, o.grad_fn is not None) == k)
    print("owners of one", k, ":")
    for line in owners(t)[:14]:
        print("     ", line[:220])
bufs = [o for o in objs if tuple(o.shape) == (524416,)]
def where(t):
    for b in bufs:
        d = t.data_ptr() - b.data_ptr()
        if 0 <= d < b.numel() * 4:
            return f"inside upload buffer @+{d // 4}"
    return "own storage"
seen = set()
for o in objs:
    k = (tuple(o.shape), o.requires_grad, o.grad_fn is not None)
    if k in seen:
        continue
    seen.add(k)
    gf = o.grad_fn
    chain = []
    n = gf
    for _ in range(6):
        if n is None:
            break
        chain.append(type(n).__name__)
        nxt = [f for f, _ in n.next_functions if f is not None]
        n = nxt[0] if nxt else None
    print(k, "leaf" if o.is_leaf else "non-leaf", "base", None if o._base is None else tuple(o._base.shape), where(o), "grad" if o.grad is not None else "", "chain", chain)
