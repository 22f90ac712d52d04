"""world_size-2 gloo test (CPU) of the view-parallel host logic: view sharding and the single packed all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util  # noqa: F401
from manigaussian_b200.parallel import FIELDS, PackedGradients, shard_views


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, P, F, M, total_views, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pk = PackedGradients(P, F, M, "cpu")
    mine = shard_views(total_views, rank, world)
    for v in mine:  # a view's gradient is (v+1) * field_index everywhere: easy to sum in closed form
        pk.accumulate({k: torch.full((P, pk.widths[k]), float((v + 1) * (i + 1))) for i, k in enumerate(FIELDS) if pk.widths[k]})
    pk.all_reduce()
    tot = sum(v + 1 for v in range(total_views))
    ok = all(torch.all(pk.views[k] == tot * (i + 1)).item() for i, k in enumerate(FIELDS) if k in pk.views)
    out.put((rank, mine, ok, pk.bytes_per_gaussian))
    dist.destroy_process_group()


def test_view_sharding_covers_every_view_once():
    for world in (1, 2, 4, 8):
        for total in (1, 4, 8, 13):
            owned = sorted(v for r in range(world) for v in shard_views(total, r, world))
            assert owned == list(range(total))


def test_packed_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 257, 32, 4, 5, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, _, ok, _ in res)
    # bytes per Gaussian at M=4, F=32: SURVEY.md 8(e)'s 232 minus the 12 bytes of the per-view screen-space gradients,
    # which stay on the rank that rendered the view
    assert res[0][3] == 220
    assert sorted(v for _, mine, _, _ in res for v in mine) == list(range(5))
