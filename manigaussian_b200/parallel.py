"""View-parallel multi-GPU plumbing.

The rasterizer shards by VIEW: Gaussians are replicated on every rank, each rank renders its own views, and the only
exchange is one SUM all-reduce of the per-Gaussian gradients.  All gradient tensors the optimiser consumes live in ONE
flat fp32 buffer (`PackedGradients.flat`) so that the exchange is a single collective on a single message:
  [means3D 3 | scales 3 | rotations 4 | opacity 1 | SH 3M | features F] x P   (220 B/Gaussian at M=4, F=32).
The screen-space mean gradients are per VIEW (the reference's `viewspace_points.grad` of each render) and stay on the rank
that rendered the view; they are not part of the message.
The reference has no counterpart: it renders one view per call on one GPU and lets DDP all-reduce MLP parameters
(train.py:92-105); Gaussians never cross GPUs there.
"""
import torch

# the feature field comes first: it is final after the blend stage of the backward, so its part of the all-reduce is issued
# before the per-Gaussian kernel that produces the other fields and overlaps it (PackedGradients.all_reduce_begin/finish)
FIELDS = ("dL_dfeature", "dL_dmeans3D", "dL_dscales", "dL_drotations", "dL_dopacity", "dL_dsh", "dL_dcolors", "dL_dmeans2D")


def shard_views(total_views, rank, world):
    """Round-robin view ownership: view v belongs to rank v % world."""
    return [v for v in range(total_views) if v % world == rank]


class PackedGradients:
    def __init__(self, P, F, M, device, colors=False, means2D=False, zero=True, registered=False):
        """colors=True adds a [P,3] field for precomputed-colour gradients (used instead of SH when M == 0); means2D=True a
        [P,3] field for the screen-space gradients summed over views (single-view callers of the raw accumulate mode).
        zero=False leaves the buffer uninitialised (rasterize_views_backward_raw overwrites every row).  registered=True
        allocates the buffer from NCCL's allocator and registers it with the communicator (user-buffer registration: NVLS /
        zero-copy collectives straight on this memory); silently falls back to a plain allocation where that is unavailable --
        meant for long-lived buffers, registration costs far more than one all-reduce."""
        self.widths = dict(dL_dmeans3D=3, dL_dscales=3, dL_drotations=4, dL_dopacity=1, dL_dsh=3 * M, dL_dfeature=F,
                           dL_dcolors=3 if colors else 0, dL_dmeans2D=3 if means2D else 0)
        self.P = P
        # every field starts on a 16-byte boundary so that 128-bit reductions can target it directly
        offs, off = {}, 0
        for k in FIELDS:
            offs[k] = off
            off += (P * self.widths[k] + 3) // 4 * 4
        self.registered = False
        self.flat = None
        if registered:
            self.flat = self._alloc_registered(off, device)
        if self.flat is None:
            self.flat = torch.empty(off, dtype=torch.float32, device=device)
        if zero:
            self.flat.zero_()
        self.split = offs["dL_dmeans3D"] if self.widths["dL_dfeature"] else 0  # [0, split) = feature field, [split, end) = the rest
        self._work = []
        self.views = {k: self.flat[offs[k]:offs[k] + P * self.widths[k]].view(P, self.widths[k]) for k in FIELDS if self.widths[k]}

    @property
    def bytes_per_gaussian(self):
        return 4 * sum(self.widths.values())

    def zero_(self):
        self.flat.zero_()

    def accumulate(self, grads):
        """grads: dict name -> tensor (any shape with P*width elements); adds one view's gradients."""
        for k, v in self.views.items():
            v.add_(grads[k].reshape(v.shape))

    def _alloc_registered(self, numel, device):
        try:
            import torch.distributed as dist
            if not (dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl"):
                return None
            backend = dist.group.WORLD._get_backend(torch.device(device))
            pool = torch.cuda.MemPool(backend.mem_allocator)
            with torch.cuda.use_mem_pool(pool):
                t = torch.empty(numel, dtype=torch.float32, device=device)
            backend.register_mem_pool(pool)
            self._pool, self.registered = pool, True
            return t
        except Exception as ex:  # pragma: no cover - depends on the NCCL / torch build
            self.registration_error = repr(ex)
            return None

    def all_reduce_begin(self, group=None):
        """Start the exchange of the feature field (call when it is final: after the blend stage of the backward)."""
        import torch.distributed as dist
        if self.split and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            self._work.append(dist.all_reduce(self.flat[:self.split], op=dist.ReduceOp.SUM, group=group, async_op=True))

    def all_reduce_finish(self, group=None):
        """Exchange the remaining fields and make the current stream wait for the whole message."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            rest = self.flat[self.split:] if self._work else self.flat
            self._work.append(dist.all_reduce(rest, op=dist.ReduceOp.SUM, group=group, async_op=True))
            for w in self._work:
                w.wait()
        self._work = []
        return self.flat

    def all_reduce(self, group=None):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return self.flat
