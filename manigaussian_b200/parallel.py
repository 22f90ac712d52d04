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

FIELDS = ("dL_dmeans3D", "dL_dscales", "dL_drotations", "dL_dopacity", "dL_dsh", "dL_dfeature", "dL_dcolors", "dL_dmeans2D")


def shard_views(total_views, rank, world):
    """Round-robin view ownership: view v belongs to rank v % world."""
    return [v for v in range(total_views) if v % world == rank]


class PackedGradients:
    def __init__(self, P, F, M, device, colors=False, means2D=False, zero=True):
        """colors=True adds a [P,3] field for precomputed-colour gradients (used instead of SH when M == 0); means2D=True a
        [P,3] field for the screen-space gradients summed over views (single-view callers of the raw accumulate mode).
        zero=False leaves the buffer uninitialised (rasterize_views_backward_raw overwrites every row)."""
        self.widths = dict(dL_dmeans3D=3, dL_dscales=3, dL_drotations=4, dL_dopacity=1, dL_dsh=3 * M, dL_dfeature=F,
                           dL_dcolors=3 if colors else 0, dL_dmeans2D=3 if means2D else 0)
        self.P = P
        # every field starts on a 16-byte boundary so that 128-bit reductions can target it directly
        offs, off = {}, 0
        for k in FIELDS:
            offs[k] = off
            off += (P * self.widths[k] + 3) // 4 * 4
        self.flat = (torch.zeros if zero else torch.empty)(off, dtype=torch.float32, device=device)
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

    def all_reduce(self, group=None):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return self.flat
