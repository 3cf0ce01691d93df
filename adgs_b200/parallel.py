"""Multi-GPU data path: the views of a multi-camera / multi-timestep batch are sharded across
ranks (one process per GPU), every rank renders its own views against a full parameter replica,
and the parameter gradients are summed with ONE all-reduce over a flat gradient buffer
(NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests).

The reference has no distributed code at all: it renders one view per
iteration (train.py:55-61). The oracle for this module is therefore "sum over views of the
single-GPU gradients" (SURVEY.md section 8e), plus the three non-linear densification statistics
that a plain gradient sum does not cover (gaussian_model.py:863-867, train.py:151):
    xyz_gradient_accum += || grad_means2D[:, :2] ||   per view   -> all-reduce SUM of per-view norms
    denom              += visible                      per view   -> all-reduce SUM
    max_radii2D         = max(max_radii2D, radii)      per view   -> all-reduce MAX
"""
from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


def shard_views(views: Sequence, rank: int, world_size: int) -> List:
    """Rank r renders views r, r+G, r+2G, ... (round-robin keeps cameras of one timestep apart)."""
    return list(views[rank::world_size])


class FlatGradBucket:
    """All hot-path parameters' gradients as views of one contiguous fp32 buffer, so the
    cross-rank reduction is a single collective launch (no per-tensor latency, full NVLink
    message size)."""

    def __init__(self, params: Dict[str, torch.Tensor]):
        self.names = list(params)
        self.shapes = {k: tuple(v.shape) for k, v in params.items()}
        self.offsets, off = {}, 0
        for k, v in params.items():
            self.offsets[k] = off
            off += (v.numel() + 3) // 4 * 4      # keep every view 16-byte aligned
        self.numel = off
        any_p = next(iter(params.values()))
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=any_p.device)

    def views(self) -> Dict[str, torch.Tensor]:
        out = {}
        for k in self.names:
            n = 1
            for s in self.shapes[k]:
                n *= s
            out[k] = self.flat[self.offsets[k]: self.offsets[k] + n].view(self.shapes[k])
        return out

    def zero_(self):
        self.flat.zero_()

    def accumulate(self, grads: Dict[str, torch.Tensor]):
        v = self.views()
        for k, g in grads.items():
            if g is not None:
                v[k].add_(g)

    def all_reduce(self, group=None, async_op=False):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return None


class DensifyStats:
    """Per-view densification statistics and their cross-rank reductions."""

    def __init__(self, n: int, device):
        self.grad_norm_sum = torch.zeros(n, 1, device=device)
        self.visible_count = torch.zeros(n, 1, device=device)
        self.max_radii = torch.zeros(n, device=device)

    def add_view(self, viewspace_grad: torch.Tensor, visibility_filter: torch.Tensor, radii: torch.Tensor):
        vis = visibility_filter
        self.grad_norm_sum[vis] += torch.norm(viewspace_grad[vis, :2], dim=-1, keepdim=True)
        self.visible_count[vis] += 1
        self.max_radii[vis] = torch.max(self.max_radii[vis], radii[vis].to(self.max_radii.dtype))

    def all_reduce(self, group=None):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad_norm_sum, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.visible_count, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self.max_radii, op=dist.ReduceOp.MAX, group=group)


class MultiViewStep:
    """One multi-view iteration: render + backward the local shard of `views`, then reduce.

    `render_fn(view)` must return the result dict of gaussian_renderer.render();
    `cotangent_fn(view, result)` returns (outputs, cotangents) for torch.autograd.backward
    (in training: the loss; in the benchmark: fixed random cotangents)."""

    def __init__(self, model, group=None):
        self.model = model
        self.group = group
        self.rank = dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        from .gaussian_model import PARAM_NAMES
        self.names = PARAM_NAMES
        self.bucket = FlatGradBucket({k: getattr(model, k) for k in PARAM_NAMES})
        self.stats = DensifyStats(model.get_pts_num, model.xyz.device)

    def run(self, views, render_fn, cotangent_fn, reduce_stats=True):
        local = shard_views(views, self.rank, self.world)
        params = {k: getattr(self.model, k) for k in self.names}
        first = True
        for view in local:
            for p in params.values():
                p.grad = None
            # first local view: the backward writes straight into the flat bucket (no copy)
            self.model._grad_sink = self.bucket.views if first else None
            try:
                res = render_fn(view)
                outs, cots = cotangent_fn(view, res)
                torch.autograd.backward(outs, cots)
            finally:
                self.model._grad_sink = None
            grads = {k: p.grad for k, p in params.items()}
            v = self.bucket.views()
            if first:
                for k, g in grads.items():
                    if g is None:
                        v[k].zero_()
                    elif g.data_ptr() != v[k].data_ptr():   # autograd cloned it (should not happen)
                        v[k].copy_(g)
                first = False
            else:
                self.bucket.accumulate(grads)
            if reduce_stats:
                self.stats.add_view(res["viewspace_points"].grad, res["visibility_filter"], res["radii"])
        if first:
            self.bucket.zero_()
        self.bucket.all_reduce(self.group)
        if reduce_stats:
            self.stats.all_reduce(self.group)
        v = self.bucket.views()
        for k, p in params.items():
            p.grad = v[k]
        return v
