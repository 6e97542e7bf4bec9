"""Multi-GPU host logic: the batch shards by sample, one process per GPU (torchrun), no data-path
collective inside the kernels.  Two NCCL all-reduces per step carry everything that is global:

  1. after loss_forward : the 8 loss partial sums (the mean-RGB term, losses.py:369, is the square of a
     difference of GLOBAL means, so its gradient scale needs them before loss_backward runs);
  2. after the backward : gradients of parameters shared by the batch (the texture map).

Per-sample pose / shape / light gradients never leave their rank.  The reference's only multi-device
mechanism is nn.DataParallel (train_hrnet.py:560); this replaces it for the hot path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ._lib import LOSS_NSUMS


def shard_range(n_global: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of the global batch owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_global, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_loss_sums(sums: torch.Tensor, group=None):
    """Sum the global partial sums in place; the per-sample IoU sums that follow them stay local."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums[:LOSS_NSUMS], group=group)
    return sums


def all_reduce_loss_sums_async(sums: torch.Tensor, group=None):
    """As all_reduce_loss_sums, but only STARTS the collective and returns the work handles (empty at world size 1):
    FusedHandStep.backward waits for them after it has enqueued the loss backward, which does not need the sums."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        return [dist.all_reduce(sums[:LOSS_NSUMS], group=group, async_op=True)]
    return []


def all_reduce_shared_grads(*grads: torch.Tensor, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for g in grads:
            dist.all_reduce(g, group=group)


def all_reduce_shared_grads_async(*grads: torch.Tensor, group=None):
    """Start the all-reduce of the shared-parameter gradients and return the work handles (empty at world size 1).
    FusedHandStep.backward calls this right after the shade/rasterize backward, so the NVLink transfer overlaps the
    geometry and hand-layer backward kernels; `wait_all` makes the current stream wait for it."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        return [dist.all_reduce(g, group=group, async_op=True) for g in grads]
    return []


def wait_all(works):
    for w in works:
        w.wait()


def loss_terms_from_sums(sums: torch.Tensor, n_local: int, n_global: int, H: int, W: int, group=None):
    """[texture, mrgb, ssim_tex, sil, iou] from all-reduced sums (IoU: local sum of per-sample IoUs, reduced)."""
    cnt = float(n_global * 3 * H * W)
    mul, add = sums[LOSS_NSUMS:LOSS_NSUMS + n_local], sums[LOSS_NSUMS + n_local:LOSS_NSUMS + 2 * n_local]
    iou_sum = (mul / (add - mul)).sum()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        iou_sum = iou_sum.clone()
        dist.all_reduce(iou_sum, group=group)
    return torch.stack([sums[0] / cnt, (sums[2] / cnt - sums[1] / cnt) ** 2, 1 - sums[4] / cnt,
                        sums[3] / float(n_global * H * W), 1 - iou_sum / n_global])
