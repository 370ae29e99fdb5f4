"""Image-sharded multibox loss over several GPUs.

Matching, CE, mining and NMS are independent per image, so a batch shards by contiguous image ranges with no
data-path collective.  The only exchange is the reference's batch-global normalisation (Losses.py:182,197):
every rank all-reduces its positive count BEFORE the mining kernel (the count scales the gradients it writes)
and the two loss sums after it - 4 + 16 bytes over NCCL.  Detect needs no collective at all.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def shard_range(rank: int, world: int, batch: int) -> Tuple[int, int]:
    """Contiguous image range [lo, hi) of `rank`; the first batch % world ranks get one extra image."""
    if not (0 <= rank < world) or batch < 0:
        raise ValueError(f"bad shard request rank={rank} world={world} batch={batch}")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_lists(rank: int, world: int, *lists: Sequence):
    """Slice per-image sequences (tensors with a leading batch dim or Python lists) to this rank's range."""
    n = len(lists[0])
    lo, hi = shard_range(rank, world, n)
    return tuple(x[lo:hi] for x in lists)


def allreduce_loss_parts(sum_l1: torch.Tensor, sum_ce: torch.Tensor, npos: torch.Tensor, group=None):
    """Combine per-rank partial results into the global (loc_loss, conf_loss) of ``ssd()``:
    loc = sum|d| / (4 * Npos), conf = sum CE / Npos with Npos the batch-global positive count.
    Works with any backend (NCCL on GPUs, gloo in the CPU tests); tensors are reduced in place."""
    import torch.distributed as dist
    dist.all_reduce(npos, op=dist.ReduceOp.SUM, group=group)
    parts = torch.stack([sum_l1.to(torch.float64), sum_ce.to(torch.float64)])
    dist.all_reduce(parts, op=dist.ReduceOp.SUM, group=group)
    n = npos.to(torch.float64)
    return (parts[0] / (4.0 * n)).to(torch.float32), (parts[1] / n).to(torch.float32)


def sharded_ssd(head, loc: torch.Tensor, conf: torch.Tensor, gt_boxes: List[torch.Tensor],
                gt_classes: List[torch.Tensor], group=None):
    """``ssd()`` on this rank's shard of a batch sharded by image: returns the GLOBAL (loc_loss, conf_loss)
    with gradients of the global loss w.r.t. this rank's loc / conf."""
    from .head import multibox_loss
    import torch.distributed as dist
    if group is None:
        group = dist.group.WORLD
    return multibox_loss(head, loc, conf, gt_boxes, gt_classes, group=group)
