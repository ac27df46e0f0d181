"""Image-sharded multi-GPU execution: one process per GPU, no data-path collective except the final result gather.

Every image is independent through backbone, neck, head, decode and NMS (eval-mode BatchNorm, per-image NMS,
reference ``utils/ops.py:250``), so the batch dimension is split contiguously over the ranks of a ``torch.distributed``
group and each rank runs its own engine replica.  The reference offers nothing here (``select_device`` collapses any
multi-GPU request to ``cuda:0`` outside training, ``utils/torch_utils.py:198-215``).

The only exchange is the gather of the fixed-capacity detections — ``[B_local, max_det, 6]`` rows + ``[B_local]`` counts,
<= 460 KB for 64 images — over NCCL (NVLink/NVSwitch) on GPUs or gloo on CPU (tests).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the images rank `rank` owns: contiguous, sizes differ by at most one, earlier ranks get the extra."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_shard(n: int, world: int) -> int:
    return (n + world - 1) // world


def gather_detections(rows: torch.Tensor, counts: torch.Tensor, n_total: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather the per-rank detections into batch order.

    rows [B_local, max_det, 6] f32 and counts [B_local] i32 of THIS rank's shard -> rows [n_total, max_det, 6], counts [n_total]
    on every rank.  Shards are padded to the largest shard so one fixed-shape collective serves uneven splits.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return rows[:n_total], counts[:n_total]
    cap = max_shard(n_total, world)
    b, max_det, six = rows.shape
    pad_rows = rows.new_zeros(cap, max_det, six)
    pad_cnt = counts.new_zeros(cap)
    pad_rows[:b], pad_cnt[:b] = rows, counts
    all_rows = rows.new_empty(world * cap, max_det, six)
    all_cnt = counts.new_empty(world * cap)
    dist.all_gather_into_tensor(all_rows, pad_rows, group=group)
    dist.all_gather_into_tensor(all_cnt, pad_cnt, group=group)
    keep = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        keep.append(torch.arange(r * cap, r * cap + (hi - lo), device=rows.device))
    idx = torch.cat(keep)
    return all_rows.index_select(0, idx), all_cnt.index_select(0, idx)


def detect_sharded(detect: Callable[[torch.Tensor], Tuple[torch.Tensor, torch.Tensor]], images: torch.Tensor, group=None,
                   rank: Optional[int] = None, world: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Run `detect` (images -> (rows [b,max_det,6], counts [b])) on this rank's contiguous shard of `images` and gather.

    `images` is the GLOBAL batch (every rank holds or can index it, e.g. a shared dataset order); only the shard is read.
    """
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_range(images.shape[0], rank, world)
    rows, counts = detect(images[lo:hi])
    return gather_detections(rows[: hi - lo], counts[: hi - lo], images.shape[0], group)
