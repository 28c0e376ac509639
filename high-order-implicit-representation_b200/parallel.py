"""Host-side multi-GPU plumbing (one process per GPU, torch.distributed over NCCL; gloo in the
CPU tests).  The path shards by pixels: training is data-parallel over disjoint
pixel blocks with ONE all-reduce (sum) of the flat [gradient | loss] vector per step; rendering is
sharded by contiguous pixel ranges with no collective."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def world_info(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def training_block(step: int, rank: int, world: int, n_blocks: int) -> int:
    """Pixel block trained by `rank` at `step`: ranks take consecutive blocks, the job walks the
    image round-robin, so within a step the ranks' blocks are disjoint whenever world <= n_blocks."""
    return (step * world + rank) % n_blocks


def render_shard(n_pixels: int, rank: int, world: int, align: int = 256) -> Tuple[int, int]:
    """(first_pixel, count) of the contiguous pixel range rendered by `rank`; ranges are disjoint,
    cover [0, n_pixels) and start on multiples of `align` (the kernel's tile)."""
    per = -(-n_pixels // world)
    per = -(-per // align) * align
    first = min(rank * per, n_pixels)
    return first, max(0, min(per, n_pixels - first))


def global_loss_scale(local_batch: int, world: int, out_features: int, global_batch: Optional[int] = None) -> float:
    """MSE(mean) over the GLOBAL batch (Quirk Q12): each rank scales its sum of squared errors and
    its gradients by 1/(B_global*out) so that the all-reduced sum equals the single-GPU result."""
    gb = global_batch if global_batch is not None else local_batch * world
    return 1.0 / (gb * out_features) if gb > 0 else 0.0


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks of the flat [gradient | loss] vector; no-op for a single rank."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat
