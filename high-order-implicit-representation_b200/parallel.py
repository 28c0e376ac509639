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


def tile_schedule(order, n_samples: int, batch: int, rank: int, world: int):
    """Steps of one epoch over contiguous blocks of ``batch`` samples visited in ``order`` (a permutation of the
    block indices, identical on every rank): a list of ``(first, count, global_batch)`` with the SAME length on every
    rank.  Ranks take consecutive entries of ``order``; when ``len(order) % world != 0`` the last group has fewer
    blocks than ranks and the surplus ranks get an EMPTY share ``(0, 0, global_batch)`` -- they still step, because
    the gradient exchange is collective (include/hoir.h: every rank calls every step)."""
    steps = []
    for k in range(0, len(order), world):
        blocks = order[k:k + world]
        sizes = [min(batch, n_samples - b * batch) for b in blocks]
        if rank < len(blocks):
            steps.append((blocks[rank] * batch, sizes[rank], sum(sizes)))
        else:
            steps.append((0, 0, sum(sizes)))
    return steps


def shuffle_schedule(n_samples: int, batch: int, rank: int, world: int, drop_last: bool = True):
    """Steps of one epoch over a shuffled index vector ``perm`` of ``n_samples`` entries: a list of
    ``(lo, hi, global_batch)``; this rank's share of a step is ``perm[lo:hi][rank::world]`` (possibly empty on the
    ragged last batch -- still a step).  ``drop_last`` drops a ragged last global batch like the reference's
    DataLoader (/root/reference/high_order_implicit_representation/single_image_dataset.py:312-320), except when it
    is the only one."""
    steps = []
    for k in range(0, n_samples, batch * world):
        hi = min(n_samples, k + batch * world)
        if drop_last and hi - k < batch * world and k > 0:
            break
        steps.append((k, hi, hi - k))
    return steps


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


class PeerBlock:
    """This rank's symmetric gradient block (include/hoir.h: hoir_peer_desc) and its peers' mappings, obtained
    through ``torch.distributed._symmetric_memory`` (CUDA IPC / fabric handles exchanged over the process
    group's store) -- plumbing only: the exchange itself happens inside the training kernel's optimizer tail
    with release / acquire flags over NVLink, no NCCL call on the step path."""

    def __init__(self, n_floats: int, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.tensor = symm.empty(n_floats, dtype=torch.float32, device=device)
        self.tensor.zero_()
        self.handle = symm.rendezvous(self.tensor, self.group)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(self.ptrs) != self.world or any(p == 0 for p in self.ptrs):
            raise RuntimeError("symmetric-memory rendezvous returned no peer mappings")
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)          # every rank's block is zeroed before anyone's first step


def try_peer_block(n_floats: int, device: torch.device, group=None):
    """PeerBlock, or None when peer memory is not available (single rank, gloo, no P2P, HOIR_NO_PEER=1):
    the caller then falls back to one NCCL all-reduce per step."""
    import os
    if os.environ.get("HOIR_NO_PEER") or not (dist.is_available() and dist.is_initialized()):
        return None
    if dist.get_world_size(group) < 2 or dist.get_backend(group) != "nccl" or dist.get_world_size(group) > 16:
        return None
    try:
        return PeerBlock(n_floats, device, group)
    except Exception as e:  # noqa: BLE001 -- any failure of the optional plumbing means "use NCCL"
        import logging
        logging.getLogger(__name__).warning("peer-memory gradient exchange unavailable (%s); using NCCL all-reduce", e)
        return None
