"""Range-sharded multi-GPU MSM: one process per GPU (torchrun), `torch.distributed` for the plumbing.

The point set shards by contiguous range -- the GPU analogue of the reference's static thread split
`range()` (src/threads/threads.ts:354-359).  Every rank runs the complete single-GPU pipeline on
its range and produces one partial point; the only exchange is an all-gather of those partials
(world x 144 bytes), after which rank 0 adds them (the "partition sum / final sum" of
src/msm-batched-affine.ts:299-322, across GPUs).  MSM is linear, so no other collective exists.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of rank `rank`: ceil(n / world) items each, like src/threads/threads.ts:354-359."""
    per = -(-n // world)
    lo = min(n, per * rank)
    return lo, min(n, lo + per)


def gather_partials(partial: torch.Tensor, group=None) -> torch.Tensor:
    """All-gathers one fixed-size partial per rank (uint8 tensor, on the GPU with NCCL or on the CPU
    with gloo) into a (world * len) tensor, ordered by rank."""
    world = dist.get_world_size(group)
    out = torch.empty(world * partial.numel(), dtype=partial.dtype, device=partial.device)
    if partial.is_cuda:
        dist.all_gather_into_tensor(out, partial.contiguous(), group=group)
    else:
        dist.all_gather(list(out.view(world, -1).unbind(0)), partial.contiguous(), group=group)
    return out


class ShardedMsm:
    """Holds this rank's engine and point range.  `points` / `scalars` are this rank's shard."""

    def __init__(self, curve: str, device: Optional[int] = None, group=None):
        from .engine import MsmEngine
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.cuda.current_device() if device is None else device
        self.stream = torch.cuda.Stream(self.device)
        self.engine = MsmEngine(curve, device=self.device, stream=self.stream.cuda_stream)
        pb = self.engine.partial_bytes()
        self.partial = torch.zeros(pb, dtype=torch.uint8, device=torch.device("cuda", self.device))

    def set_bases(self, points, n_local: int, layout: int = 1, on_device: bool = False):
        if on_device:
            self.engine.set_bases_device(int(points), n_local, layout)
        else:
            self.engine.set_bases(points, n_local, layout)

    def msm(self, scalars, n_local: int, layout: int = 1, on_device: bool = False, window_bits: int = 0):
        """Returns the full MsmResult on rank 0, None elsewhere."""
        with torch.cuda.stream(self.stream):
            # MSM, collective and combine are ordered on this stream: no host synchronisation in between
            self.engine.run_partial(scalars, n_local, self.partial.data_ptr(), layout=layout,
                                    window_bits=window_bits, on_device=on_device, timing=False)
            if self.world == 1:
                return self.engine.combine(self.partial.data_ptr(), 1)
            allp = gather_partials(self.partial, self.group)
            if self.rank == 0:
                return self.engine.combine(allp.data_ptr(), self.world)
            self.stream.synchronize()
            return None
