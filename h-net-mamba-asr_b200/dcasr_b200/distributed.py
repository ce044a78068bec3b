"""Data-parallel plumbing of the hot path: utterance sharding and the gradient all-reduce.

The path shards by utterance (every scan, reduction and recurrence is per row; the ratio loss is per rank, as under
the reference's DDP, src/dcasr/training/trainer.py:99-102).  The only exchange is the gradient all-reduce at the
end of a step: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> list[int]:
    """Stride slicing `items[rank:usable:world]` of the reference's DistributedBucketBatchSampler
    (src/dcasr/data/librispeech.py:195-196): every rank gets the same number of items."""
    usable = n_items - n_items % world
    return list(range(rank, usable, world))


class GradAllReducer:
    """Bucketed mean all-reduce of parameter gradients (DDP semantics) through flat buffers."""

    def __init__(self, params, bucket_mb: float = 64.0):
        self.params = [p for p in params if p.requires_grad]
        self.buckets, cur, size = [], [], 0
        limit = int(bucket_mb * (1 << 20))
        for p in self.params:
            cur.append(p)
            size += p.numel() * 4
            if size >= limit:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.flat = [torch.zeros(sum(p.numel() for p in b), dtype=torch.float32, device=b[0].device)
                     for b in self.buckets]

    def __call__(self) -> None:
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        world = dist.get_world_size()
        works = []
        for b, flat in zip(self.buckets, self.flat):
            views = flat.split([p.numel() for p in b])
            torch._foreach_copy_(list(views), [p.grad.reshape(-1).float() if p.grad is not None
                                               else torch.zeros_like(v) for p, v in zip(b, views)])
            works.append(dist.all_reduce(flat, async_op=True))
        for (b, flat), w in zip(zip(self.buckets, self.flat), works):
            w.wait()
            flat.div_(world)
            for p, v in zip(b, flat.split([p.numel() for p in b])):
                if p.grad is None:
                    p.grad = v.view_as(p).to(p.dtype).clone()
                else:
                    p.grad.copy_(v.view_as(p))
