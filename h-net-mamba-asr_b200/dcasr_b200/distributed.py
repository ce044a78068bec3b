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
    """Bucketed mean all-reduce of parameter gradients (DDP semantics) through flat fp32 buffers.

    ``overlap=True`` registers post-accumulate-grad hooks: a bucket's all-reduce is launched (async, on NCCL's own
    stream) as soon as its last gradient has been produced, so the exchange overlaps the rest of the backward pass;
    ``__call__`` (end of the step) launches whatever is still pending, waits, and writes the means back into ``.grad``.
    Buckets are filled in reverse parameter order, the order in which backward produces gradients."""

    def __init__(self, params, bucket_mb: float = 64.0, overlap: bool = False):
        self.params = [p for p in params if p.requires_grad]
        self.buckets, cur, size = [], [], 0
        limit = int(bucket_mb * (1 << 20))
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * 4
            if size >= limit:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.flat = [torch.zeros(sum(p.numel() for p in b), dtype=torch.float32, device=b[0].device)
                     for b in self.buckets]
        # persistent views of the flat buffers, shaped like their parameters: a step copies gradients in and out with
        # one multi-tensor launch per bucket and no per-parameter tensor ops on the host (a reshape + a cast going in
        # and a view coming out per parameter were ~1000 small host ops, ~1.5 ms, per step of the 330-parameter encoder)
        self.views = [[v.view_as(p) for p, v in zip(b, flat.split([p.numel() for p in b]))]
                      for b, flat in zip(self.buckets, self.flat)]
        self.works = [None] * len(self.buckets)
        self.pending = [len(b) for b in self.buckets]
        self.overlap = overlap and dist.is_initialized() and dist.get_world_size() > 1
        if self.overlap:
            where = {id(p): i for i, b in enumerate(self.buckets) for p in b}
            for p in self.params:
                p.register_post_accumulate_grad_hook(lambda q, i=where[id(p)]: self._ready(i))

    def _launch(self, i: int) -> None:
        b, views = self.buckets[i], self.views[i]
        if all(p.grad is not None for p in b):
            torch._foreach_copy_(views, [p.grad for p in b])
        else:                                                    # a parameter that got no gradient contributes zeros
            have = [(v, p.grad) for p, v in zip(b, views) if p.grad is not None]
            for p, v in zip(b, views):
                if p.grad is None:
                    v.zero_()
            if have:
                torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        self.works[i] = dist.all_reduce(self.flat[i], async_op=True)

    def _ready(self, i: int) -> None:
        self.pending[i] -= 1
        if self.pending[i] == 0:
            self._launch(i)

    def __call__(self) -> None:
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        world = dist.get_world_size()
        for i in range(len(self.buckets)):
            if self.works[i] is None:
                self._launch(i)
        for i, (b, flat) in enumerate(zip(self.buckets, self.flat)):
            self.works[i].wait()
            flat.div_(world)
            for p in b:
                if p.grad is None:
                    p.grad = torch.empty_like(p)
            torch._foreach_copy_([p.grad for p in b], self.views[i])   # one launch per bucket
            self.works[i] = None
            self.pending[i] = len(b)


class HostBatchPrefetcher:
    """Double-buffered host -> device staging of input batches on a side stream.

    ``push(*pinned_host_tensors)`` starts the copies on the prefetcher's own stream; ``pop()`` makes the compute stream
    wait for them and returns the device tensors.  Pushing batch i+1 before running step i hides the copy (20 MB of
    log-mel features per 40 x 16 s batch, ~0.4 ms over PCIe) behind step i's kernels -- the role the reference gives to
    its DataLoader's pin_memory + non_blocking copies (src/dcasr/training/trainer.py:187-190)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._slot = None

    def empty(self) -> bool:
        return self._slot is None

    def push(self, *host_tensors) -> None:
        if self._slot is not None:
            raise RuntimeError("HostBatchPrefetcher: one batch is already in flight")
        with torch.cuda.stream(self.stream):
            devs = [t.to(self.device, non_blocking=True) for t in host_tensors]
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._slot = (devs, ev)

    def pop(self):
        if self._slot is None:
            raise RuntimeError("HostBatchPrefetcher: nothing was pushed")
        devs, ev = self._slot
        self._slot = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in devs:
            t.record_stream(cur)       # the caching allocator must not hand the buffer back while `cur` still reads it
        return devs
