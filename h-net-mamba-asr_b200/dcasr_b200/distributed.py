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
    ``__call__`` (end of the step) launches whatever is still pending, waits, and leaves every ``.grad`` pointing AT its
    slice of the reduced flat buffer (no copy back, no separate division pass: the reduction op is AVG).
    Buckets are filled in reverse parameter order, the order in which backward produces gradients; the bucket that
    completes LAST (the first parameters of the model) is kept small (``tail_mb``) because its all-reduce is the one
    that cannot hide behind backward compute.

    Gradient accumulation (the reference trainer's ``accum_grad > 1``, src/dcasr/training/trainer.py:196-260): wrap
    every micro-batch but the last in ``with reducer.no_sync():`` -- hooks then only let autograd accumulate, exactly
    like DDP's ``no_sync``.  A hook that fires for a bucket whose all-reduce is already in flight is an error (it would
    silently publish a partial sum), and so is a parameter that received no gradient on this rank when the step ends
    with hooks outstanding: buckets are always launched in index order so that every rank issues the same collectives."""

    def __init__(self, params, bucket_mb: float = 64.0, overlap: bool = False, tail_mb: float = 4.0):
        self.params = [p for p in params if p.requires_grad]
        self.buckets, cur, size = [], [], 0
        limit = int(bucket_mb * (1 << 20))
        tail = int(min(tail_mb, bucket_mb) * (1 << 20))
        total = sum(p.numel() * 4 for p in self.params)
        seen = 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * 4
            seen += p.numel() * 4
            # close a bucket when it is full, or when what remains is the small tail bucket
            if size >= limit or (total - seen <= tail and total - seen > 0 and size >= tail):
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.flat = [torch.zeros(sum(p.numel() for p in b), dtype=torch.float32, device=b[0].device)
                     for b in self.buckets]
        # persistent views of the flat buffers, shaped like their parameters: a step copies gradients in with one
        # multi-tensor launch per bucket and afterwards hands the views out as the parameters' .grad
        self.views = [[v.view_as(p) for p, v in zip(b, flat.split([p.numel() for p in b]))]
                      for b, flat in zip(self.buckets, self.flat)]
        self.works = [None] * len(self.buckets)
        self.pending = [len(b) for b in self.buckets]
        self.next_launch = 0                      # buckets go out strictly in index order (same on every rank)
        self._sync = True
        self._op = None
        self.overlap = overlap and dist.is_initialized() and dist.get_world_size() > 1
        if self.overlap:
            where = {id(p): i for i, b in enumerate(self.buckets) for p in b}
            for p in self.params:
                p.register_post_accumulate_grad_hook(lambda q, i=where[id(p)]: self._ready(i))

    # -- accumulation ----------------------------------------------------------------------------------------------
    class _NoSync:
        def __init__(self, red):
            self.red = red

        def __enter__(self):
            self.prev, self.red._sync = self.red._sync, False

        def __exit__(self, *a):
            self.red._sync = self.prev

    def no_sync(self):
        """Context manager for the non-final micro-batches of an accumulated step (DDP.no_sync semantics)."""
        return GradAllReducer._NoSync(self)

    # -- internals -------------------------------------------------------------------------------------------------
    def _reduce_op(self):
        if self._op is None:
            # NCCL averages inside the collective; gloo (CPU tests) has no AVG: SUM and one division in __call__
            self._op = dist.ReduceOp.AVG if dist.get_backend() == "nccl" else dist.ReduceOp.SUM
        return self._op

    def _launch(self, i: int) -> None:
        b, views = self.buckets[i], self.views[i]
        src = [(v, p.grad) for p, v in zip(b, views) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        for p, v in zip(b, views):
            if p.grad is None:                                   # a parameter that got no gradient contributes zeros
                v.zero_()
        if src:
            torch._foreach_copy_([v for v, _ in src], [g for _, g in src])
        self.works[i] = dist.all_reduce(self.flat[i], op=self._reduce_op(), async_op=True)

    def _ready(self, i: int) -> None:
        if not self._sync:
            return
        if self.works[i] is not None or self.pending[i] <= 0:
            raise RuntimeError(
                "GradAllReducer: a gradient arrived for a bucket whose all-reduce is already in flight -- a second "
                "backward pass before reducer(); wrap all micro-batches but the last in `with reducer.no_sync():`")
        self.pending[i] -= 1
        while self.next_launch < len(self.buckets) and self.pending[self.next_launch] == 0 \
                and self.works[self.next_launch] is None:
            self._launch(self.next_launch)
            self.next_launch += 1

    def __call__(self) -> None:
        """Finish the step's exchange: launch what is pending, wait, and point every .grad at its reduced slice."""
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return
        world = dist.get_world_size()
        for i in range(len(self.buckets)):                       # whatever the hooks have not launched, in index order
            if self.works[i] is None:
                self._launch(i)
        avg_in_op = self._reduce_op() != dist.ReduceOp.SUM
        for i, (b, flat) in enumerate(zip(self.buckets, self.flat)):
            self.works[i].wait()
            if not avg_in_op:
                flat.div_(world)
            for p, v in zip(b, self.views[i]):                   # .grad IS the reduced slice: no copy back
                p.grad = v
            self.works[i] = None
            self.pending[i] = len(b)
        self.next_launch = 0


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
