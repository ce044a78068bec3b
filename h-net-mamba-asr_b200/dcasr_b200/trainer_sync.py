"""The reference Trainer's per-micro-batch host sync under DDP, without the GPU in the loop.

Reference: ``Trainer._any_rank_oom`` (/root/reference/src/dcasr/training/trainer.py:200-208) all-reduces a one-element CUDA
tensor over NCCL and reads it back with ``.item()`` after EVERY micro-batch when world_size > 1.  The flag itself is host
knowledge (a Python ``except torch.cuda.OutOfMemoryError``), but the read-back waits for everything queued on the stream --
the whole forward of the micro-batch -- so the host can never run ahead of the GPU and every backward starts with an empty
launch queue.  Here the same MAX-reduction runs over a host-side (gloo) process group on a CPU tensor: identical result on
every rank, exactly one matched collective per micro-batch as the reference requires, nothing enqueued on or read from the
CUDA stream.  ``dcasr_b200.install()`` rebinds the method when the reference's trainer module is importable.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

_HOST_GROUP = None


def host_group():
    """A gloo group over all ranks, created on first use (collectively: every rank reaches its first micro-batch)."""
    global _HOST_GROUP
    if _HOST_GROUP is None:
        _HOST_GROUP = dist.group.WORLD if dist.get_backend() == "gloo" else dist.new_group(backend="gloo")
    return _HOST_GROUP


def any_rank_flag(flag: bool, world_size: int) -> bool:
    """True on every rank iff `flag` is true on at least one (identity without a process group / on one rank)."""
    if world_size <= 1 or not (dist.is_available() and dist.is_initialized()):
        return bool(flag)
    t = torch.tensor([1 if flag else 0], dtype=torch.int32)          # host memory: the CUDA stream is not involved
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=host_group())
    return bool(int(t[0]) > 0)


def patch_trainer(trainer_cls) -> None:
    """Rebind ``trainer_cls._any_rank_oom`` (same signature, same return value on every rank)."""
    def _any_rank_oom(self, oom_local: bool) -> bool:
        return any_rank_flag(oom_local, self.world_size)

    _any_rank_oom.__doc__ = "host-side group OOM flag (dcasr_b200.trainer_sync): no CUDA sync per micro-batch"
    trainer_cls._any_rank_oom = _any_rank_oom
