"""Trainer._any_rank_oom per micro-batch (reference training/trainer.py:200-208) while the GPU has a forward's worth of work
queued: the reference's NCCL all-reduce + .item() against dcasr_b200.trainer_sync's host-side collective.  torchrun, 2 ranks."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "h-net-mamba-asr_b200"))
import torch, torch.distributed as dist
from dcasr_b200.trainer_sync import any_rank_flag
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world = dist.get_world_size()
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)

def queue_work():                       # ~8 ms of GPU work, as the forward of a micro-batch leaves behind
    for _ in range(8):
        a @ a

def ref_flag(flag):                     # the reference's implementation, verbatim semantics
    t = torch.tensor([1.0 if flag else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return bool(t.item() > 0)

for name, fn in (("reference: NCCL all-reduce + .item()", ref_flag), ("dcasr_b200.trainer_sync: host-side gloo", lambda f: any_rank_flag(f, world))):
    for _ in range(3):
        queue_work(); fn(False)
    torch.cuda.synchronize(); dist.barrier()
    blocked = []
    for _ in range(20):
        queue_work()
        t0 = time.perf_counter(); fn(False); blocked.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    if rank == 0:
        blocked.sort()
        print(f"{name}: host blocked {1e3 * blocked[len(blocked) // 2]:.3f} ms per micro-batch (median of 20, {world} ranks, GPU queue ~8 ms deep)")
dist.destroy_process_group()
