"""Pure-write, pure-read and copy bandwidth on this GPU (1 GB buffers, CUDA events): the ceiling for write-only kernels."""
import torch
DEV = "cuda"
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device=DEV); b = torch.empty(n, dtype=torch.uint8, device=DEV)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
s = t(lambda: a.zero_()); print(f"memset 1 GiB: {n / s / 1e12:.2f} TB/s written")
af = a.view(torch.float32)
s = t(lambda: af.fill_(1.5)); print(f"fill fp32 1 GiB: {n / s / 1e12:.2f} TB/s written")
s = t(lambda: b.copy_(a)); print(f"copy 1 GiB: {2 * n / s / 1e12:.2f} TB/s read+written")
s = t(lambda: af.sum()); print(f"sum fp32 1 GiB: {n / s / 1e12:.2f} TB/s read")
