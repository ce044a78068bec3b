"""One forward call per implementation at the outer / main shapes inside a profiler range (ncu launch list target)."""
import sys, os
sys.path.insert(0, "tests"); import _util
import torch, torch.nn.functional as F
from dcasr_b200 import ops
DEV = "cuda"; torch.manual_seed(0)
shapes = ((40, 398, 12, "outer"), (40, 196, 16, "main"), (10, 1498, 16, "large-outer"))
data = []
for (B, L, H, tag) in shapes:
    ndir = 2; di, N = 64 * H, 128
    xconv = (torch.randn(ndir, B * L, di + 2 * N, device=DEV) * 0.8).to(torch.bfloat16)
    dt = F.softplus(torch.randn(ndir, B * L, H, device=DEV) - 2.0)
    A_log = torch.log(torch.rand(ndir, H, device=DEV) * 15 + 1); Dk = torch.randn(ndir, H, device=DEV)
    data.append((xconv, dt, A_log, Dk, ndir, B, L, di, N, H))
impls = [int(a) for a in sys.argv[1:]] or [1, 4]
for d in data:
    for fi in impls:
        ops.ssd_fwd(*d, impl=fi)
torch.cuda.synchronize(); torch.cuda.profiler.start()
for d in data:
    for fi in impls:
        ops.ssd_fwd(*d, impl=fi)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("ok")
