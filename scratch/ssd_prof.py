"""Per-kernel durations of ssd_bwd (torch.profiler / CUPTI) at the outer- and main-stack shapes, impl 1 vs 3."""
import sys, os
sys.path.insert(0, "tests"); import _util
import torch, torch.nn.functional as F
from torch.profiler import profile, ProfilerActivity
from dcasr_b200 import ops
DEV = "cuda"; torch.manual_seed(0)
for (B, L, H, tag) in ((40, 398, 12, "outer"), (40, 196, 16, "main")):
    ndir = 2; di, N = 64 * H, 128
    xconv = (torch.randn(ndir, B * L, di + 2 * N, device=DEV) * 0.8).to(torch.bfloat16)
    dt = F.softplus(torch.randn(ndir, B * L, H, device=DEV) - 2.0)
    A_log = torch.log(torch.rand(ndir, H, device=DEV) * 15 + 1); Dk = torch.randn(ndir, H, device=DEV)
    dy = (torch.randn(ndir, B * L, di, device=DEV) * 0.5).to(torch.bfloat16)
    y, ws = ops.ssd_fwd(xconv, dt, A_log, Dk, ndir, B, L, di, N, H, impl=1)
    for impl in (1, 3):
        for _ in range(3):
            ops.ssd_bwd(dy, xconv, y, dt, A_log, Dk, ws, ndir, B, L, di, N, H, impl=impl, keep_parts=True)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(10):
                ops.ssd_bwd(dy, xconv, y, dt, A_log, Dk, ws, ndir, B, L, di, N, H, impl=impl, keep_parts=True)
            torch.cuda.synchronize()
        rows = [(e.key[:60], e.count, e.device_time_total / max(e.count, 1)) for e in prof.key_averages() if e.device_time_total > 0]
        print(tag, "impl", impl, [(k, n, round(t, 1)) for k, n, t in sorted(rows, key=lambda r: -r[2])], flush=True)
