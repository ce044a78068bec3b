"""ssd_fwd / ssd_bwd timings at the outer- and main-stack shapes (CUDA events, L2 flushed between reps)."""
import sys, os
sys.path.insert(0, "tests"); import _util
import torch, torch.nn.functional as F
from dcasr_b200 import ops
DEV = "cuda"; torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
def timeit(fn, reps=5, inner=10):
    for _ in range(2): fn()
    ts = []
    for _ in range(reps):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner): fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / inner)
    ts.sort(); return ts[len(ts) // 2]
out = []
for (L, H, tag) in ((398, 12, "outer"), (196, 16, "main")):
    ndir, B = 2, 40; di, N = 64 * H, 128
    xconv = (torch.randn(ndir, B * L, di + 2 * N, device=DEV) * 0.8).to(torch.bfloat16)
    dt = F.softplus(torch.randn(ndir, B * L, H, device=DEV) - 2.0)
    A_log = torch.log(torch.rand(ndir, H, device=DEV) * 15 + 1); Dk = torch.randn(ndir, H, device=DEV)
    dy = (torch.randn(ndir, B * L, di, device=DEV) * 0.5).to(torch.bfloat16)
    y, ws = ops.ssd_fwd(xconv, dt, A_log, Dk, ndir, B, L, di, N, H, impl=1)
    tf = timeit(lambda: ops.ssd_fwd(xconv, dt, A_log, Dk, ndir, B, L, di, N, H, impl=1))
    tb = timeit(lambda: ops.ssd_bwd(dy, xconv, y, dt, A_log, Dk, ws, ndir, B, L, di, N, H, impl=1, keep_parts=True))
    out.append(f"{tag}: fwd {tf:.1f} us, bwd {tb:.1f} us")
print(os.environ.get("HNB_SSD_DX_HEADS", "auto"), " | ".join(out), flush=True)
