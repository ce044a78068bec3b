"""ssd_fwd / ssd_bwd timings at the outer- and main-stack shapes (CUDA events, L2 flushed between reps).
bwd impl 1 = state-gradient pass + fused dx | dB/dC kernel, impl 3 = the three-kernel backward of round 1."""
import sys, os
sys.path.insert(0, "tests"); import _util
import torch, torch.nn.functional as F
from dcasr_b200 import ops
from dcasr_b200._lib import lib
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
shapes = ((40, 398, 12, "outer"), (40, 196, 16, "main"), (10, 1498, 16, "large-outer"), (10, 640, 24, "large-main"))
for (B, L, H, tag) in shapes:
    ndir = 2; di, N = 64 * H, 128
    xconv = (torch.randn(ndir, B * L, di + 2 * N, device=DEV) * 0.8).to(torch.bfloat16)
    dt = F.softplus(torch.randn(ndir, B * L, H, device=DEV) - 2.0)
    A_log = torch.log(torch.rand(ndir, H, device=DEV) * 15 + 1); Dk = torch.randn(ndir, H, device=DEV)
    dy = (torch.randn(ndir, B * L, di, device=DEV) * 0.5).to(torch.bfloat16)
    y, ws = ops.ssd_fwd(xconv, dt, A_log, Dk, ndir, B, L, di, N, H, impl=1)
    res = []
    for fi in (5, 4):          # 5: split (states pass + scan), 4: persistent two-CTA kernel
        tf = timeit(lambda: ops.ssd_fwd(xconv, dt, A_log, Dk, ndir, B, L, di, N, H, impl=fi))
        res.append(f"fwd impl {fi} {tf:.1f}")
    for impl in (1, 3):
        parts = int(lib().raw("ssd_dbc_parts")(ndir, B, L, H, impl))
        tb = timeit(lambda: ops.ssd_bwd(dy, xconv, y, dt, A_log, Dk, ws, ndir, B, L, di, N, H, impl=impl, keep_parts=True))
        res.append(f"bwd impl {impl} ({parts} part{'s' if parts > 1 else ''}) {tb:.1f}")
    out.append(f"{tag} B={B} L={L} H={H}: " + ", ".join(res) + " us")
print("\n".join(out), flush=True)
