"""fp32 (decode) SSD forward on the CUDA-core exact path at the headline shapes: CUDA-event time per call and error vs fp64."""
import sys
sys.path.insert(0, "tests"); import _util
import torch
from dcasr_b200 import ops
DEV = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
def timeit(fn, reps=3, inner=4):
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
torch.manual_seed(0)
for (T, di, H, tag) in ((15920, 768, 12, "outer"), (7840, 1024, 16, "main")):
    N = 128; C = di + 2 * N; B = 40; L = T // B
    xconv = torch.randn(2, T, C, device=DEV) * 0.5
    dt = torch.rand(2, T, H, device=DEV) * 0.1 + 0.01
    Al = torch.log(torch.rand(2, H, device=DEV) * 15 + 1); Dk = torch.randn(2, H, device=DEV)
    us = timeit(lambda: ops.ssd_fwd(xconv, dt, Al, Dk, 2, B, L, di, N, H, impl=0))
    y, _ = ops.ssd_fwd(xconv, dt, Al, Dk, 2, B, L, di, N, H, impl=0)
    print(f"{tag} exact fp32 ssd_fwd {us:.0f} us  checksum {float(y.double().abs().mean()):.8f}", flush=True)
if len(sys.argv) > 1:        # one more call inside the profiler range
    torch.cuda.synchronize(); torch.cuda.profiler.start()
    ops.ssd_fwd(xconv, dt, Al, Dk, 2, B, L, di, N, H, impl=0)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
