"""Split-K sweep of the weight-gradient GEMMs at the headline shapes."""
import sys
sys.path.insert(0, "tests"); import _util
import torch
from dcasr_b200 import ops
DEV = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
def timeit(fn, reps=5, inner=6):
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
def bf(*shape): return (torch.randn(*shape, device=DEV) * 0.5).to(torch.bfloat16)
for (T, d, di, H, tag) in ((15920, 384, 768, 12, "outer"), (7840, 512, 1024, 16, "main")):
    N = 128; dip = 2 * di + 2 * N + H; ds = (dip + 7) // 8 * 8; ldz = 2 * ds
    h = bf(T, d); zx = bf(T, ldz); yn = bf(T, 2 * di); x = bf(T, d)
    for name, a, b_, M_, N_ in (("in_proj wgrad", zx, h, ldz, d), ("out_proj wgrad", x, yn, d, 2 * di)):
        hint = ops.wgrad_splitk(T, M_, N_)
        res = []
        for sk in (1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 24):
            us = timeit(lambda: ops.gemm(a, b_, trans_a=True, trans_b=True, splitk=sk, out_dtype=torch.float32))
            res.append(f"{sk}:{us:.1f}")
        print(f"{tag} {name} [{M_}x{N_}x{T}] hint sk={hint}  " + "  ".join(res), flush=True)
