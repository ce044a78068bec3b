"""conv_fwd / conv_bwd / gated_norm / layernorm timings at the headline shapes (L2 flushed between reps)."""
import sys, os
sys.path.insert(0, "tests"); import _util
import torch
from dcasr_b200 import ops
DEV = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
def timeit(fn, reps=5, inner=8):
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
print("variant", os.environ.get("HNB_CONV_BWD_VARIANT", "0"))
for (T, d, di, H, tag) in ((15920, 384, 768, 12, "outer"), (8400, 512, 1024, 16, "main")):
    N = 128; C = di + 2 * N; dip = 2 * di + 2 * N + H; ds = (dip + 7) // 8 * 8; ldz = 2 * ds
    B = 40; L = T // B
    zx = bf(T, ldz); x = bf(T, d); h = bf(T, d)
    cw, cb = torch.randn(2, C, 4, device=DEV), torch.randn(2, C, device=DEV)
    dtb, nw = torch.randn(2, H, device=DEV), torch.randn(2, di, device=DEV)
    lens = torch.full((B,), L, dtype=torch.int32, device=DEV)
    us = timeit(lambda: ops.conv_fwd(zx, ds, lens, cw, cb, dtb, 2, B, L, di, N, H)); by = 2 * T * (C * 4 + H * 6)
    print(f"{tag} conv_fwd {us:.1f} us  {by / us / 1e3:.0f} GB/s")
    dxc, dBC, ddt = bf(2, T, di), bf(2, T, 2 * N), torch.randn(2, T, H, device=DEV)
    dzx = torch.zeros_like(zx)
    us = timeit(lambda: ops.conv_bwd(zx, dxc, dBC, ddt, ds, lens, cw, cb, dtb, 2, B, L, di, N, H, dzx)); by = 2 * T * (C * 6 + H * 10)
    print(f"{tag} conv_bwd {us:.1f} us  {by / us / 1e3:.0f} GB/s")
    yy = bf(2, T, di); yn = bf(T, 2 * di)
    us = timeit(lambda: ops.gated_norm_fwd(yy, zx, ds, lens, nw, 2, B, L, di)); by = 2 * T * di * 6
    print(f"{tag} gated_norm_fwd {us:.1f} us  {by / us / 1e3:.0f} GB/s")
    yn2, rs = ops.gated_norm_fwd(yy, zx, ds, lens, nw, 2, B, L, di)
    us = timeit(lambda: ops.gated_norm_bwd(yn, yy, zx, ds, lens, nw, rs, 2, B, L, di, dzx)); by = 2 * T * di * 10
    print(f"{tag} gated_norm_bwd {us:.1f} us  {by / us / 1e3:.0f} GB/s")
    g, b_ = torch.randn(d, device=DEV), torch.randn(d, device=DEV)
    us = timeit(lambda: ops.layernorm_fwd(x, g, b_, 1e-5, torch.bfloat16)); by = T * d * 4
    print(f"{tag} layernorm_fwd {us:.1f} us  {by / us / 1e3:.0f} GB/s")
    y_, mean, rstd = ops.layernorm_fwd(x, g, b_, 1e-5, torch.bfloat16)
    us = timeit(lambda: ops.layernorm_bwd(h, x, g, mean, rstd, x)); by = T * d * 8
    print(f"{tag} layernorm_bwd {us:.1f} us  {by / us / 1e3:.0f} GB/s")
