"""One launch of each bandwidth kernel at the outer-stack shape: ncu target."""
import sys
sys.path.insert(0, "tests"); import _util
import torch
from dcasr_b200 import ops
DEV = "cuda"
def bf(*shape): return (torch.randn(*shape, device=DEV) * 0.5).to(torch.bfloat16)
T, d, di, H, N = 15920, 384, 768, 12, 128
C = di + 2 * N; dip = 2 * di + 2 * N + H; ds = (dip + 7) // 8 * 8; ldz = 2 * ds; B = 40; L = T // B
zx = bf(T, ldz); x = bf(T, d); h = bf(T, d)
cw, cb = torch.randn(2, C, 4, device=DEV), torch.randn(2, C, device=DEV)
dtb, nw = torch.randn(2, H, device=DEV), torch.randn(2, di, device=DEV)
lens = torch.full((B,), L, dtype=torch.int32, device=DEV)
dxc, dBC, ddt = bf(2, T, di), bf(2, T, 2 * N), torch.randn(2, T, H, device=DEV)
dzx = torch.zeros_like(zx); yy = bf(2, T, di); yn = bf(T, 2 * di)
g, b_ = torch.randn(d, device=DEV), torch.randn(d, device=DEV)
for _ in range(2):
    ops.conv_fwd(zx, ds, lens, cw, cb, dtb, 2, B, L, di, N, H)
    ops.conv_bwd(zx, dxc, dBC, ddt, ds, lens, cw, cb, dtb, 2, B, L, di, N, H, dzx)
    yn2, rs = ops.gated_norm_fwd(yy, zx, ds, lens, nw, 2, B, L, di)
    ops.gated_norm_bwd(yn, yy, zx, ds, lens, nw, rs, 2, B, L, di, dzx)
    y_, mean, rstd = ops.layernorm_fwd(x, g, b_, 1e-5, torch.bfloat16)
    ops.layernorm_bwd(h, x, g, mean, rstd, x)
torch.cuda.synchronize(); print("ok")
