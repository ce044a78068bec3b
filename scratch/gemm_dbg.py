"""The two GEMM shapes that dominate the step (outer in_proj fwd, K = 384; main in_proj dgrad, K = 4640): ncu target."""
import sys
sys.path.insert(0, "tests"); import _util
import torch
from dcasr_b200 import ops
DEV = "cuda"
def bf(*shape): return (torch.randn(*shape, device=DEV) * 0.5).to(torch.bfloat16)
h, Win = bf(15920, 384), bf(3616, 384)
zx, W2 = bf(7840, 4640), bf(4640, 512)
for _ in range(3):
    ops.gemm(h, Win)
    ops.gemm(zx, W2, trans_b=True)
torch.cuda.synchronize()
print("ok")
