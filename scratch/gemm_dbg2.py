import sys, os
sys.path.insert(0, "tests"); import _util
import torch
from dcasr_b200 import ops
DEV = "cuda"
def bf(*s): return (torch.randn(*s, device=DEV) * 0.5).to(torch.bfloat16)
T, d, ldz, di2 = 15920, 384, 3616, 1536
h, Win, yn, Wout, x, zx = bf(T, d), bf(ldz, d), bf(T, di2), bf(d, di2), bf(T, d), bf(T, ldz)
for name, fn in (("in_proj fwd", lambda: ops.gemm(h, Win)), ("out_proj fwd", lambda: ops.gemm(yn, Wout, residual=x)),
                 ("out_proj dgrad", lambda: ops.gemm(x, Wout, trans_b=True)), ("in_proj dgrad", lambda: ops.gemm(zx, Win, trans_b=True))):
    print("==", name, file=sys.stderr, flush=True)
    for _ in range(2): fn()
    torch.cuda.synchronize()
