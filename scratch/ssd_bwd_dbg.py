import sys, os
sys.path.insert(0, "tests"); import _util
import torch, torch.nn.functional as F
from dcasr_b200 import ops
DEV = "cuda"; torch.manual_seed(0)
for (B, L, H) in ((40, 398, 12), (40, 196, 16)):
    ndir = 2; di, N = 64 * H, 128
    xconv = (torch.randn(ndir, B * L, di + 2 * N, device=DEV) * 0.8).to(torch.bfloat16)
    dt = F.softplus(torch.randn(ndir, B * L, H, device=DEV) - 2.0)
    A_log = torch.log(torch.rand(ndir, H, device=DEV) * 15 + 1); Dk = torch.randn(ndir, H, device=DEV)
    dy = (torch.randn(ndir, B * L, di, device=DEV) * 0.5).to(torch.bfloat16)
    os.environ.pop("HNB_SSD_DEBUG", None)
    y, ws = ops.ssd_fwd(xconv, dt, A_log, Dk, ndir, B, L, di, N, H, impl=1)
    os.environ["HNB_SSD_DEBUG"] = "1"
    for _ in range(2):
        ops.ssd_bwd(dy, xconv, y, dt, A_log, Dk, ws, ndir, B, L, di, N, H, impl=1, keep_parts=True)
    torch.cuda.synchronize()
