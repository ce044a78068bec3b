"""Round-2 kernels under compute-sanitizer: split SSD forward (impl 5) at a ragged, multi-chunk shape, the fp32 tensor-core
GEMM (all layouts, ragged K), the CTC head (ragged lengths, empty / infeasible targets) and a small fp32 + bf16 encoder step."""
import sys
sys.path.insert(0, "tests"); import _util
import torch, torch.nn.functional as F, dcasr_b200 as dd
from dcasr_b200 import ops
torch.manual_seed(0)
dev = "cuda"
# split SSD forward: 10 chunks (>= 8: also what impl 1 picks), partial last chunk, 5 heads (odd head grouping)
ndir, B, L, H = 2, 2, 1200, 5
di, N = 64 * H, 128
xconv = (torch.randn(ndir, B * L, di + 2 * N, device=dev) * 0.8).to(torch.bfloat16)
dt = F.softplus(torch.randn(ndir, B * L, H, device=dev) - 2.0)
A_log = torch.log(torch.rand(ndir, H, device=dev) * 15 + 1); Dk = torch.randn(ndir, H, device=dev)
y5, ws5 = ops.ssd_fwd(xconv, dt, A_log, Dk, ndir, B, L, di, N, H, impl=5)
y4, ws4 = ops.ssd_fwd(xconv, dt, A_log, Dk, ndir, B, L, di, N, H, impl=4)
dy = (torch.randn(ndir, B * L, di, device=dev) * 0.5).to(torch.bfloat16)
ops.ssd_bwd(dy, xconv, y5, dt, A_log, Dk, ws5, ndir, B, L, di, N, H, impl=1)
print("ssd split vs persistent", float((y5.float() - y4.float()).abs().max()))
# fp32 tensor-core GEMM
for ta, tb in ((0, 0), (0, 1), (1, 0), (1, 1)):
    M, Nn, K = 520, 131, 1001
    a = torch.randn((K, M) if ta else (M, K), device=dev); b = torch.randn((K, Nn) if tb else (Nn, K), device=dev)
    c = ops.gemm(a, b, trans_a=bool(ta), trans_b=bool(tb))
print("gemm_f32_tc", float(c.abs().mean()))
# CTC head
head = dd.CTCHead(32, 20).to(dev)
x = torch.randn(4, 37, 32, device=dev, requires_grad=True)
fl = torch.tensor([37, 30, 4, 0], device=dev); tl = torch.tensor([5, 0, 8, 3], device=dev)
tg = torch.randint(0, 20, (4, 8), device=dev)
for red in ("mean", "sum"):
    loss = head.loss(x, fl, tg, tl, reduction=red); loss.backward()
with torch.autocast("cuda", dtype=torch.bfloat16):
    head.loss(x, fl, tg, tl).backward()
print("ctc", float(loss), head.greedy_decode(x, fl)[0][:4])
# encoder, fp32 (exact / tensor-core fp32 GEMMs) and bf16
enc = dd.DCASREncoder(n_mels=80, d_outer=128, d_main=128, n_enc=1, n_main=1, n_dec=1, arch_type="B", N=4).to(dev)
feats = torch.randn(2, 1230, 80, device=dev); lens = torch.tensor([1230, 901], device=dev)
o = enc(feats, lens); (o.features.pow(2).mean() + 0.03 * o.ratio_loss).backward()
with torch.autocast("cuda", dtype=torch.bfloat16):
    o = enc(feats, lens)
(o.features.float().pow(2).mean() + 0.03 * o.ratio_loss).backward()
torch.cuda.synchronize()
print("ok", [float(k) for k in o.kept_fractions])
