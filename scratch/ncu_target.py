"""One bidirectional MambaBlock fwd+bwd at the headline shape (B=40 x 398 frames, d=384, bf16 autocast) plus one
chunk/dechunk round trip: every hand-written kernel of the hot path is launched once.  Target for ncu."""
import sys
sys.path.insert(0, "tests"); import _util
import torch, dcasr_b200 as dd
torch.manual_seed(0)
dev = "cuda"
B, L, d = 40, 398, 384
blk = dd.MambaBlock(d).to(dev)
ch = dd.DynamicChunker(d, N=2).to(dev)
with torch.no_grad():
    ch.router.W_k.weight.copy_(torch.randn(d, d, device=dev) / d ** 0.5)
x = torch.randn(B, L, d, device=dev, requires_grad=True)
lens = torch.full((B,), L, device=dev, dtype=torch.int32)
for it in range(3):
    if it == 2:
        torch.cuda.synchronize(); torch.cuda.profiler.start()      # ncu --profile-from-start off: third iteration only
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = blk(x, lens)
        co = ch.chunk(y.float())
        z = ch.dechunk(co.z.to(torch.bfloat16), co, residual=y.float())
    (z.float().pow(2).mean() + 0.03 * co.ratio_loss).backward()
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("ok", float(co.kept_fraction))
