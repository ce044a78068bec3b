"""ConvSubsampling4 forward + backward at the headline shape (40 x 16 s, bf16 autocast): target for ncu."""
import sys
sys.path.insert(0, "tests"); import _util
import torch, dcasr_b200 as dd
torch.manual_seed(0)
dev = "cuda"
sub = dd.ConvSubsampling4(80, 384).to(dev)
feats = torch.randn(40, 1598, 80, device=dev); lens = torch.full((40,), 1598, device=dev)
for it in range(3):
    if it == 2:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    for p in sub.parameters(): p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y, l = sub(feats, lens)
    y.float().pow(2).mean().backward()
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("ok", tuple(y.shape))
