import sys, time
sys.path.insert(0, "tests"); import _util
import torch, dcasr_b200 as dd
dev = "cuda"
for fmt in ("contiguous", "channels_last"):
    torch.manual_seed(0)
    sub = dd.ConvSubsampling4(80, 384).to(dev)
    if fmt == "channels_last":
        sub = sub.to(memory_format=torch.channels_last)
    feats = torch.randn(40, 1598, 80, device=dev); lens = torch.full((40,), 1598, device=dev)
    def step():
        for p in sub.parameters(): p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            f = feats.unsqueeze(1)
            if fmt == "channels_last": f = f.contiguous(memory_format=torch.channels_last)
            x = sub.conv(f)
            B, C, T, F = x.shape
            y = sub.proj(x.transpose(1, 2).reshape(B, T, C * F))
        y.float().pow(2).mean().backward()
    for _ in range(3): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): step()
    torch.cuda.synchronize(); print(fmt, "subsample fwd+bwd ms:", (time.perf_counter() - t0) * 100)
