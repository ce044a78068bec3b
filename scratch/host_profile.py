"""Host-side cost of one encoder step: tiny batch (GPU time negligible), cProfile over 10 steps."""
import sys, cProfile, pstats, io, time
sys.path.insert(0, "tests"); import _util
import torch, dcasr_b200 as dd
torch.manual_seed(1)
dev = "cuda"
enc = dd.DCASREncoder(n_mels=80, d_outer=384, d_main=512, n_enc=4, n_main=12, n_dec=4, arch_type="A", N=2).to(dev)
with torch.no_grad():
    g = torch.Generator().manual_seed(7)
    enc.chunk.router.W_k.weight.copy_(torch.randn(384, 384, generator=g) / 384 ** 0.5)
feats = torch.randn(2, 198, 80, device=dev); lens = torch.tensor([198, 198], device=dev)
params = list(enc.parameters())
def step():
    for p in params: p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = enc(feats, lens)
    loss = out.features.float().pow(2).mean() + 0.03 * out.ratio_loss
    loss.backward()
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): step()
torch.cuda.synchronize()
print("ms/step (host-bound):", (time.perf_counter() - t0) * 100)
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(30); print(s.getvalue()[:6000])
