"""Small bf16-autocast encoder step (Type A N=2, two ragged utterances, L not a multiple of 128) for compute-sanitizer."""
import sys
sys.path.insert(0, "tests"); import _util
import torch, dcasr_b200 as dd
torch.manual_seed(0)
dev = "cuda"
enc = dd.DCASREncoder(n_mels=80, d_outer=128, d_main=128, n_enc=1, n_main=1, n_dec=1, arch_type="A", N=2).to(dev)
with torch.no_grad():
    enc.chunk.router.W_k.weight.copy_(torch.randn(128, 128, device=dev) / 128 ** 0.5)
feats = torch.randn(2, 1230, 80, device=dev); lens = torch.tensor([1230, 901], device=dev)
with torch.autocast("cuda", dtype=torch.bfloat16):
    o = enc(feats, lens)
(o.features.float().pow(2).mean() + 0.03 * o.ratio_loss).backward()
torch.cuda.synchronize()
print("ok", float(o.kept_fractions[0]), all(p.grad is not None and torch.isfinite(p.grad).all() for p in enc.parameters()))
