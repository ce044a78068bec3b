import sys, os
sys.path.insert(0, "tests"); import _util
from _util import fill_weights, rel_err
import torch, dcasr_b200 as dd
from oracle.encoder_ref import EncoderRef
DEV = "cuda"
for mode in ("fp32", "bf16"):
    kw = dict(n_mels=80, d_outer=128, d_main=256, n_enc=2, n_main=2, n_dec=2, n_mid=1, arch_type="A", N=2)
    ref = EncoderRef(**kw); fill_weights(ref, 77, router_identity=True)
    enc = dd.DCASREncoder(**kw); enc.load_state_dict(ref.state_dict()); enc = enc.to(DEV)
    torch.manual_seed(3)
    B, L = 3, 180
    lengths = torch.tensor([180, 131, 64])
    x = torch.randn(B, L, 128); x = x + 1.5 * torch.roll(x, 1, 1) * (torch.rand(B, L, 1) > 0.5)
    xr = x.clone().requires_grad_(True)
    o_ref = ref.forward_from_subsampled(xr, lengths)
    xg = x.to(DEV).requires_grad_(True)
    if mode == "bf16":
        with torch.autocast("cuda", dtype=torch.bfloat16):
            o = enc.forward_hot_path(xg, lengths.to(DEV))
    else:
        o = enc.forward_hot_path(xg, lengths.to(DEV))
    mask = (torch.arange(L)[None] < lengths[:, None]).unsqueeze(-1)
    print(mode, "b equal:", torch.equal(o.boundaries[0][1].cpu(), o_ref.boundaries[0][1]), "feat err", rel_err(o.features.cpu()*mask, o_ref.features*mask))
    w = torch.randn(B, L, 128)
    ((o.features * (w * mask).to(DEV)).sum() + 0.03 * o.ratio_loss).backward()
    ((o_ref.features * w * mask).sum() + 0.03 * o_ref.ratio_loss).backward()
    print(" dx", rel_err(xg.grad, xr.grad))
    gs, gr = dict(enc.named_parameters()), dict(ref.named_parameters())
    rows = sorted(((rel_err(gs[k].grad, gr[k].grad), k, gr[k].grad.numel()) for k in gr if gr[k].grad is not None), reverse=True)
    for e, k, n in rows[:14]: print(f"  {e:.3e} {k} ({n})")
    import collections
    by = collections.defaultdict(list)
    for e, k, n in rows: by[k.split(".")[-2] + "." + k.split(".")[-1]].append(e)
    print("  median by class:", {k: f"{sorted(v)[len(v)//2]:.2e}" for k, v in by.items()})
