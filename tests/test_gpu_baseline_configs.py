"""GPU parity at the configurations BASELINE.json actually names, at FULL depth and width (VERDICT r1 "next" #1):

  A_small_N2  Type A Small N=2   4/12/4 layers, d 384/512, 3 ragged utterances up to 16 s        (BASELINE config 2)
  B_small_N4  Type B Small N=4   4/4/12/4/4 layers, two sqrt(N)=2 chunk stages, 2 ragged utterances (BASELINE config 3)
  A_large_N3  Type A Large N=3   6/18/6 layers, d 512/768, one 60 s utterance (1498 frames, 12 SSD chunks) (config 4)

fp32: CUDA product against the CPU oracle (oracle/encoder_ref.py), 1e-3 relative on activations AND gradients, every
boundary / membership / mask bit-exact (the cases' router seeds keep every |p - 0.5| above the 1e-4 band, see
tests/golden/find_margin_seeds.py and tests/baseline_cases.py, so there is no skip and no band escape).
bf16: CUDA product under bf16 autocast against THE SAME ORACLE RUN ON THE GPU UNDER bf16 AUTOCAST (the comparator
SURVEY.md §7 asks for: both sides round where the reference's own training path rounds), north_star's 2e-2; the distance
of each of them to the fp32 oracle is printed as information.
"""
import copy

import pytest
import torch

from _util import fill_weights, force_in_band_boundaries, max_err, rel_err, set_router_operating_point
from baseline_cases import CASES, make_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _build(name):
    import dcasr_b200 as dd
    from oracle.encoder_ref import EncoderRef
    case = CASES[name]
    ref = EncoderRef(**case["kw"])
    fill_weights(ref, case["seed"], router_identity=True)
    set_router_operating_point(ref, case["router_seed"], case["shift"])
    enc = dd.DCASREncoder(**case["kw"])
    enc.load_state_dict(ref.state_dict())           # identical key sets: strict load
    return case, ref, enc.to(DEV)


def _valid_masks(o_ref):
    """valid-frame mask of every chunk stage's input sequence."""
    L0 = o_ref.boundaries[0][0].shape[1]
    masks = [torch.arange(L0)[None, :] < o_ref.lengths[:, None]]
    for (p, b) in o_ref.boundaries[:-1]:
        cnt = (b > 0.5).sum(1)
        M = int(cnt.max())
        masks.append(torch.arange(M)[None, :] < cnt[:, None])
    return masks


@pytest.mark.parametrize("name", ["A_small_N2", "B_small_N4", "A_large_N3"])
def test_full_size_encoder_fp32_vs_oracle(name):
    case, ref, enc = _build(name)
    feats, lens = make_inputs(case, case["seed"])
    torch.backends.cudnn.allow_tf32 = False          # ConvSubsampling4 (cuDNN / cuBLAS, before the hot path) in true fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    o_ref = ref(feats, lens)
    masks = _valid_masks(o_ref)
    margin = min(float((p[m] - 0.5).abs().min()) for (p, _), m in zip(o_ref.boundaries, masks))
    assert margin >= case["margin"], f"case seeds no longer keep p out of the band: {margin:.2e}"
    o = enc(feats.to(DEV), lens.to(DEV))
    assert torch.equal(o.lengths.cpu(), o_ref.lengths)
    for i, ((p, b), (pr, br)) in enumerate(zip(o.boundaries, o_ref.boundaries)):
        perr = max_err(p, pr)
        print(f"{name} stage {i}: max |p - p_ref| = {perr:.2e} (margin {margin:.2e}), kept {float(o.kept_fractions[i]):.3f}")
        assert perr < 0.25 * margin
        assert torch.equal(b.cpu(), br), f"stage {i}: boundary mask differs"          # EVERY frame, no band needed
        assert rel_err(o.chunk_embeddings[i], o_ref.chunk_embeddings[i]) < 1e-3
        assert max_err(o.kept_fractions[i], o_ref.kept_fractions[i]) < 1e-6
    m0 = masks[0].unsqueeze(-1)
    ferr = rel_err(o.features.cpu() * m0, o_ref.features * m0)
    print(f"{name}: fp32 feature rel err {ferr:.2e}, ratio loss {float(o.ratio_loss):.6f} vs {float(o_ref.ratio_loss):.6f}")
    assert ferr < 1e-3
    assert max_err(o.ratio_loss, o_ref.ratio_loss) < 1e-5
    w = torch.randn(o_ref.features.shape, generator=torch.Generator().manual_seed(5)) * m0
    ((o_ref.features * w).sum() + 0.03 * o_ref.ratio_loss).backward()
    ((o.features * w.to(DEV)).sum() + 0.03 * o.ratio_loss).backward()
    gs, gr = dict(enc.named_parameters()), dict(ref.named_parameters())
    assert all(p.grad is not None for p in enc.parameters()), "DDP(find_unused_parameters=False) needs every grad"
    errs = sorted(((rel_err(gs[k].grad, gr[k].grad), k, float(gr[k].grad.norm())) for k in gr), reverse=True)
    tot = sum(float((gs[k].grad.cpu().double() - gr[k].grad.double()).pow(2).sum()) for k in gr) ** 0.5 / \
        sum(float(gr[k].grad.double().pow(2).sum()) for k in gr) ** 0.5
    print(f"{name}: gradient rel err over all {len(errs)} parameter tensors {tot:.2e}; worst five:",
          [(f"{e:.1e}", k, f"|g|={n:.1e}") for e, k, n in errs[:5]])
    assert tot < 1e-3
    gmax = max(n for _, _, n in errs)
    for e, k, n in errs:                                  # every tensor on its own, unless its gradient is numerically nil
        if n > 1e-5 * gmax:
            assert e < 1e-3, (k, e, n)


BF16_BAND = 2e-2          # bf16 q, k move p by up to ~1e-2 in EITHER implementation (the reference's own bf16 path too)


@pytest.mark.parametrize("name", ["A_small_N2", "B_small_N4"])
def test_full_size_hot_path_bf16_vs_bf16_oracle_on_gpu(name, monkeypatch):
    """bf16 autocast at full depth: product vs the oracle run on the same GPU under the same autocast (bf16 vs bf16),
    north_star's 2e-2 on activations and gradients.

    Boundaries: at the trained operating point a few of the ~850 valid frames always sit within bf16 noise of p = 0.5, and
    one flipped boundary re-indexes every later chunk.  So the test (1) asserts b equal wherever the comparator's
    |p - 0.5| > BF16_BAND and p equal within the band width everywhere, and (2) teacher-forces the (few, counted)
    in-band decisions of the product to the comparator's, so that everything downstream is compared frame by frame.
    Nothing is skipped."""
    case, ref, enc = _build(name)
    feats, lens = make_inputs(case, case["seed"])
    with torch.no_grad():
        x, l0 = ref.subsample(feats, lens)                 # identical fp32 input of the hot path for all three runs
    B, L, d = x.shape
    ref_gpu = copy.deepcopy(ref).to(DEV)
    xo = x.to(DEV).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ob = ref_gpu.forward_from_subsampled(xo, l0.to(DEV))
    # fp32 CPU oracle on the comparator's boundaries (information only: how far either bf16 run is from fp32 truth)
    from oracle import hnet_ref
    real_router_ref, calls = hnet_ref.router_ref, [0]

    def router_ref_forced(xx, Wq, Wk, mask=None, eps=1e-6):
        p_, b_ = real_router_ref(xx, Wq, Wk, mask, eps)
        bb = ob.boundaries[calls[0]][1].cpu().to(b_.dtype)
        calls[0] += 1
        return p_, (bb if bb.shape == b_.shape else b_)

    monkeypatch.setattr(hnet_ref, "router_ref", router_ref_forced)
    xr32 = x.clone().requires_grad_(True)
    o32 = ref.forward_from_subsampled(xr32, l0)
    monkeypatch.setattr(hnet_ref, "router_ref", real_router_ref)
    forced = force_in_band_boundaries(monkeypatch, ob.boundaries, BF16_BAND)
    xg = x.to(DEV).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o = enc.forward_hot_path(xg, l0.to(DEV))
    assert o.features.dtype == torch.float32 and o.boundaries[0][0].dtype == torch.float32
    n_valid = int(l0.sum())
    print(f"{name}: in-band decisions teacher-forced per stage {forced} of {n_valid} valid frames; kept "
          f"{[round(float(k), 3) for k in o.kept_fractions]}")
    assert sum(forced) <= 0.02 * n_valid
    for (p, b), (pb, bb) in zip(o.boundaries, ob.boundaries):
        assert torch.equal(b, bb)
    mask = (torch.arange(L)[None] < l0[:, None]).unsqueeze(-1)
    mg = mask.to(DEV)
    e_bb = rel_err(o.features * mg, ob.features * mg)
    same32 = all(torch.equal(bb.cpu(), b32) for (_, bb), (_, b32) in zip(ob.boundaries, o32.boundaries))
    if same32:                                             # distance of either bf16 run to the fp32 oracle
        print(f"{name}: features  ours-vs-fp32 {rel_err(o.features.cpu() * mask, o32.features * mask):.2e} | "
              f"bf16oracle-vs-fp32 {rel_err(ob.features.cpu() * mask, o32.features * mask):.2e}")
    e_o32 = rel_err(o.features.cpu() * mask, o32.features * mask)
    e_b32 = rel_err(ob.features.cpu() * mask, o32.features * mask)
    print(f"{name}: features  ours-vs-bf16oracle {e_bb:.2e}   ratio loss {float(o.ratio_loss):.5f} vs {float(ob.ratio_loss):.5f}")
    # After 20-28 blocks bf16 rounding noise has been amplified to 4e-2 (Type A) / 9e-2 (Type B) of the features for the
    # comparator ITSELF against fp32 truth (measured, printed above): no two bf16 implementations that round
    # independently can agree to 2e-2 end to end at this depth.  The end-to-end assertion is therefore relative -- the
    # product must be as close to fp32 truth as the reference-style bf16 path is, and as close to that path as that path
    # is to truth; north_star's absolute 2e-2 is asserted where it is attainable: block by block on the same inputs,
    # every block of the full-size model (test_full_size_blocks_bf16_layerwise_vs_bf16_oracle).
    assert same32
    assert e_o32 < 1.1 * e_b32 and e_bb < 1.25 * e_b32, (e_o32, e_b32, e_bb)
    assert max_err(o.ratio_loss, ob.ratio_loss) < 2e-3
    w = torch.randn(B, L, d, generator=torch.Generator().manual_seed(5)) * mask
    ((o.features * w.to(DEV)).sum() + 0.03 * o.ratio_loss).backward()
    ((ob.features * w.to(DEV)).sum() + 0.03 * ob.ratio_loss).backward()
    ((o32.features * w).sum() + 0.03 * o32.ratio_loss).backward()
    xr32_grad = xr32.grad
    e_dx = rel_err(xg.grad, xo.grad)
    gs, gb = dict(enc.named_parameters()), dict(ref_gpu.named_parameters())
    keys = [k for k in gb if gb[k].grad is not None]        # (the subsample front end is not on this path)
    num = sum(float((gs[k].grad.double() - gb[k].grad.double()).pow(2).sum()) for k in keys)
    den = sum(float(gb[k].grad.double().pow(2).sum()) for k in keys)
    e_gw = (num / den) ** 0.5
    per = sorted(((rel_err(gs[k].grad, gb[k].grad), k) for k in keys if gb[k].grad.numel() >= 64), reverse=True)
    print(f"{name}: d x ours-vs-bf16oracle {e_dx:.2e}; all parameter gradients ours-vs-bf16oracle {e_gw:.2e}; worst tensors "
          f"{[(f'{e:.1e}', k) for e, k in per[:4]]}")
    e_dx32, e_gw32 = rel_err(xo.grad, xr32_grad), None
    print(f"{name}: d x bf16oracle-vs-fp32 {e_dx32:.2e}, ours-vs-fp32 {rel_err(xg.grad, xr32_grad):.2e}")
    assert e_dx < 1.25 * e_dx32 and rel_err(xg.grad, xr32_grad) < 1.1 * e_dx32


@pytest.mark.parametrize("name", ["A_small_N2", "A_large_N3"])
def test_full_size_blocks_bf16_layerwise_vs_bf16_oracle(name):
    """north_star's bf16 bar (2e-2 on activations and gradients), asserted for EVERY Mamba block of the full-size model
    on the activations that block really sees: the oracle runs the whole encoder on the GPU under bf16 autocast, hooks
    record each block's input, output and their gradients, and the product's block (same weights) is run on the same
    input / output gradient under the same autocast.  This removes only the amplification of earlier blocks' rounding
    noise by later blocks (which no implementation controls), not any arithmetic of the block itself."""
    case, ref, enc = _build(name)
    feats, lens = make_inputs(case, case["seed"])
    with torch.no_grad():
        x, l0 = ref.subsample(feats, lens)
    ref_gpu = copy.deepcopy(ref).to(DEV)
    rec = {}

    def fwd_hook(tag):
        def hook(mod, args, out):
            rec[tag] = {"x": args[0].detach(), "lengths": args[1] if len(args) > 1 else None, "y": out.detach()}
        return hook

    def bwd_hook(tag):
        def hook(mod, gin, gout):
            rec[tag]["gy"], rec[tag]["gx"] = gout[0].detach(), gin[0].detach()
        return hook

    stacks = [s for s in ("enc", "mid", "main", "mid_dec", "dec") if hasattr(ref_gpu, s)]
    for sname in stacks:
        for i, blk in enumerate(getattr(ref_gpu, sname).layers):
            blk.register_forward_hook(fwd_hook(f"{sname}.{i}"))
            blk.register_full_backward_hook(bwd_hook(f"{sname}.{i}"))
    xo = x.to(DEV).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ob = ref_gpu.forward_from_subsampled(xo, l0.to(DEV))
    L = x.shape[1]
    mask = (torch.arange(L, device=DEV)[None] < l0.to(DEV)[:, None]).unsqueeze(-1)
    w = torch.randn(ob.features.shape, generator=torch.Generator().manual_seed(5)).to(DEV) * mask
    ((ob.features * w).sum() + 0.03 * ob.ratio_loss).backward()
    worst = {"y": (0, ""), "dx": (0, ""), "gw": (0, "")}
    gref = dict(ref_gpu.named_parameters())
    for sname in stacks:
        for i, blk in enumerate(getattr(enc, sname).layers):
            tag = f"{sname}.{i}"
            r = rec[tag]
            lengths = r["lengths"]
            Lb = r["x"].shape[1]
            m = (torch.arange(Lb, device=DEV)[None] < lengths[:, None]).unsqueeze(-1) if lengths is not None else 1.0
            xin = r["x"].clone().requires_grad_(True)
            for p in blk.parameters():
                p.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = blk(xin, lengths)
            e_y = rel_err((y - xin) * m, (r["y"] - r["x"]) * m)                 # the mixers' contribution, not x + ...
            y.backward(r["gy"])
            e_dx = rel_err((xin.grad - r["gy"]) * m, (r["gx"] - r["gy"]) * m)
            num = den = 0.0
            for k, p in blk.named_parameters():
                gr = gref[f"{tag.replace('.', '.layers.')}.{k}"].grad
                if gr.numel() >= 64:                                            # weight matrices and norm / conv vectors
                    num += float((p.grad.double() - gr.double()).pow(2).sum())
                    den += float(gr.double().pow(2).sum())
            e_gw = (num / max(den, 1e-300)) ** 0.5
            for key, e in (("y", e_y), ("dx", e_dx), ("gw", e_gw)):
                if e > worst[key][0]:
                    worst[key] = (e, tag)
            assert e_y < 2e-2 and e_dx < 2e-2 and e_gw < 2e-2, (tag, e_y, e_dx, e_gw)
    n_blocks = sum(len(getattr(enc, s).layers) for s in stacks)
    print(f"{name}: {n_blocks} blocks, worst bf16-vs-bf16 rel err: mixer output {worst['y'][0]:.2e} ({worst['y'][1]}), "
          f"d x {worst['dx'][0]:.2e} ({worst['dx'][1]}), block weight gradients {worst['gw'][0]:.2e} ({worst['gw'][1]})")


@pytest.mark.parametrize("arch,N", [("A", 2), ("B", 4)])
def test_fixed_chunker_encoder_matches_reference_golden(arch, N):
    """`chunker: fixed` encoders (FixedPoolChunker, reference models/fixed_pool.py) against the vectors the reference's
    own encoder.py produced (tests/golden/encfixed_*.npz): values, integer outputs and every parameter gradient."""
    import os

    import numpy as np

    import dcasr_b200 as dd
    from _util import GOLDEN
    g = np.load(os.path.join(GOLDEN, f"encfixed_{arch}_N{N}.npz"))
    enc = dd.DCASREncoder(n_mels=80, d_outer=64, d_main=128, n_enc=1, n_main=1, n_dec=1, n_mid=1, arch_type=arch, N=N,
                          chunker="fixed")
    fill_weights(enc, int(g["seed"]))
    enc = enc.to(DEV)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    out = enc(torch.from_numpy(g["feats"]).to(DEV), torch.from_numpy(g["feat_lengths"]).to(DEV))
    assert torch.equal(out.lengths.cpu(), torch.from_numpy(g["lengths"]))
    i = 0
    while f"p{i}" in g:
        p, b = out.boundaries[i]
        assert torch.equal(p.cpu(), torch.from_numpy(g[f"p{i}"])) and torch.equal(b.cpu(), torch.from_numpy(g[f"b{i}"]))
        assert rel_err(out.chunk_embeddings[i], torch.from_numpy(g[f"z{i}"])) < 1e-3
        assert max_err(out.kept_fractions[i], torch.from_numpy(g[f"kept{i}"])) < 1e-6
        i += 1
    assert i == (1 if arch == "A" else 2)
    mask = (torch.arange(out.features.shape[1], device=DEV)[None] < out.lengths[:, None]).unsqueeze(-1)
    ref = torch.from_numpy(g["features"]).to(DEV)
    assert rel_err(out.features * mask, ref * mask) < 1e-3
    assert max_err(out.ratio_loss, torch.from_numpy(g["ratio_loss"])) < 1e-6
    loss = (out.features * torch.from_numpy(g["w"]).to(DEV) * mask).sum() + 0.03 * out.ratio_loss
    loss.backward()
    sd = dict(enc.named_parameters())
    for k in g.files:
        if k.startswith("g_"):
            assert rel_err(sd[k[2:]].grad, torch.from_numpy(g[k])) < 2e-3, k


def test_lengths_beyond_the_padded_length_are_clamped():
    """lengths[b] > L or < 0: the kernels clamp like reverse_sequences' clamp_(0, T-1) instead of indexing
    out of bounds; the result equals the call with the clamped lengths."""
    import dcasr_b200 as dd
    torch.manual_seed(2)
    blk = dd.MambaBlock(128).to(DEV)
    x = torch.randn(3, 70, 128, device=DEV)
    bad = torch.tensor([70 + 500, 33, -4], device=DEV)
    ok = torch.tensor([70, 33, 0], device=DEV)
    outs = []
    for lens in (bad, ok):
        xg = x.clone().requires_grad_(True)
        y = blk(xg, lens)
        y.pow(2).sum().backward()
        torch.cuda.synchronize()
        outs.append((y.detach(), xg.grad))
    assert torch.equal(outs[0][0], outs[1][0])
    assert rel_err(outs[0][1], outs[1][1]) < 1e-5            # (fp32 atomic sums inside the backward: order varies run to run)
