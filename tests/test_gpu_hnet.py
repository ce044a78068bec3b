"""GPU parity of the H-Net stage kernels (through the C ABI via dcasr_b200) against
(a) golden vectors produced by the reference's own hnet_chunk.py, (b) the CPU oracle on seeded inputs,
(c) size-independent properties at BASELINE sizes.  Tolerances: integer/boolean outputs bit-exact
wherever |p - 0.5| > 1e-4 (north_star); fp32 activations/gradients 1e-3 relative (we assert tighter)."""
import glob
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN, max_err, rel_err

pytestmark = pytest.mark.gpu
HNET = sorted(glob.glob(os.path.join(GOLDEN, "hnet_*.npz")))
EMA = sorted(glob.glob(os.path.join(GOLDEN, "ema_*.npz")))
DEV = "cuda"


def _t(a, grad=False):
    t = torch.from_numpy(np.asarray(a)).to(DEV)
    return t.requires_grad_(True) if grad else t


def _chunker(g):
    import dcasr_b200 as d
    N = float(g["N"])
    N = int(N) if N == int(N) else N
    ch = d.DynamicChunker(g["x"].shape[-1], N=N, ema_smoothing=bool(g["ema"])).to(DEV)
    with torch.no_grad():
        ch.router.W_q.weight.copy_(_t(g["Wq"]))
        ch.router.W_k.weight.copy_(_t(g["Wk"]))
    return ch


@pytest.mark.parametrize("path", HNET, ids=[os.path.basename(p)[:-4] for p in HNET])
def test_chunk_dechunk_matches_reference_golden(path):
    g = np.load(path)
    ch = _chunker(g)
    x = _t(g["x"], True)
    mask = _t(g["mask"]) if "mask" in g else None
    co = ch.chunk(x, mask)
    p_ref, b_ref = _t(g["p"]), _t(g["b"])
    assert max_err(co.p, p_ref) < 2e-6
    safe = (p_ref - 0.5).abs() > 1e-4
    assert torch.equal(co.b[safe], b_ref[safe])
    assert torch.equal(co.b, b_ref), "a boundary inside the 1e-4 band flipped; golden case needs a wider margin"
    assert co.membership.dtype == torch.int64 and torch.equal(co.membership, _t(g["membership"]))
    assert co.z_mask.dtype == torch.bool and torch.equal(co.z_mask, _t(g["z_mask"]))
    assert torch.equal(co.z, _t(g["z"]))
    assert max_err(co.ratio_loss, _t(g["ratio_loss"])) < 1e-6
    assert max_err(co.kept_fraction, _t(g["kept_fraction"])) < 1e-7
    z_proc = _t(g["z_proc"], True)
    y = ch.dechunk(z_proc, co)
    assert rel_err(y, _t(g["y"])) < 1e-5
    loss = (y * _t(g["w"])).sum() + (co.z * _t(g["wz"])).sum() + 0.03 * co.ratio_loss
    assert rel_err(loss, _t(g["loss"])) < 1e-5
    loss.backward()
    assert rel_err(x.grad, _t(g["gx"])) < 1e-4
    assert rel_err(z_proc.grad, _t(g["gz"])) < 1e-4
    assert rel_err(ch.router.W_q.weight.grad, _t(g["gWq"])) < 1e-4
    assert rel_err(ch.router.W_k.weight.grad, _t(g["gWk"])) < 1e-4


@pytest.mark.parametrize("path", EMA, ids=[os.path.basename(p)[:-4] for p in EMA])
def test_ema_matches_reference_golden(path):
    import dcasr_b200 as d
    g = np.load(path)
    x, p = _t(g["x"], True), _t(g["p"], True)
    out = d.DynamicChunker._ema(x, p)
    assert max_err(out, _t(g["out"])) < 1e-5
    (out * _t(g["w"])).sum().backward()
    assert rel_err(x.grad, _t(g["gx"])) < 1e-5
    assert rel_err(p.grad, _t(g["gp"])) < 1e-4
    sat = (_t(g["p"]) >= 1 - 1e-4) | (_t(g["p"]) <= 1e-4)
    sat[:, 0] = False
    assert (p.grad[sat] == 0).all(), "hard clamp: gradient must be exactly zero at saturated P"


def test_reference_unit_properties():
    """The reference's own tests/test_hnet_chunk.py properties, on the CUDA path."""
    import dcasr_b200 as d
    torch.manual_seed(0)
    B, L, D = 4, 40, 32
    r = d.RoutingModule(D).to(DEV)
    x = torch.randn(B, L, D, device=DEV)
    p, b = r(x)
    assert torch.all(p >= 0) and torch.all(p <= 1) and torch.all(p[:, 0] == 1) and torch.all(b[:, 0] == 1)
    assert torch.all((b == 0) | (b == 1))
    p1, b1 = r(torch.ones(1, L, D, device=DEV))
    assert torch.allclose(p1[0, 1:], torch.zeros(L - 1, device=DEV), atol=1e-4) and b1[0, 1:].sum() == 0
    x2 = x.clone(); x2[0, 21:] = torch.randn(L - 21, D, device=DEV)
    assert torch.allclose(r(x)[0][0, :21], r(x2)[0][0, :21], atol=1e-6)          # causal
    # N = 1 exact identity, forward and gradient
    ch1 = d.DynamicChunker(D, N=1)
    xi = torch.randn(B, L, D, device=DEV, requires_grad=True)
    co = ch1.chunk(xi)
    assert torch.equal(co.z, xi) and float(co.ratio_loss) == 0.0 and float(co.kept_fraction) == 1.0
    y = ch1.dechunk(co.z, co)
    assert torch.equal(y, xi)
    y.sum().backward()
    assert torch.allclose(xi.grad, torch.ones_like(xi))
    # ratio loss
    pr = torch.rand(B, L, device=DEV); br = (pr >= 0.5).float()
    assert float(d.ratio_loss(pr, br, N=1)) == 0.0
    pg = torch.rand(B, L, device=DEV, requires_grad=True)
    d.ratio_loss(pg, (pg.detach() >= 0.5).float(), N=3).backward()
    assert torch.any(pg.grad != 0)
    F_, G_ = br.mean(), pr.mean()
    assert abs(float(d.ratio_loss(pr, br, 4)) - float(4 / 3 * (3 * F_ * G_ + (1 - F_) * (1 - G_)))) < 1e-6
    # masking ignores padding
    ch = d.DynamicChunker(D, N=2).to(DEV)
    mask = torch.ones(B, L, dtype=torch.bool, device=DEV); mask[:, L // 2:] = False
    co = ch.chunk(x, mask)
    assert float((co.b * (~mask).float()).sum()) == 0.0
    # gradcheck-style fp64 is not built (kernels are fp32/bf16); check saturated p instead
    xs = torch.randn(2, 120, D, device=DEV, requires_grad=True)
    ps = torch.rand(2, 120, device=DEV) * 0.8 + 0.1
    ps[:, ::10] = 1.0
    ps.requires_grad_(True)
    out = d.DynamicChunker._ema(xs, ps)
    out.sum().backward()
    assert torch.isfinite(out).all() and torch.isfinite(xs.grad).all() and torch.isfinite(ps.grad).all()
    assert (ps.grad[:, 10::10] == 0).all() and ps.grad[:, 1:].abs().sum() > 0


def test_bf16_long_sequence_compaction_exact():
    import dcasr_b200 as d
    torch.manual_seed(1)
    ch = d.DynamicChunker(32, N=2).to(DEV).to(torch.bfloat16)
    x = torch.randn(2, 1200, 32, device=DEV, dtype=torch.bfloat16)
    co = ch.chunk(x)
    assert int(co.membership.max()) < co.z.shape[1]
    for i in range(2):
        idx = torch.nonzero(co.b[i] > 0.5).squeeze(-1)
        assert torch.equal(co.z[i, :idx.numel()], x[i, idx])
        assert int(co.z_mask[i].sum()) == idx.numel()


@pytest.mark.parametrize("B,L,D,N", [(40, 398, 384, 2), (8, 1498, 512, 3), (3, 5000, 64, 2)])
def test_full_size_properties(B, L, D, N):
    """BASELINE-size inputs: integer outputs against torch's own cumsum/nonzero, EMA against the sequential
    recurrence, gather+STE against index arithmetic, compaction round trip (scatter then gather == identity on kept)."""
    import dcasr_b200 as d
    torch.manual_seed(2)
    ch = d.DynamicChunker(D, N=N).to(DEV)
    with torch.no_grad():
        ch.router.W_q.weight.add_(0.3 * torch.randn(D, D, device=DEV) / D ** 0.5)
    base = torch.randn(B, L, D, device=DEV)
    x = base + 1.5 * torch.roll(base, 1, 1) * (torch.rand(B, L, 1, device=DEV) > 0.5)
    lengths = torch.randint(L // 3, L + 1, (B,), device=DEV); lengths[0] = L
    mask = torch.arange(L, device=DEV)[None] < lengths[:, None]
    co = ch.chunk(x, mask)
    keep = co.b > 0.5
    assert torch.equal(co.membership, (keep.long().cumsum(1) - 1).clamp_min(0))
    counts = keep.sum(1)
    M = int(counts.max())
    assert co.z.shape == (B, M, D)
    assert torch.equal(co.z_mask, torch.arange(M, device=DEV)[None] < counts[:, None])
    bi, ti = keep.nonzero(as_tuple=True)
    assert torch.equal(co.z[bi, co.membership[bi, ti]], x[bi, ti])
    assert float(co.z[~co.z_mask].abs().sum()) == 0.0
    assert not keep[~mask].any() and keep[:, 0].all()
    # dechunk: EMA (sequential reference in fp64) + gather
    zp = torch.randn(B, M, D, device=DEV)
    y = ch.dechunk(zp, co)
    P = torch.zeros(B, M, device=DEV, dtype=torch.float64)
    P[bi, co.membership[bi, ti]] = co.p[bi, ti].double()
    pc = P.clamp(1e-4, 1 - 1e-4)
    zb = torch.empty(B, M, D, device=DEV, dtype=torch.float64)
    prev = zp[:, 0].double(); zb[:, 0] = prev
    for t in range(1, M):
        prev = pc[:, t, None] * zp[:, t].double() + (1 - pc[:, t, None]) * prev
        zb[:, t] = prev
    ref = torch.gather(zb, 1, co.membership[..., None].expand(B, L, D))
    assert rel_err(y, ref) < 1e-5


FIXED = sorted(p for p in glob.glob(os.path.join(GOLDEN, "fixed_*.npz")) if "identity" not in p)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("path", FIXED, ids=[os.path.basename(p)[:-4] for p in FIXED])
def test_fixed_pool_matches_reference_golden(path, dtype):
    """CUDA FixedPoolChunker (hnb_window_reduce / hnb_window_broadcast) against the reference's own outputs and
    autograd gradients: integer / boolean fields bit-exact, values 1e-6 in fp32 and bf16 rounding in bf16."""
    import dcasr_b200 as dd
    g = np.load(path)
    tol = 1e-6 if dtype == torch.float32 else 2e-2
    x = torch.from_numpy(g["x"]).to(DEV, dtype).requires_grad_(True)
    z_proc = torch.from_numpy(g["z_proc"]).to(DEV, dtype).requires_grad_(True)
    mask = torch.from_numpy(g["mask"]).to(DEV) if "mask" in g else None
    ch = dd.FixedPoolChunker(x.shape[-1], N=int(g["N"]))
    co = ch.chunk(x, mask)
    assert torch.equal(co.membership.cpu(), torch.from_numpy(g["membership"])) and co.membership.dtype == torch.int64
    assert torch.equal(co.z_mask.cpu(), torch.from_numpy(g["z_mask"])) and co.z_mask.dtype == torch.bool
    assert torch.equal(co.b.float().cpu(), torch.from_numpy(g["b"])) and torch.equal(co.p, co.b)
    assert rel_err(co.z.float().cpu(), torch.from_numpy(g["z"])) < max(tol, 2e-7)
    assert abs(float(co.kept_fraction) - float(g["kept_fraction"])) < 1e-6 and float(co.ratio_loss) == 0.0
    out = ch.dechunk(z_proc, co)
    assert torch.equal(out.float().cpu(), torch.from_numpy(g["z_proc"]).to(dtype).float()[
        torch.arange(out.shape[0])[:, None], torch.from_numpy(g["membership"])])       # a pure gather: bit exact
    w, v = torch.from_numpy(g["w"]).to(DEV), torch.from_numpy(g["v"]).to(DEV)
    ((out.float() * w).sum() + (co.z.float() * v).sum()).backward()
    assert rel_err(x.grad.float().cpu(), torch.from_numpy(g["dx"])) < max(tol, 1e-6)
    assert rel_err(z_proc.grad.float().cpu(), torch.from_numpy(g["dz_proc"])) < max(tol, 1e-6)
    # the encoder's fused form: residual added in the same pass
    r = torch.randn_like(out)
    assert rel_err(ch.dechunk(z_proc, co, residual=r).float(), (r + out).float()) < max(tol, 1e-6)


def test_encoder_with_fixed_chunker_runs_and_backpropagates():
    """`chunker: fixed` through the whole Type A encoder (bf16 autocast): finite features, a gradient on every parameter."""
    import dcasr_b200 as dd
    torch.manual_seed(3)
    enc = dd.DCASREncoder(n_mels=80, d_outer=128, d_main=128, n_enc=1, n_main=1, n_dec=1, arch_type="A", N=2,
                          chunker="fixed").to(DEV)
    feats, lens = torch.randn(3, 330, 80, device=DEV), torch.tensor([330, 250, 61], device=DEV)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o = enc(feats, lens)
    (o.features.float().pow(2).mean() + 0.03 * o.ratio_loss).backward()
    assert torch.isfinite(o.features).all() and abs(float(o.kept_fractions[0]) - 0.5) < 0.02
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in enc.parameters())
