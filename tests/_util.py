"""Shared test helpers: repo paths, deterministic weights, error metrics."""
from __future__ import annotations

import math
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(REPO, "h-net-mamba-asr_b200")
GOLDEN = os.path.join(REPO, "tests", "golden")
for _p in (REPO, PKG_DIR):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def fill_weights(module: torch.nn.Module, seed: int = 0, router_identity: bool = False) -> None:
    """Deterministically (re)initialise every parameter from a seeded CPU generator, visiting
    the state_dict in sorted-key order so that the reference modules, the oracle and the CUDA
    product (identical key sets) receive identical values regardless of construction order."""
    g = torch.Generator().manual_seed(seed)
    sd = module.state_dict()
    new = {}
    for k in sorted(sd):
        t = sd[k]
        leaf = k.split(".")[-1]
        parent = k.split(".")[-2] if "." in k else ""
        if leaf == "A_log":
            v = torch.log(torch.empty(t.shape).uniform_(1.0, 16.0, generator=g))
        elif leaf == "dt_bias":
            dt = torch.exp(torch.rand(t.shape, generator=g) * (math.log(0.1) - math.log(0.001)) + math.log(0.001))
            dt = dt.clamp(min=1e-4)
            v = dt + torch.log(-torch.expm1(-dt))
        elif leaf == "D":
            v = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
        elif parent in ("norm",) and leaf == "weight":
            v = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
        elif leaf == "bias":
            v = 0.1 * torch.randn(t.shape, generator=g)
        elif parent in ("W_q", "W_k"):
            if router_identity:
                v = torch.eye(t.shape[0])
            else:
                v = torch.eye(t.shape[0]) + 0.5 * torch.randn(t.shape, generator=g) / math.sqrt(t.shape[1])
        elif parent == "conv1d":
            v = torch.randn(t.shape, generator=g) * 0.5
        elif t.dim() >= 2:
            fan_in = t[0].numel()
            v = torch.randn(t.shape, generator=g) / math.sqrt(fan_in)
        else:
            v = torch.randn(t.shape, generator=g)
        new[k] = v.to(t.dtype)
    module.load_state_dict(new)


def set_router_operating_point(module: torch.nn.Module, seed: int = 0, shift: float = 0.0) -> None:
    """Untrained routers (W_q = W_k = I, reference hnet_chunk.py:89-90) keep < 1 % of the frames of random input; the
    metric's configurations run at the TRAINED operating point, keep fraction ~ 1/N.  This sets every router in
    `module` to W_q = I, W_k = shift I + R / sqrt(d) with a seeded Gaussian R: cos(q_t, k_{t-1}) ~ N(c(shift), 1/d), so
    about half of the frames cross p >= 0.5 at shift 0 and fewer as shift grows (same device-independent values for
    the reference modules, the oracle and the CUDA product: the keys are visited in sorted order)."""
    g = torch.Generator().manual_seed(7000 + seed)
    sd = module.state_dict()
    new = {}
    for k in sorted(sd):
        parts = k.split(".")
        if len(parts) >= 3 and parts[-3] == "router" and parts[-1] == "weight":
            d = sd[k].shape[0]
            if parts[-2] == "W_q":
                new[k] = torch.eye(d)
            elif parts[-2] == "W_k":
                new[k] = shift * torch.eye(d) + torch.randn(d, d, generator=g) / math.sqrt(d)
    with torch.no_grad():
        for k, v in new.items():
            sd[k].copy_(v.to(sd[k].dtype))


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b|| in fp64 (the 'relative' of north_star's 1e-3 / 2e-2 tolerances)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = b.norm().item()
    return (a - b).norm().item() / max(den, 1e-30)


def max_err(a: torch.Tensor, b: torch.Tensor) -> float:
    return (a.detach().double().cpu() - b.detach().double().cpu()).abs().max().item()


def force_in_band_boundaries(monkeypatch, comparator_boundaries, band: float):
    """Teacher-force the product's boundary decisions INSIDE the band |p_cmp - 0.5| <= band to the comparator's, after
    asserting that p agrees within the band width and that every decision OUTSIDE the band already agrees (north_star:
    "bit-exact wherever |p - 0.5| > band").  One flipped in-band boundary re-indexes every later chunk; forcing the few
    in-band frames lets everything downstream be compared frame by frame, with no skip.  Returns the list that receives
    the number of forced frames per chunk stage."""
    from dcasr_b200 import ops
    forced, stage = [], [0]
    real_router_fwd = ops.router_fwd

    def router_fwd_forced(qk, mask_u8, Bq, Lq, D, pb_dtype, N):
        p, b, stats = real_router_fwd(qk, mask_u8, Bq, Lq, D, pb_dtype, N)
        pc, bc = comparator_boundaries[stage[0]]
        stage[0] += 1
        pc, bc = pc.to(p.device), bc.to(p.device)
        assert p.shape == pc.shape, "an earlier stage drew different boundaries"
        valid = mask_u8.view(Bq, Lq) > 0 if mask_u8 is not None else torch.ones_like(bc, dtype=torch.bool)
        assert max_err(p[valid], pc[valid]) < band
        outside = valid & ((pc - 0.5).abs() > band)
        assert torch.equal(b[outside], bc[outside].to(b.dtype)), "boundary differs outside the band"
        forced.append(int((b != bc.to(b.dtype))[valid].sum()))
        b = bc.to(b.dtype).clone()
        return p, b, ops.masked_ratio_stats(p, b, mask_u8, N)

    monkeypatch.setattr(ops, "router_fwd", router_fwd_forced)
    return forced
