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


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b|| in fp64 (the 'relative' of north_star's 1e-3 / 2e-2 tolerances)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    den = b.norm().item()
    return (a - b).norm().item() / max(den, 1e-30)


def max_err(a: torch.Tensor, b: torch.Tensor) -> float:
    return (a.detach().double().cpu() - b.detach().double().cpu()).abs().max().item()
