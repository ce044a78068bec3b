"""Oracle: CPU restatement of the reference H-Net dynamic-chunking stage (test infrastructure).

Follows /root/reference/src/dcasr/models/hnet_chunk.py function by function, written
as plain per-row / per-timestep torch code (differentiable through autograd) rather
than the reference's vectorised scatter / O(M^2) matmul, so that it is an independent
statement of the same algorithm:

  router_ref        <- RoutingModule.forward            hnet_chunk.py:92-111
  ratio_loss_ref    <- ratio_loss                       hnet_chunk.py:117-136
  chunk_ref         <- DynamicChunker.chunk             hnet_chunk.py:164-196
  ema_ref           <- DynamicChunker._ema              hnet_chunk.py:226-248
                       (== the sequential recurrence of tests/test_hnet_chunk.py:183-193)
  dechunk_ref       <- DynamicChunker.dechunk           hnet_chunk.py:199-224

PINNED: tests/golden/make_golden.py runs the reference module itself on seeded inputs
and tests/test_oracle_hnet.py checks this file against those vectors (values, integer
outputs bit-exact, and gradients).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass
class ChunkRef:
    z: torch.Tensor
    z_mask: torch.Tensor
    p: torch.Tensor
    b: torch.Tensor
    membership: torch.Tensor
    ratio_loss: torch.Tensor
    kept_fraction: torch.Tensor


def router_ref(x, Wq, Wk, mask=None, eps: float = 1e-6):
    """p_t = 0.5 (1 - cos(Wq x_t, Wk x_{t-1})), p_0 = 1, clamp, b = [p >= 0.5], mask.

    cos = normalise-each-then-dot with norms clamped at eps (F.cosine_similarity semantics,
    SURVEY.md App. A)."""
    q = x @ Wq.t()
    k = x @ Wk.t()
    qn = q / q.norm(dim=-1, keepdim=True).clamp_min(eps)
    kn = k / k.norm(dim=-1, keepdim=True).clamp_min(eps)
    cos = torch.zeros(x.shape[:2], dtype=q.dtype, device=x.device)
    cos[:, 1:] = (qn[:, 1:] * kn[:, :-1]).sum(-1)
    p = 0.5 * (1.0 - cos)
    first = torch.zeros_like(p, dtype=torch.bool)
    first[:, 0] = True
    p = torch.where(first, torch.ones_like(p), p)          # p_1 := 1 (no grad through it)
    p = p.clamp(0.0, 1.0)
    b = (p >= 0.5).to(p.dtype)
    if mask is not None:
        m = mask.to(p.dtype)
        p = p * m
        b = b * m
    return p, b


def ratio_loss_ref(p, b, N, mask=None):
    if N == 1:
        return p.new_zeros(())
    p, b = p.float(), b.float()
    if mask is None:
        Fh, G = b.mean(), p.mean()
    else:
        m = mask.float()
        den = m.sum().clamp_min(1.0)
        Fh, G = (b * m).sum() / den, (p * m).sum() / den
    return (N / (N - 1.0)) * ((N - 1.0) * Fh * G + (1.0 - Fh) * (1.0 - G))


def chunk_ref(x, Wq, Wk, N, mask=None):
    Bsz, L, D = x.shape
    p, b = router_ref(x, Wq, Wk, mask)
    rl = ratio_loss_ref(p, b, N, mask)
    keep = b > 0.5
    memb = torch.zeros(Bsz, L, dtype=torch.int64)
    counts = []
    for i in range(Bsz):
        c = 0
        for t in range(L):
            if keep[i, t]:
                c += 1
            memb[i, t] = max(c - 1, 0)
        counts.append(c)
    M = max(max(counts) if counts else 0, 1)
    rows, zm = [], torch.zeros(Bsz, M, dtype=torch.bool)
    for i in range(Bsz):
        idx = torch.nonzero(keep[i]).squeeze(-1)
        zi = x[i, idx]
        if idx.numel() < M:
            zi = torch.cat([zi, x.new_zeros(M - idx.numel(), D)], 0)
        rows.append(zi)
        zm[i, :idx.numel()] = True
    z = torch.stack(rows, 0)
    valid = mask.sum() if mask is not None else torch.tensor(Bsz * L)
    kept = keep.sum().float() / valid.float().clamp_min(1.0)
    return ChunkRef(z, zm, p, b, memb, rl, kept)


def _hard_clamp(p, lo, hi):
    """clamp with zero gradient outside [lo, hi] (torch.clamp semantics)."""
    return p.clamp(lo, hi)


def ema_ref(x, P, p_clamp: float = 1e-4):
    """out_0 = x_0; out_t = pc_t x_t + (1 - pc_t) out_{t-1}, pc = clamp(P, 1e-4, 1-1e-4)."""
    Bsz, M, D = x.shape
    if M == 1:
        return x.clone()
    pc = _hard_clamp(P, p_clamp, 1.0 - p_clamp)
    wd = torch.promote_types(P.dtype, torch.float32)
    outs = [x[:, 0].to(wd)]
    for t in range(1, M):
        a = pc[:, t, None].to(wd)
        outs.append(a * x[:, t].to(wd) + (1.0 - a) * outs[-1])
    return torch.stack(outs, 1).to(x.dtype)


def dechunk_ref(z_proc, co: ChunkRef, ema: bool = True):
    Bsz, L = co.membership.shape
    M, D = z_proc.shape[1], z_proc.shape[2]
    if ema:
        keep = co.b > 0.5
        P = co.p.new_zeros(Bsz, M)
        for i in range(Bsz):
            idx = torch.nonzero(keep[i]).squeeze(-1)
            if idx.numel():
                P = P.index_put((torch.full_like(idx, i), torch.arange(idx.numel())), co.p[i, idx])
        z_proc = ema_ref(z_proc, P)
    x_up = torch.stack([z_proc[i, co.membership[i]] for i in range(Bsz)], 0)
    c = torch.where(co.b > 0.5, co.p, 1.0 - co.p)
    ste = (c + (1.0 - c).detach()).unsqueeze(-1)
    return x_up * ste.to(x_up.dtype)
