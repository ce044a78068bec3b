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

A second, VECTORISED statement of the same three stages (chunk_vec / ema_vec / dechunk_vec:
rank-by-cumsum, stable-sort compaction, log-step linear scan) runs on any device in a few
tensor ops.  It is what EncoderRef uses: bench.py's CPU arm should time tensor code, as the
reference's own module is, not a Python loop per frame, and the bf16-vs-bf16
comparator of the GPU tests runs this oracle on the B200 under autocast.

PINNED: tests/golden/make_golden.py runs the reference module itself on seeded inputs
and tests/test_oracle_hnet.py checks BOTH statements in this file against those vectors
(values, integer outputs bit-exact, and gradients).
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


def _linear(x, W):
    """x W^T through F.linear so that CUDA autocast treats it like the reference's nn.Linear (hnet_chunk.py:95-96)."""
    return torch.nn.functional.linear(x, W)


def router_ref(x, Wq, Wk, mask=None, eps: float = 1e-6):
    """p_t = 0.5 (1 - cos(Wq x_t, Wk x_{t-1})), p_0 = 1, clamp, b = [p >= 0.5], mask.

    cos = normalise-each-then-dot with norms clamped at eps (F.cosine_similarity semantics,
    SURVEY.md App. A)."""
    q = _linear(x, Wq)
    k = _linear(x, Wk)
    if q.dtype != torch.float64:            # autocast: cosine_similarity runs in fp32 (hnet_chunk.py:99)
        q, k = q.float(), k.float()
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


# ------------------------------------------------------------------------------------------------
# vectorised statement (any device)
# ------------------------------------------------------------------------------------------------
def chunk_vec(x, Wq, Wk, N, mask=None):
    """chunk_ref without Python loops: membership = rank of the latest kept frame (integer cumsum, hnet_chunk.py:182-183),
    compaction by a STABLE sort that moves kept frames to the front in time order (the reference scatters through
    nonzero(), hnet_chunk.py:188-191: same result, different mechanism)."""
    Bsz, L, D = x.shape
    p, b = router_ref(x, Wq, Wk, mask)
    rl = ratio_loss_ref(p, b, N, mask)
    keep = b > 0.5
    rank = keep.to(torch.int64).cumsum(1)
    memb = (rank - 1).clamp_min(0)
    counts = keep.sum(1)
    M = max(int(counts.max()) if counts.numel() else 0, 1)
    order = torch.argsort((~keep).to(torch.int8), dim=1, stable=True)[:, :M]        # kept frames first, in time order
    zm = torch.arange(M, device=x.device)[None, :] < counts[:, None]
    z = torch.gather(x, 1, order.unsqueeze(-1).expand(Bsz, M, D)) * zm.unsqueeze(-1).to(x.dtype)
    valid = mask.sum() if mask is not None else torch.tensor(Bsz * L, device=x.device)
    kept = keep.sum().float() / valid.float().clamp_min(1.0)
    return ChunkRef(z, zm, p, b, memb, rl, kept)


def ema_vec(x, P, p_clamp: float = 1e-4):
    """ema_ref as a log-step (Hillis-Steele) scan of the affine maps  h -> a_t h + u_t : ceil(log2 M) tensor steps, linear
    memory (the reference materialises the M x M weight matrix, hnet_chunk.py:243-248; same recurrence)."""
    Bsz, M, D = x.shape
    if M == 1:
        return x.clone()
    wd = torch.promote_types(P.dtype, torch.float32)
    pc = _hard_clamp(P, p_clamp, 1.0 - p_clamp).to(wd)
    first = torch.zeros(1, M, dtype=torch.bool, device=x.device)
    first[:, 0] = True
    a = torch.where(first, torch.zeros_like(pc), 1.0 - pc)                          # out_0 = x_0: a_0 = 0, u_0 = x_0
    u = torch.where(first.unsqueeze(-1), x.to(wd), pc.unsqueeze(-1) * x.to(wd))
    step = 1
    while step < M:
        a_prev = torch.nn.functional.pad(a[:, :-step], (step, 0), value=1.0)
        u_prev = torch.nn.functional.pad(u[:, :-step], (0, 0, step, 0), value=0.0)
        u = u + a.unsqueeze(-1) * u_prev
        a = a * a_prev
        step *= 2
    return u.to(x.dtype)


def dechunk_vec(z_proc, co: ChunkRef, ema: bool = True):
    Bsz, L = co.membership.shape
    M, D = z_proc.shape[1], z_proc.shape[2]
    keep = co.b > 0.5
    if ema:
        order = torch.argsort((~keep).to(torch.int8), dim=1, stable=True)[:, :M]
        live = torch.arange(M, device=z_proc.device)[None, :] < keep.sum(1)[:, None]
        P = torch.gather(co.p, 1, order) * live.to(co.p.dtype)                      # p at the j-th kept frame, pad slots 0
        z_proc = ema_vec(z_proc, P)
    x_up = torch.gather(z_proc, 1, co.membership.unsqueeze(-1).expand(Bsz, L, D))
    c = torch.where(keep, co.p, 1.0 - co.p)
    ste = (c + (1.0 - c).detach()).unsqueeze(-1)
    return x_up * ste.to(x_up.dtype)
