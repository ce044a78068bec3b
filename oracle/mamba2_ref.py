"""Oracle: pure-torch restatement of ``mamba_ssm.Mamba2`` (test infrastructure).

The reference imports ``from mamba_ssm import Mamba2`` (src/dcasr/models/mamba_block.py:12)
and constructs it with ``d_model, d_state=128, d_conv=4, expand=2, headdim=64``
(src/dcasr/models/mamba_block.py:45); everything else is the package default
(ngroups=1, rmsnorm=True, norm_before_gate=False, bias=False, conv_bias=True,
chunk_size=256, dt_limit=(0, inf)).  mamba-ssm 2.3.2.post1 is not vendored, so
this file restates its *published* forward:

    zxbcdt = in_proj(u);  z, xBC, dt = split
    xBC = silu(causal_depthwise_conv1d(xBC))
    x, B, C = split(xBC);  dt = softplus(dt + dt_bias);  A = -exp(A_log)
    h_t = exp(dt_t A) h_{t-1} + dt_t B_t (x) x_t ;  y_t = C_t . h_t + D x_t
    y = rmsnorm(y * silu(z)) * norm.weight ;  out = out_proj(y)

Parameter names / shapes / ``_no_weight_decay`` tags equal the upstream module so
reference checkpoints load (dimension formulas: src/dcasr/eval/efficiency.py:49-57).

Under CUDA autocast (the reference's training precision, src/dcasr/training/trainer.py:187-190) the upstream conv1d /
SSD / gated-norm kernels are opaque custom ops: autocast decides the dtype of what goes IN (bf16 zxbcdt from the bf16
in_proj GEMM) and they return that dtype, computing in fp32 inside.  `forward` therefore runs those three stages with
autocast disabled on the bf16-rounded tensors and rounds their results back (conv output, y, normed y) -- autocast must
not be left to cast the einsums of this restatement, which would add roundings the reference's path does not have.

PARITY UNPINNED by the reference (no numeric vectors exist for this path); see
``oracle/__init__.py``.  Cross-checks live in tests/test_oracle_mamba2.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# functional pieces
# ----------------------------------------------------------------------------
def causal_conv1d_silu(xBC: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """xBC [B,L,C]; weight [C,1,K]; out[t] = silu(bias + sum_j w[:,j] * xBC[t-(K-1)+j]), zero left pad."""
    K = weight.shape[-1]
    xt = F.pad(xBC.transpose(1, 2), (K - 1, 0))
    out = F.conv1d(xt, weight, bias, groups=weight.shape[0])
    return F.silu(out).transpose(1, 2)


def ssd_sequential(x, dt, A, Bm, Cm, D):
    """Ground-truth recurrence, one timestep at a time.

    x [B,L,H,P]; dt [B,L,H] (already softplus'ed); A [H] (negative); Bm, Cm [B,L,N]; D [H].
    Returns y [B,L,H,P].  Works in whatever dtype it is given (use fp64 for truth).
    """
    Bsz, L, H, P = x.shape
    N = Bm.shape[-1]
    h = x.new_zeros(Bsz, H, P, N)
    ys = []
    for t in range(L):
        a = torch.exp(dt[:, t] * A)                                   # [B,H]
        u = dt[:, t, :, None] * x[:, t]                               # [B,H,P]
        h = a[:, :, None, None] * h + u[..., None] * Bm[:, t, None, None, :]
        y = (h * Cm[:, t, None, None, :]).sum(-1) + D[None, :, None] * x[:, t]
        ys.append(y)
    return torch.stack(ys, dim=1)


def ssd_chunked(x, dt, A, Bm, Cm, D, chunk: int = 64):
    """Chunked SSD (the 'ssd_minimal' dual form): same maths as ssd_sequential, O(L*chunk)."""
    Bsz, L, H, P = x.shape
    N = Bm.shape[-1]
    pad = (-L) % chunk
    if pad:
        x = F.pad(x, (0, 0, 0, 0, 0, pad))
        dt = F.pad(dt, (0, 0, 0, pad))
        Bm = F.pad(Bm, (0, 0, 0, pad))
        Cm = F.pad(Cm, (0, 0, 0, pad))
    nc = (L + pad) // chunk
    xs = x.reshape(Bsz, nc, chunk, H, P)
    dts = dt.reshape(Bsz, nc, chunk, H)
    Bs = Bm.reshape(Bsz, nc, chunk, N)
    Cs = Cm.reshape(Bsz, nc, chunk, N)
    dA = dts * A                                                     # [B,nc,Q,H]
    cs = dA.cumsum(dim=2)                                            # inclusive
    # intra-chunk: L[t,s] = exp(cs_t - cs_s) for s<=t
    seg = cs[:, :, :, None, :] - cs[:, :, None, :, :]                # [B,nc,t,s,H]
    tri = torch.ones(chunk, chunk, dtype=torch.bool, device=x.device).tril()
    Lmat = torch.exp(seg.masked_fill(~tri[None, None, :, :, None], float("-inf")))
    CB = torch.einsum("bctn,bcsn->bcts", Cs, Bs)
    u = xs * dts[..., None]                                          # dt * x
    y_diag = torch.einsum("bcts,bctsh,bcshp->bcthp", CB, Lmat, u)
    # per-chunk end states
    decay_to_end = torch.exp(cs[:, :, -1:, :] - cs)                  # [B,nc,Q,H]
    S_local = torch.einsum("bcsh,bcshp,bcsn->bchpn", decay_to_end, u, Bs)
    # inter-chunk pass
    S_in = []
    S = x.new_zeros(Bsz, H, P, N)
    tot = torch.exp(cs[:, :, -1, :])                                 # [B,nc,H]
    for c in range(nc):
        S_in.append(S)
        S = tot[:, c, :, None, None] * S + S_local[:, c]
    S_in = torch.stack(S_in, dim=1)                                  # [B,nc,H,P,N]
    y_off = torch.einsum("bctn,bchpn,bcth->bcthp", Cs, S_in, torch.exp(cs))
    y = (y_diag + y_off).reshape(Bsz, nc * chunk, H, P)[:, :L]
    return y + D[None, None, :, None] * x[:, :L]


def gated_rmsnorm(y: torch.Tensor, z: torch.Tensor, weight: torch.Tensor, eps: float = 1e-5):
    """norm_before_gate=False, ngroups=1: rmsnorm(y * silu(z)) * weight over the full width."""
    g = y * F.silu(z)
    return g * torch.rsqrt(g.pow(2).mean(-1, keepdim=True) + eps) * weight


# ----------------------------------------------------------------------------
# module with the upstream parameter layout
# ----------------------------------------------------------------------------
class _Norm(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.eps = 1e-5


class Mamba2Ref(nn.Module):
    """Same constructor subset and state_dict layout as ``mamba_ssm.Mamba2``."""

    def __init__(self, d_model, d_state=128, d_conv=4, expand=2, headdim=64, ngroups=1,
                 dt_min=0.001, dt_max=0.1, dt_init_floor=1e-4, A_init_range=(1, 16),
                 chunk_size=64, mode="chunked"):
        super().__init__()
        assert ngroups == 1
        self.d_model, self.d_state, self.d_conv, self.headdim = d_model, d_state, d_conv, headdim
        self.d_inner = expand * d_model
        assert self.d_inner % headdim == 0
        self.nheads = self.d_inner // headdim
        self.chunk_size, self.mode = chunk_size, mode
        d_in_proj = 2 * self.d_inner + 2 * d_state + self.nheads
        conv_dim = self.d_inner + 2 * d_state
        self.in_proj = nn.Linear(d_model, d_in_proj, bias=False)
        self.conv1d = nn.Conv1d(conv_dim, conv_dim, d_conv, groups=conv_dim, padding=d_conv - 1, bias=True)
        dt = torch.exp(torch.rand(self.nheads) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min))
        dt = dt.clamp(min=dt_init_floor)
        self.dt_bias = nn.Parameter(dt + torch.log(-torch.expm1(-dt)))
        self.dt_bias._no_weight_decay = True
        A = torch.empty(self.nheads).uniform_(*A_init_range)
        self.A_log = nn.Parameter(torch.log(A))
        self.A_log._no_weight_decay = True
        self.D = nn.Parameter(torch.ones(self.nheads))
        self.D._no_weight_decay = True
        self.norm = _Norm(self.d_inner)
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=False)

    def forward(self, u: torch.Tensor) -> torch.Tensor:
        Bsz, L, _ = u.shape
        di, N, H, P = self.d_inner, self.d_state, self.nheads, self.headdim
        zxbcdt = self.in_proj(u)                                  # bf16 under autocast
        act = zxbcdt.dtype
        cd = torch.promote_types(act, torch.float32)              # kernel-internal math in >= fp32
        with torch.autocast(u.device.type, enabled=False):
            z, xBC, dt = torch.split(zxbcdt, [di, di + 2 * N, H], dim=-1)
            xBC = causal_conv1d_silu(xBC.to(cd), self.conv1d.weight.to(cd), self.conv1d.bias.to(cd)).to(act)
            x, Bm, Cm = torch.split(xBC, [di, N, N], dim=-1)
            dtp = F.softplus(dt.to(cd) + self.dt_bias.to(cd))
            A = -torch.exp(self.A_log.to(cd))
            xh = x.to(cd).reshape(Bsz, L, H, P)
            if self.mode == "sequential":
                y = ssd_sequential(xh, dtp, A, Bm.to(cd), Cm.to(cd), self.D.to(cd))
            else:
                y = ssd_chunked(xh, dtp, A, Bm.to(cd), Cm.to(cd), self.D.to(cd), self.chunk_size)
            y = y.reshape(Bsz, L, di).to(act)
            y = gated_rmsnorm(y.to(cd), z.to(cd), self.norm.weight.to(cd), self.norm.eps).to(act)
        return self.out_proj(y)
