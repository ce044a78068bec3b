"""Fixed-stride pooling chunker on the CUDA path — host-side mirror of the reference's H2 control
``dcasr.models.fixed_pool.FixedPoolChunker`` (/root/reference/src/dcasr/models/fixed_pool.py:31-110).

Same constructor, attributes (``stride``, ``N``, ``identity``, ``ema_smoothing``), ``ValueError`` behaviour,
``chunk / dechunk / forward`` and ``ChunkOutput`` contract as the reference, no parameters, zero ratio loss.
The masked mean and the broadcast run in ``hnb_window_reduce`` / ``hnb_window_broadcast``
(csrc/fixed_pool_kernels.cu); this file wires autograd and the small index tensors of the contract.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .hnet_chunk import ChunkOutput, _mask_u8


class _PoolFn(torch.autograd.Function):
    """z[b,j] = masked mean of x over window j (fp32 accumulation, fixed_pool.py:84-89); cnt = valid frames per window."""

    @staticmethod
    def forward(ctx, x, mask_u8, M, stride):
        x = x if x.is_contiguous() else x.contiguous()
        z, cnt = ops.window_reduce(x, mask_u8, M, stride, True, x.dtype)
        ctx.save_for_backward(mask_u8, cnt)
        ctx.meta = (x.shape[1], stride, x.dtype)
        ctx.mark_non_differentiable(cnt)
        return z, cnt

    @staticmethod
    def backward(ctx, dz, _dcnt):
        mask_u8, cnt = ctx.saved_tensors
        L, stride, xdt = ctx.meta
        dz = dz if dz.is_contiguous() else dz.contiguous()
        dx = ops.window_broadcast(dz, mask_u8, cnt, None, L, stride, xdt)     # m / max(cnt,1) * dz[window]
        return dx, None, None, None


class _BroadcastFn(torch.autograd.Function):
    """out[b,t] = z[b, min(t // stride, M-1)] (+ resid[b,t]): the gather of fixed_pool.py:96-104."""

    @staticmethod
    def forward(ctx, z, resid, L, stride):
        z = z if z.is_contiguous() else z.contiguous()
        if resid is not None:
            resid = resid if resid.is_contiguous() else resid.contiguous()
        out = ops.window_broadcast(z, None, None, resid, L, stride, resid.dtype if resid is not None else z.dtype)
        ctx.meta = (z.shape[1], stride, z.dtype, resid is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        M, stride, zdt, has_resid = ctx.meta
        dout = dout if dout.is_contiguous() else dout.contiguous()
        dz, _ = ops.window_reduce(dout, None, M, stride, False, zdt, want_cnt=False)      # sum over the window's frames
        return dz, (dout if has_resid else None), None, None


class FixedPoolChunker(nn.Module):
    """Fixed-stride masked mean pooling to rate 1/N (no learned boundaries; fixed_pool.py:31-110)."""

    def __init__(self, d_model: int, N=1, ema_smoothing: bool = True):
        super().__init__()
        n = float(N)
        stride = int(round(n))
        if abs(n - stride) > 1e-6:
            raise ValueError(
                f"FixedPoolChunker needs an integer stride; got N={N!r}. Fixed-stride "
                "pooling has no fractional window — Type B fixed-pool is only defined at "
                "perfect-square N (so √N is an integer).")
        if stride < 1:
            raise ValueError(f"FixedPoolChunker stride must be >= 1, got {stride}")
        self.d_model = d_model
        self.stride = stride
        self.N = stride
        self.identity = (stride == 1)
        self.ema_smoothing = ema_smoothing      # interface parity: fixed pooling has no probability signal to smooth with

    def chunk(self, x: torch.Tensor, mask: torch.Tensor | None = None) -> ChunkOutput:
        B, L, D = x.shape
        s = self.stride
        if self.identity:                       # exact passthrough, field for field DynamicChunker's N = 1 (:56-67)
            ones = x.new_ones(B, L)
            memb = torch.arange(L, device=x.device).unsqueeze(0).expand(B, L).clone()
            if mask is not None:
                ones = ones * mask.to(x.dtype)
            return ChunkOutput(z=x, z_mask=(mask if mask is not None else x.new_ones(B, L, dtype=torch.bool)),
                               p=ones, b=ones, membership=memb, ratio_loss=x.new_zeros(()),
                               kept_fraction=x.new_ones(()))
        if mask is not None:
            lengths = mask.sum(dim=1)
            m = mask.to(x.dtype)
        else:
            lengths = torch.full((B,), L, device=x.device, dtype=torch.long)
            m = x.new_ones(B, L)
        nwin = ((lengths + s - 1) // s).clamp_min(1)
        M = int(nwin.max().item())              # the one host sync of the stage (sizes z), as in the reference (:78)
        pos = torch.arange(L, device=x.device)
        memb = (pos // s).clamp(max=M - 1).unsqueeze(0).expand(B, L)
        z, cnt = _PoolFn.apply(x, _mask_u8(mask), M, s)
        z_mask = cnt > 0
        b = (pos % s == 0).to(x.dtype).unsqueeze(0).expand(B, L) * m
        kept = nwin.sum().float() / lengths.sum().clamp_min(1).float()
        return ChunkOutput(z=z, z_mask=z_mask, p=b, b=b, membership=memb, ratio_loss=x.new_zeros(()),
                           kept_fraction=kept)

    def dechunk(self, z_proc: torch.Tensor, co: ChunkOutput, residual: torch.Tensor | None = None) -> torch.Tensor:
        """Broadcast each processed window vector back over its fine frames (identity at N = 1).  ``residual`` (an
        extension used by the encoder) is added in the same pass.  The window of a frame is recomputed from its
        position, min(t // stride, M-1): exactly ``co.membership`` as chunk() of this class builds it."""
        if self.identity:
            return z_proc if residual is None else residual + z_proc
        L = co.membership.shape[1]
        return _BroadcastFn.apply(z_proc, residual, L, self.stride)

    def forward(self, x, mask=None):
        return self.chunk(x, mask)
