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


def _passthrough(x: torch.Tensor, mask: torch.Tensor | None) -> ChunkOutput:
    """Stride 1: every frame is its own window.  Field for field what DynamicChunker returns at N = 1
    (fixed_pool.py:56-67): z is x itself, p = b = the mask as floats, membership = frame index."""
    B, L, _ = x.shape
    valid = x.new_ones(B, L) if mask is None else mask.to(x.dtype)
    index = torch.arange(L, device=x.device).repeat(B, 1)
    z_mask = torch.ones(B, L, dtype=torch.bool, device=x.device) if mask is None else mask
    return ChunkOutput(z=x, z_mask=z_mask, p=valid, b=valid, membership=index, ratio_loss=x.new_zeros(()),
                       kept_fraction=x.new_ones(()))


class FixedPoolChunker(nn.Module):
    """Masked mean over windows of ``stride`` frames, rate 1/N, nothing learned (fixed_pool.py:31-110)."""

    def __init__(self, d_model: int, N=1, ema_smoothing: bool = True):
        super().__init__()
        stride = int(round(float(N)))
        if abs(float(N) - stride) > 1e-6:               # Type B passes sqrt(N): only perfect squares have a window
            raise ValueError(f"FixedPoolChunker: the stride must be an integer, got N={N!r}")
        if stride < 1:
            raise ValueError(f"FixedPoolChunker: the stride must be at least 1, got {stride}")
        self.d_model, self.stride, self.N = d_model, stride, stride
        self.identity = stride == 1
        self.ema_smoothing = ema_smoothing              # accepted for interface parity; there is no probability to smooth with

    def chunk(self, x: torch.Tensor, mask: torch.Tensor | None = None) -> ChunkOutput:
        if self.identity:
            return _passthrough(x, mask)
        B, L, _ = x.shape
        s = self.stride
        n_valid = mask.sum(dim=1) if mask is not None else torch.full((B,), L, device=x.device, dtype=torch.long)
        windows = torch.div(n_valid + (s - 1), s, rounding_mode="floor").clamp_min(1)     # per row (:76)
        M = int(windows.max().item())                   # the one host sync of the stage: it sizes z (:77)
        t = torch.arange(L, device=x.device)
        membership = torch.div(t, s, rounding_mode="floor").clamp(max=M - 1).unsqueeze(0).expand(B, L)   # (:80-83)
        z, cnt = _PoolFn.apply(x, _mask_u8(mask), M, s)
        starts = (t % s == 0).to(x.dtype).unsqueeze(0).expand(B, L)                       # window starts (:91)
        b = starts if mask is None else starts * mask.to(x.dtype)
        kept = windows.sum().float() / n_valid.sum().clamp_min(1).float()                 # (:92)
        return ChunkOutput(z=z, z_mask=cnt > 0, p=b, b=b, membership=membership, ratio_loss=x.new_zeros(()),
                           kept_fraction=kept)

    def dechunk(self, z_proc: torch.Tensor, co: ChunkOutput, residual: torch.Tensor | None = None) -> torch.Tensor:
        """Every frame takes its window's processed vector (fixed_pool.py:96-104); ``residual`` (an extension the
        encoder uses) is added in the same pass.  The window of a frame is recomputed from its position,
        min(t // stride, M-1): exactly ``co.membership`` as chunk() of this class builds it."""
        if self.identity:
            return z_proc if residual is None else residual + z_proc
        return _BroadcastFn.apply(z_proc, residual, co.membership.shape[1], self.stride)

    def forward(self, x, mask=None):
        return self.chunk(x, mask)
