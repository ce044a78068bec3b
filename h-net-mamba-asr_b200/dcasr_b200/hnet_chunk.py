"""H-Net dynamic chunking on the CUDA hot path — host-side mirror of the reference module
``dcasr.models.hnet_chunk`` (/root/reference/src/dcasr/models/hnet_chunk.py).

Same public names, constructor signatures, ``state_dict`` keys and error behaviour:
``ChunkOutput`` (:61-70), ``RoutingModule`` (:76-111), ``ratio_loss`` (:117-136),
``DynamicChunker.chunk / dechunk / _ema / forward`` (:142-252).  The arithmetic runs in the
hand-written sm_100a kernels behind include/hnet_b200.h; this file only wires autograd.

Differences in mechanism (not in results): one host sync per ``chunk`` (the read of M) instead of
three; a linear-time EMA scan instead of the O(M^2) matmul; the STE/upsample/residual are one kernel.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn

from . import ops
from ._lib import HnbError


def _autocast_dtype():
    if torch.is_autocast_enabled("cuda"):
        dt = torch.get_autocast_dtype("cuda")
        if dt != torch.bfloat16:
            raise HnbError(f"hnet_b200 supports bf16 autocast only, got {dt}")
        return dt
    return None


def _mask_u8(mask):
    if mask is None:
        return None
    m = mask if mask.dtype == torch.bool else (mask != 0)
    return m.contiguous().view(torch.uint8)


@dataclass
class ChunkOutput:
    """Field-for-field the reference's ChunkOutput; the underscored extras feed dechunk()."""
    z: torch.Tensor
    z_mask: torch.Tensor
    p: torch.Tensor
    b: torch.Tensor
    membership: torch.Tensor
    ratio_loss: torch.Tensor
    kept_fraction: torch.Tensor
    _P: torch.Tensor | None = None          # [B, M] float: p at the kept frames (downsampled)
    _starts: torch.Tensor | None = None     # [B, M] int32: frame index of each kept frame
    _counts: torch.Tensor | None = None     # [B] int32: chunks per row


# ------------------------------------------------------------------------------------------------
# autograd functions
# ------------------------------------------------------------------------------------------------
_WQK_CACHE: dict = {}


def _packed_router_weights(Wq, Wk, adt):
    """W_q | W_k stacked and cast for the one router GEMM; rebuilt only when a parameter was written (its version counter
    moves with every optimiser step) -- under no_grad / decode this is once per model, not once per call.  An entry is
    valid only for the very tensor objects it was built from (weak references): a freed parameter's address is reused by
    the allocator, and a new tensor at the same address with the same version count must not hit."""
    import weakref
    key = (Wq.data_ptr(), Wk.data_ptr(), adt)
    ver = (Wq._version, Wk._version)
    hit = _WQK_CACHE.get(key)
    if hit is not None and hit[0] == ver and hit[2]() is Wq and hit[3]() is Wk:
        return hit[1]
    w = torch.cat([Wq.detach(), Wk.detach()], 0).to(adt)
    if len(_WQK_CACHE) > 64:
        _WQK_CACHE.clear()
    _WQK_CACHE[key] = (ver, w, weakref.ref(Wq), weakref.ref(Wk))
    return w


def _router_core(x, Wq, Wk, mask_u8, N):
    B, L, D = x.shape
    adt = _autocast_dtype() or x.dtype
    x2 = x.reshape(B * L, D)
    xa = x2 if x2.dtype == adt else x2.to(adt)
    xa = xa if xa.is_contiguous() else xa.contiguous()
    Wqk = _packed_router_weights(Wq, Wk, adt)
    qk = ops.gemm(xa, Wqk)                                              # [B*L, 2D]
    pb_dtype = torch.float32 if (adt != x.dtype or x.dtype == torch.float32) else x.dtype
    p, b, stats = ops.router_fwd(qk, mask_u8, B, L, D, pb_dtype, N)
    return xa, Wqk, qk, p, b, stats


def _router_backward(xa, Wqk, qk, mask_u8, stats, shape, x_dtype, N, dp, dratio):
    B, L, D = shape
    if dp is not None:
        dp = dp.to(torch.float32).contiguous()
    if dratio is not None:
        dratio = dratio.to(torch.float32).reshape(1).contiguous()
    dqk = ops.router_bwd(qk, mask_u8, B, L, D, dp, dratio, stats, N)
    dx = ops.gemm(dqk, Wqk, trans_b=True, out_dtype=x_dtype)            # dgrad  [B*L, D]
    if dx.dtype != x_dtype:
        dx = dx.to(x_dtype)
    sk = ops.wgrad_splitk(B * L, 2 * D, D) if dqk.dtype == torch.bfloat16 else 1
    dW = ops.gemm(dqk, xa, trans_a=True, trans_b=True, splitk=sk, out_dtype=torch.float32)   # [2D, D]
    return dx.view(B, L, D), dW


class _RouterFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Wq, Wk, mask_u8, N):
        xa, Wqk, qk, p, b, stats = _router_core(x, Wq, Wk, mask_u8, N)
        ctx.save_for_backward(xa, Wqk, qk, mask_u8, stats)
        ctx.meta = (tuple(x.shape), x.dtype, N, Wq.dtype)
        ctx.mark_non_differentiable(b)
        return p, b

    @staticmethod
    def backward(ctx, dp, _db):
        xa, Wqk, qk, mask_u8, stats = ctx.saved_tensors
        shape, xdt, N, wdt = ctx.meta
        dx, dW = _router_backward(xa, Wqk, qk, mask_u8, stats, shape, xdt, N, dp, None)
        D = shape[2]
        return dx, dW[:D].to(wdt), dW[D:].to(wdt), None, None


class _ChunkFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Wq, Wk, mask_u8, N):
        B, L, D = x.shape
        xc = x if x.is_contiguous() else x.contiguous()
        xa, Wqk, qk, p, b, stats = _router_core(xc, Wq, Wk, mask_u8, N)
        memb, counts = ops.boundary_scan(b, B, L)
        M = max(1, int(counts.max().item()))                           # the one host sync: z's shape is data dependent
        z, zmask, P, starts = ops.compact_rows(xc, p, b, memb, counts, M)
        ratio, kept = stats[0].clone(), stats[1].clone()
        zmask = zmask.view(torch.bool)
        ctx.save_for_backward(xa, Wqk, qk, mask_u8, stats, b, memb)
        ctx.meta = (tuple(x.shape), x.dtype, N, Wq.dtype)
        ctx.mark_non_differentiable(b, zmask, memb, kept, P, starts, counts)
        return z, p, ratio, b, zmask, memb, kept, P, starts, counts

    @staticmethod
    def backward(ctx, dz, dp, dratio, *_):
        xa, Wqk, qk, mask_u8, stats, b, memb = ctx.saved_tensors
        shape, xdt, N, wdt = ctx.meta
        B, L, D = shape
        dx = dWq = dWk = None
        if dp is not None or dratio is not None:
            dx, dW = _router_backward(xa, Wqk, qk, mask_u8, stats, shape, xdt, N, dp, dratio)
            dWq, dWk = dW[:D].to(wdt), dW[D:].to(wdt)
        if dz is not None:
            dzc = dz.to(xdt).contiguous()
            dx = ops.compact_rows_bwd(dzc, b, memb, L, dx=dx if dx is None else dx.contiguous())
        return dx, dWq, dWk, None, None


class _DechunkFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z_proc, p, resid, b, memb, P, starts, counts, ema, p_clamp):
        zc = z_proc if z_proc.is_contiguous() else z_proc.contiguous()
        zbar = ops.ema_fwd(zc, P, p_clamp) if ema else zc
        y_dtype = torch.promote_types(zc.dtype, p.dtype)
        if resid is not None:
            y_dtype = torch.promote_types(y_dtype, resid.dtype)
            resid = resid.to(y_dtype).contiguous()
        y = ops.upsample_fwd(zbar, memb, p.contiguous(), b, resid, y_dtype)
        ctx.save_for_backward(zc, zbar, p, b, memb, P, starts, counts)
        ctx.meta = (ema, p_clamp, resid is not None, z_proc.dtype, p.dtype,
                    resid.dtype if resid is not None else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        zc, zbar, p, b, memb, P, starts, counts = ctx.saved_tensors
        ema, p_clamp, has_res, zdt, pdt, rdt = ctx.meta
        dyc = dy.contiguous()
        dzbar, dp = ops.upsample_bwd(dyc, zbar, memb, starts, counts, p.contiguous(), b)
        if ema:
            dz, dP = ops.ema_bwd(dzbar, zc, zbar, P, p_clamp)
            dp = dp + torch.where(b > 0.5, dP.gather(1, memb), dP.new_zeros(()))   # P_j = p at the j-th kept frame
        else:
            dz = dzbar
        return dz.to(zdt), dp.to(pdt), (dyc.to(rdt) if has_res else None), None, None, None, None, None, None, None


class _EmaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, p_clamp):
        xc = x.contiguous()
        # weights in >= fp32 (hnet_chunk.py:242); float64 inputs (the reference's gradcheck tests) keep float64 throughout
        Pf = p.to(torch.float64 if xc.dtype == torch.float64 else torch.float32).contiguous()
        out = ops.ema_fwd(xc, Pf, p_clamp)
        ctx.save_for_backward(xc, out, Pf)
        ctx.meta = (p_clamp, p.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        xc, out, Pf = ctx.saved_tensors
        p_clamp, pdt = ctx.meta
        dx, dP = ops.ema_bwd(dout.contiguous(), xc, out, Pf, p_clamp)
        return dx, dP.to(pdt), None


class _RatioLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, b, mask_u8, N):
        stats = ops.masked_ratio_stats(p.contiguous(), b.to(p.dtype).contiguous(), mask_u8, N)
        ctx.save_for_backward(stats, mask_u8)
        ctx.meta = (N, p.dtype, tuple(p.shape))
        return stats[0].clone()

    @staticmethod
    def backward(ctx, g):
        stats, mask_u8 = ctx.saved_tensors
        N, pdt, shape = ctx.meta
        F_, den = stats[2], stats[4]
        coef = (N / (N - 1.0)) * ((N - 1.0) * F_ - (1.0 - F_)) / den * g
        dp = coef.expand(shape)
        if mask_u8 is not None:
            dp = dp * mask_u8.view(shape).to(dp.dtype)
        return dp.to(pdt), None, None, None


# ------------------------------------------------------------------------------------------------
# public API (reference names)
# ------------------------------------------------------------------------------------------------
class RoutingModule(nn.Module):
    """p_t = 0.5 (1 - cos(W_q x_t, W_k x_{t-1})), b_t = [p_t >= 0.5], p_1 = 1  (hnet_chunk.py:76-111)."""

    def __init__(self, d_model: int, eps: float = 1e-6):
        super().__init__()
        if eps != 1e-6:
            raise HnbError("RoutingModule: the kernels are built for eps=1e-6")
        self.d_model, self.eps = d_model, eps
        self.W_q = nn.Linear(d_model, d_model, bias=False)
        self.W_k = nn.Linear(d_model, d_model, bias=False)
        nn.init.eye_(self.W_q.weight)
        nn.init.eye_(self.W_k.weight)

    def forward(self, x: torch.Tensor, mask: torch.Tensor | None = None):
        return _RouterFn.apply(x, self.W_q.weight, self.W_k.weight, _mask_u8(mask), 2.0)


def ratio_loss(p: torch.Tensor, b: torch.Tensor, N, mask: torch.Tensor | None = None) -> torch.Tensor:
    """N/(N-1) [(N-1) F G + (1-F)(1-G)], F = masked mean b, G = masked mean p, in fp32 (hnet_chunk.py:117-136)."""
    if N == 1:
        return p.new_zeros(())
    return _RatioLossFn.apply(p, b, _mask_u8(mask), float(N))


class DynamicChunker(nn.Module):
    """One H-Net dynamic-chunking block (hnet_chunk.py:142-252): chunk() downsamples, dechunk() restores."""

    def __init__(self, d_model: int, N=1, ema_smoothing: bool = True):
        super().__init__()
        assert N >= 1
        self.d_model, self.N, self.ema_smoothing = d_model, N, ema_smoothing
        self.identity = (N == 1)
        self.router = None if self.identity else RoutingModule(d_model)

    def chunk(self, x: torch.Tensor, mask: torch.Tensor | None = None) -> ChunkOutput:
        B, L, D = x.shape
        if self.identity:                                   # N = 1: exact passthrough, nothing to compute
            ones = x.new_ones(B, L)
            memb = torch.arange(L, device=x.device).unsqueeze(0).expand(B, L).clone()
            if mask is not None:
                ones = ones * mask.to(x.dtype)
            zm = mask if mask is not None else x.new_ones(B, L, dtype=torch.bool)
            return ChunkOutput(z=x, z_mask=zm, p=ones, b=ones, membership=memb,
                               ratio_loss=x.new_zeros(()), kept_fraction=x.new_ones(()))
        z, p, ratio, b, zmask, memb, kept, P, starts, counts = _ChunkFn.apply(
            x, self.router.W_q.weight, self.router.W_k.weight, _mask_u8(mask), float(self.N))
        return ChunkOutput(z=z, z_mask=zmask, p=p, b=b, membership=memb, ratio_loss=ratio,
                           kept_fraction=kept, _P=P, _starts=starts, _counts=counts)

    def dechunk(self, z_proc: torch.Tensor, co: ChunkOutput, residual: torch.Tensor | None = None) -> torch.Tensor:
        """EMA over the compressed sequence -> gather by membership -> confidence STE (Eq. 5 -> 8 -> 9).
        ``residual`` (not in the reference signature) fuses the encoder's ``x_enc + dechunk(...)`` add."""
        if self.identity:
            return z_proc if residual is None else residual + z_proc
        P, starts, counts = co._P, co._starts, co._counts
        if P is None:                                       # a hand-built ChunkOutput: rebuild the helpers
            B, L = co.membership.shape
            bb = co.b.contiguous()
            _, counts = ops.boundary_scan(bb, B, L)
            if int(counts.max()) > z_proc.shape[1]:         # the reference's gather raises here too (index out of range)
                raise IndexError(f"dechunk: z_proc holds {z_proc.shape[1]} chunks per row but the boundaries define "
                                 f"{int(counts.max())}")
            _, _, P, starts = ops.compact_rows(co.p.detach().reshape(B, L, 1).contiguous(), co.p.detach().contiguous(),
                                               bb, co.membership.contiguous(), counts, z_proc.shape[1])
        return _DechunkFn.apply(z_proc, co.p, residual, co.b, co.membership, P, starts, counts,
                                bool(self.ema_smoothing), 1e-4)

    @staticmethod
    def _ema(x: torch.Tensor, p: torch.Tensor, p_clamp: float = 1e-4) -> torch.Tensor:
        """out_0 = x_0; out_t = pc_t x_t + (1 - pc_t) out_{t-1}, pc hard-clamped (zero grad at saturation)."""
        return _EmaFn.apply(x, p, float(p_clamp))

    def forward(self, x, mask=None):
        return self.chunk(x, mask)
