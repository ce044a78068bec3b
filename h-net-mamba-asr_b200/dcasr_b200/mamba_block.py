"""Mamba-2 mixer, bidirectional pre-norm block and stack on the CUDA hot path — host-side mirror of
``mamba_ssm.Mamba2`` (as used by the reference) and of ``dcasr.models.mamba_block``
(/root/reference/src/dcasr/models/mamba_block.py: reverse_sequences :19-28, MambaBlock :31-56,
MambaStack :59-73).  Parameter names, shapes and ``_no_weight_decay`` tags equal the upstream module
(SURVEY.md §3.4), so reference checkpoints load unchanged.

One ``MambaBlock`` call is ONE autograd node:
    LayerNorm -> in_proj of BOTH directions as one GEMM -> conv1d+SiLU / softplus(dt) written in scan
    order (the length-aware reversal is index arithmetic, never a gather) -> SSD scan -> gated RMSNorm
    (back to natural order) -> out_proj of both directions as one GEMM with the residual add.
"""
from __future__ import annotations

import ctypes
import math
import os

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import HnbError, dtype_code, lib, stream
from .hnet_chunk import _autocast_dtype

# MambaBlock through the library's one-call-per-block composites (hnb_block_fwd / hnb_block_bwd).  "0" keeps the
# per-kernel Python orchestration (_MixerFn), which is also what a profiled step uses: bench.py times every entry point.
BLOCK_COMPOSITE = os.environ.get("HNB_BLOCK_COMPOSITE", "1") != "0"
# MambaStack packs the weights of all its blocks in one launch ("0": every block packs its own inside hnb_block_fwd)
STACK_PACK = os.environ.get("HNB_STACK_PACK", "1") != "0"


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def reverse_sequences(x: torch.Tensor, lengths: torch.Tensor | None = None) -> torch.Tensor:
    """Reverse each row's valid span, padding left in place (mamba_block.py:19-28).  Kept for API parity;
    the block itself never calls it (the kernels index through the same map)."""
    if lengths is None:
        return torch.flip(x, dims=[1])
    B, T, _ = x.shape
    pos = torch.arange(T, device=x.device).unsqueeze(0).expand(B, T)
    L = lengths.to(x.device).view(B, 1)
    idx = torch.where(pos < L, L - 1 - pos, pos).clamp_(0, T - 1)
    return torch.gather(x, 1, idx.unsqueeze(-1).expand_as(x))


class _RMSNormWeight(nn.Module):
    def __init__(self, d: int, eps: float = 1e-5):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.eps = eps


class Mamba2(nn.Module):
    """Drop-in for ``mamba_ssm.Mamba2(d_model, d_state, d_conv, expand, headdim)`` with package defaults
    (ngroups=1, rmsnorm, norm_before_gate=False, no in/out bias, conv bias, dt_limit=(0, inf))."""

    def __init__(self, d_model, d_state=128, d_conv=4, expand=2, headdim=64, ngroups=1,
                 dt_min=0.001, dt_max=0.1, dt_init_floor=1e-4, A_init_range=(1, 16), **unused):
        super().__init__()
        if ngroups != 1 or d_conv != 4:
            raise HnbError("Mamba2: kernels are built for ngroups=1, d_conv=4")
        self.d_model, self.d_state, self.d_conv, self.headdim = d_model, d_state, d_conv, headdim
        self.d_inner = expand * d_model
        if self.d_inner % headdim:
            raise HnbError("Mamba2: expand*d_model must be divisible by headdim")
        self.nheads = self.d_inner // headdim
        conv_dim = self.d_inner + 2 * d_state
        self.in_proj = nn.Linear(d_model, 2 * self.d_inner + 2 * d_state + self.nheads, bias=False)
        self.conv1d = nn.Conv1d(conv_dim, conv_dim, d_conv, groups=conv_dim, padding=d_conv - 1, bias=True)
        dt = torch.exp(torch.rand(self.nheads) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min))
        dt = dt.clamp(min=dt_init_floor)
        self.dt_bias = nn.Parameter(dt + torch.log(-torch.expm1(-dt)))
        self.dt_bias._no_weight_decay = True
        self.A_log = nn.Parameter(torch.log(torch.empty(self.nheads).uniform_(*A_init_range)))
        self.A_log._no_weight_decay = True
        self.D = nn.Parameter(torch.ones(self.nheads))
        self.D._no_weight_decay = True
        self.norm = _RMSNormWeight(self.d_inner)
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=False)

    def _params(self):
        return (self.in_proj.weight, self.conv1d.weight, self.conv1d.bias, self.dt_bias, self.A_log, self.D,
                self.norm.weight, self.out_proj.weight)

    def forward(self, u: torch.Tensor) -> torch.Tensor:
        return _MixerFn.apply(u, None, None, None, 1, self.d_inner, self.d_state, self.nheads, False,
                              *self._params())


NP = 8   # parameters per direction, in _params() order


def _mixer_forward(h2, lengths, B, L, ndir, di, N, H, params):
    """h2 [B*L, d] (activation dtype) -> (ynorm [B*L, ndir*di], saved tensors)."""
    adt = h2.dtype
    dip = 2 * di + 2 * N + H
    dstride = _round_up(dip, 8)
    d = h2.shape[1]
    Win, Wout, conv_w, conv_b, dt_bias, A_log, Dk, norm_w = ops.pack_mixer_params(
        [p_.detach() if p_.is_contiguous() else p_.detach().contiguous() for p_ in params], ndir, d, di, N, H, dstride,
        adt, h2.device)
    zx = ops.gemm(h2, Win)                                               # [B*L, ndir*dstride]
    C = di + 2 * N
    xconv, dt = ops.conv_fwd(zx, dstride, lengths, conv_w, conv_b, dt_bias, ndir, B, L, di, N, H)
    y, ws = ops.ssd_fwd(xconv, dt, A_log, Dk, ndir, B, L, di, N, H)
    yn, rstd = ops.gated_norm_fwd(y, zx, dstride, lengths, norm_w, ndir, B, L, di)
    saved = (Win, zx, conv_w, conv_b, dt_bias, A_log, Dk, norm_w, xconv, dt, y, ws, yn, rstd, Wout)
    return yn, Wout, saved, dstride


def _acc_sizes(ndir, C, di, H, ln_acc_d):
    """fp32 accumulators of one block backward: conv_w, conv_b, norm_w, dA_log, dD, ddt_bias, LayerNorm (dgamma | dbeta)."""
    return [ndir * C * 4, ndir * C, ndir * di, ndir * H, ndir * H, ndir * H, 2 * ln_acc_d]


def _mixer_backward(dyn, h2, lengths, B, L, ndir, di, N, H, dstride, saved, ln_acc_d=0, arena=None):
    """dyn = d loss / d ynorm [B*L, ndir*di] -> (dh2, per-direction parameter grads except out_proj, LN accumulator)."""
    Win, zx, conv_w, conv_b, dt_bias, A_log, Dk, norm_w, xconv, dt, y, ws, yn, rstd, Wout = saved
    dip = 2 * di + 2 * N + H
    C = di + 2 * N
    dzx = torch.empty_like(zx)
    if dstride != dip:
        dzx.view(-1, ndir, dstride)[:, :, dip:] = 0                      # pad columns feed the GEMMs below
    # one zero-filled buffer for every accumulated parameter gradient of the block (one fill instead of ten)
    # (conv accumulators first: C % 4 == 0 keeps them 16-byte aligned for the kernel's vector reductions)
    sizes = _acc_sizes(ndir, C, di, H, ln_acc_d)
    d = h2.shape[1]
    if arena is None:
        arena = torch.zeros(ndir * dstride * d + sum(sizes), dtype=torch.float32, device=zx.device)
    dWin_buf, *accs = arena[-(ndir * dstride * d + sum(sizes)):].split([ndir * dstride * d] + sizes)
    a_cw, a_cb, a_nw = accs[0].view(ndir, C, 4), accs[1].view(ndir, C), accs[2].view(ndir, di)
    a_dA, a_dD, a_dtb = accs[3].view(ndir, H), accs[4].view(ndir, H), accs[5].view(ndir, H)
    dy, dnorm_w = ops.gated_norm_bwd(dyn, y, zx, dstride, lengths, norm_w, rstd, ndir, B, L, di, dzx, acc=a_nw)
    dxc, dBC, ddt, dA, dD = ops.ssd_bwd(dy, xconv, y, dt, A_log, Dk, ws, ndir, B, L, di, N, H, acc=(a_dA, a_dD),
                                        keep_parts=True)
    dconv_w, dconv_b, ddt_bias = ops.conv_bwd(zx, dxc, dBC, ddt, dstride, lengths, conv_w, conv_b, dt_bias,
                                              ndir, B, L, di, N, H, dzx, acc=(a_cw, a_cb, a_dtb))
    dh2 = ops.gemm(dzx, Win, trans_b=True)                               # dgrad [B*L, d]
    sk = ops.wgrad_splitk(B * L, Win.shape[0], Win.shape[1]) if dzx.dtype == torch.bfloat16 else 1
    dWin = ops.gemm(dzx, h2, trans_a=True, trans_b=True, splitk=sk, out_dtype=torch.float32,
                    out=dWin_buf.view(ndir * dstride, d))                # [ndir*dstride, d], pre-zeroed (split-K accumulates)
    grads = []
    for r in range(ndir):
        grads.append((dWin[r * dstride: r * dstride + dip], dconv_w[r].reshape(-1, 1, 4), dconv_b[r], ddt_bias[r],
                      dA[r], dD[r], dnorm_w[r]))
    return dh2, grads, (accs[6].view(2, ln_acc_d) if ln_acc_d else None)


class _MixerFn(torch.autograd.Function):
    """(optional LayerNorm) -> mixer(s) -> (optional residual).  Serves Mamba2 (1 direction, no norm, no
    residual) and MambaBlock (norm + residual, 1 or 2 directions)."""

    @staticmethod
    def forward(ctx, x, lengths, ln_w, ln_b, ndir, di, N, H, block, *params):
        B, L, d = x.shape
        adt = _autocast_dtype() or x.dtype
        x2 = x.reshape(B * L, d)
        x2 = x2 if x2.is_contiguous() else x2.contiguous()
        if lengths is not None:
            lengths = lengths.to(device=x.device, dtype=torch.int32).contiguous()
        if block:
            h2, mean, rstd_ln = ops.layernorm_fwd(x2, ln_w.float(), ln_b.float(), 1e-5, adt)
        else:
            h2, mean, rstd_ln = (x2 if x2.dtype == adt else x2.to(adt)), None, None
        yn, Wout, saved, dstride = _mixer_forward(h2, lengths, B, L, ndir, di, N, H, params)
        if block:
            # residual stream keeps the dtype it arrived in (reference: x + y under autocast)
            if adt == torch.float32:
                out2 = ops.gemm(yn, Wout, residual=x2)
            elif x2.dtype == torch.float32:
                out2 = ops.gemm(yn, Wout, residual=x2, out_dtype=torch.float32)
            else:
                out2 = ops.gemm(yn, Wout, residual=x2, out_dtype=torch.bfloat16)
        else:
            out2 = ops.gemm(yn, Wout)
        ctx.save_for_backward(x2, lengths, ln_w, mean, rstd_ln, h2, *saved)
        ctx.meta = (B, L, d, ndir, di, N, H, block, dstride, x.dtype, [p.dtype for p in params])
        return out2.view(B, L, d)

    @staticmethod
    def backward(ctx, dout):
        x2, lengths, ln_w, mean, rstd_ln, h2, *saved = ctx.saved_tensors
        B, L, d, ndir, di, N, H, block, dstride, xdt, pdts = ctx.meta
        Wout, yn = saved[-1], saved[-3]
        adt = h2.dtype
        dout2 = dout.reshape(B * L, d)
        dout2 = dout2 if dout2.is_contiguous() else dout2.contiguous()
        da = dout2 if dout2.dtype == adt else dout2.to(adt)
        dyn = ops.gemm(da, Wout, trans_b=True)                           # [B*L, ndir*di]
        # ONE zero fill per block backward: the split-K weight gradients and every accumulated parameter gradient are
        # views of this arena (three fills per block before).  out_proj's gradient stays ONE GEMM over both directions:
        # a GEMM per direction (contiguous [d, di] blocks, no strided copy in AccumulateGrad) was measured 0.3 ms per
        # step slower -- half-width weight-gradient GEMMs fill the SMs worse than the copies cost.
        n_acc = sum(_acc_sizes(ndir, di + 2 * N, di, H, d if block else 0))
        arena = torch.zeros(ndir * d * di + ndir * dstride * d + n_acc, dtype=torch.float32, device=da.device)
        sk = ops.wgrad_splitk(B * L, d, ndir * di) if adt == torch.bfloat16 else 1
        dWout = ops.gemm(da, yn, trans_a=True, trans_b=True, splitk=sk, out_dtype=torch.float32,
                         out=arena[:ndir * d * di].view(d, ndir * di))   # [d, ndir*di]
        dh2, grads, ln_acc = _mixer_backward(dyn, h2, lengths, B, L, ndir, di, N, H, dstride, saved, d if block else 0,
                                             arena=arena)
        if block:
            dres = dout2 if dout2.dtype == xdt else dout2.to(xdt)
            dx2, dg, db = ops.layernorm_bwd(dh2, x2, ln_w.float(), mean, rstd_ln, dres, acc=ln_acc)
        else:
            dx2, dg, db = (dh2 if dh2.dtype == xdt else dh2.to(xdt)), None, None
        pg = []
        for r in range(ndir):
            g = list(grads[r]) + [dWout[:, r * di:(r + 1) * di]]
            pg += [t.to(pdts[r * NP + i]) for i, t in enumerate(g)]
        return (dx2.view(B, L, d), None, dg, db, None, None, None, None, None, *pg)


class _BlockFn(torch.autograd.Function):
    """One MambaBlock = ONE host call forward and ONE backward (hnb_block_fwd / hnb_block_bwd): same kernels as
    _MixerFn, sequenced inside the library in one workspace.  Host time per encoder step drops from ~13 ms to a few ms,
    which is what keeps eight ranks sharing one host from becoming launch-bound."""

    @staticmethod
    def forward(ctx, x, lengths, ln_w, ln_b, ndir, di, N, H, packed, *params):
        B, L, d = x.shape
        adt = _autocast_dtype() or x.dtype
        x2 = x.reshape(B * L, d)
        x2 = x2 if x2.is_contiguous() else x2.contiguous()
        if not x2.is_cuda:
            raise HnbError(f"MambaBlock: got a {x2.device} tensor; the hot path is CUDA-only")
        if lengths is not None:
            lengths = lengths.to(device=x.device, dtype=torch.int32).contiguous()
        ps = [p_.detach() if p_.is_contiguous() else p_.detach().contiguous() for p_ in params]
        for t in ps:
            if t.dtype != torch.float32 or t.device != x2.device:
                raise HnbError("mixer parameters must be fp32 masters on the input's device")
        ptrs = (ctypes.c_void_p * len(ps))(*[t.data_ptr() for t in ps])
        L_ = lib()
        act, xdt, impl = dtype_code(adt), dtype_code(x2.dtype), ops.ssd_impl_for(adt)
        ws = torch.empty(L_.raw("block_fwd_ws_bytes")(B, L, d, ndir, di, N, H, act), dtype=torch.uint8, device=x2.device)
        out2 = torch.empty_like(x2)
        ln_w32, ln_b32 = ln_w.detach().float().contiguous(), ln_b.detach().float().contiguous()
        L_.call("block_fwd", x2, xdt, lengths, ln_w32, ln_b32, ctypes.addressof(ptrs), B, L, d, ndir, di, N, H, act, impl,
                packed, out2, ws, stream())
        ctx.save_for_backward(x2, lengths, ln_w32, ws, packed)
        ctx.meta = (B, L, d, ndir, di, N, H, act, xdt, impl, x.dtype, [p_.dtype for p_ in params], ln_w.dtype)
        return out2.view(B, L, d)

    @staticmethod
    def backward(ctx, dout):
        x2, lengths, ln_w32, ws, packed = ctx.saved_tensors
        B, L, d, ndir, di, N, H, act, xdt, impl, x_dtype, pdts, ln_dtype = ctx.meta
        dout2 = dout.reshape(B * L, d)
        dout2 = dout2 if dout2.is_contiguous() else dout2.contiguous()
        dout2 = dout2 if dout2.dtype == x2.dtype else dout2.to(x2.dtype)
        L_ = lib()
        offs = (ctypes.c_longlong * 9)()
        n = L_.raw("block_grad_floats")(B, L, d, ndir, di, N, H, ctypes.addressof(offs))
        # the split-K weight gradients and every accumulated parameter gradient are views of this arena; the call clears it
        # on its own stream (no fill kernel, no extra host round trip)
        arena = torch.empty(n, dtype=torch.float32, device=x2.device)
        scratch = torch.empty(L_.raw("block_bwd_ws_bytes")(B, L, d, ndir, di, N, H, act, impl), dtype=torch.uint8,
                              device=x2.device)
        dx2 = torch.empty_like(x2)
        L_.call("block_bwd", dout2, x2, xdt, lengths, ln_w32, ws, B, L, d, ndir, di, N, H, act, impl, packed, dx2, arena, 1,
                scratch, stream())
        C, dip = di + 2 * N, 2 * di + 2 * N + H
        dstride = _round_up(dip, 8)
        o = list(offs)
        dWout = arena[o[0]:o[0] + d * ndir * di].view(ndir, d, di)       # one contiguous [d, di] matrix per direction
        dWin = arena[o[1]:o[1] + ndir * dstride * d].view(ndir * dstride, d)
        cw = arena[o[2]:o[2] + ndir * C * 4].view(ndir, C, 1, 4)
        cb = arena[o[3]:o[3] + ndir * C].view(ndir, C)
        nw = arena[o[4]:o[4] + ndir * di].view(ndir, di)
        dA = arena[o[5]:o[5] + ndir * H].view(ndir, H)
        dD = arena[o[6]:o[6] + ndir * H].view(ndir, H)
        dtb = arena[o[7]:o[7] + ndir * H].view(ndir, H)
        ln = arena[o[8]:o[8] + 2 * d].view(2, d)
        pg = []
        for r in range(ndir):                    # _params() order: in_proj.w, conv1d.w, conv1d.b, dt_bias, A_log, D, norm.w, out_proj.w
            g = (dWin[r * dstride: r * dstride + dip], cw[r], cb[r], dtb[r], dA[r], dD[r], nw[r],
                 dWout[r])
            pg += [t if t.dtype == pdts[r * NP + i] else t.to(pdts[r * NP + i]) for i, t in enumerate(g)]
        dg, db = ln[0], ln[1]
        if ln_dtype != torch.float32:
            dg, db = dg.to(ln_dtype), db.to(ln_dtype)
        return (dx2.view(B, L, d), None, dg, db, None, None, None, None, None, *pg)


class MambaBlock(nn.Module):
    """y = x + Mamba2_fwd(norm(x)) [+ reverse(Mamba2_bwd(reverse(norm(x))))]  (mamba_block.py:31-56)."""

    def __init__(self, d_model: int, bidirectional: bool = True, d_state: int = 128,
                 d_conv: int = 4, expand: int = 2, headdim: int = 64):
        super().__init__()
        assert (expand * d_model) % headdim == 0, \
            f"expand*d_model ({expand * d_model}) must be divisible by headdim ({headdim})"
        self.bidirectional = bidirectional
        self.norm = nn.LayerNorm(d_model)
        kw = dict(d_model=d_model, d_state=d_state, d_conv=d_conv, expand=expand, headdim=headdim)
        self.fwd = Mamba2(**kw)
        self.bwd = Mamba2(**kw) if bidirectional else None

    def forward(self, x: torch.Tensor, lengths: torch.Tensor | None = None, packed: torch.Tensor | None = None) -> torch.Tensor:
        """``packed`` (not in the reference signature): this block's slice of the weights MambaStack packed for all its
        blocks in one launch; without it the block packs its own."""
        m = self.fwd
        params = m._params() + (self.bwd._params() if self.bwd is not None else ())
        ndir = 2 if self.bwd is not None else 1
        if BLOCK_COMPOSITE and _lib._PROFILE is None:
            return _BlockFn.apply(x, lengths if ndir == 2 else None, self.norm.weight, self.norm.bias, ndir,
                                  m.d_inner, m.d_state, m.nheads, packed, *params)
        return _MixerFn.apply(x, lengths if ndir == 2 else None, self.norm.weight, self.norm.bias, ndir,
                              m.d_inner, m.d_state, m.nheads, True, *params)


class _FinalNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps):
        B, L, d = x.shape
        x2 = x.reshape(B * L, d)
        x2 = x2 if x2.is_contiguous() else x2.contiguous()
        # autocast runs layer_norm in fp32: the stack output is fp32 under bf16 autocast
        out_dtype = torch.float32 if _autocast_dtype() is not None else x.dtype
        y, mean, rstd = ops.layernorm_fwd(x2, w.float(), b.float(), eps, out_dtype)
        ctx.save_for_backward(x2, w, mean, rstd)
        ctx.meta = (B, L, d, w.dtype)
        return y.view(B, L, d)

    @staticmethod
    def backward(ctx, dy):
        x2, w, mean, rstd = ctx.saved_tensors
        B, L, d, wdt = ctx.meta
        dy2 = dy.reshape(B * L, d)
        dy2 = dy2 if dy2.is_contiguous() else dy2.contiguous()
        dx, dg, db = ops.layernorm_bwd(dy2, x2, w.float(), mean, rstd, None)
        return dx.view(B, L, d), dg.to(wdt), db.to(wdt), None


class MambaStack(nn.Module):
    """n_layers MambaBlocks + a final LayerNorm (mamba_block.py:59-73)."""

    def __init__(self, n_layers: int, d_model: int, bidirectional: bool = True, **block_kw):
        super().__init__()
        self.layers = nn.ModuleList(
            MambaBlock(d_model, bidirectional=bidirectional, **block_kw) for _ in range(n_layers))
        self.norm = nn.LayerNorm(d_model)

    def forward(self, x: torch.Tensor, lengths: torch.Tensor | None = None) -> torch.Tensor:
        if lengths is not None and lengths.dtype != torch.int32:
            lengths = lengths.to(torch.int32)
        packed = self._pack_all(x) if (STACK_PACK and BLOCK_COMPOSITE and _lib._PROFILE is None and x.is_cuda) else None
        for i, layer in enumerate(self.layers):
            x = layer(x, lengths) if packed is None else layer(x, lengths, packed=packed[i])
        return _FinalNormFn.apply(x, self.norm.weight, self.norm.bias, self.norm.eps)

    def _pack_all(self, x: torch.Tensor):
        """The weights of EVERY block of the stack cast to the activation dtype and stacked per direction in ONE launch
        (hnb_pack_mixer_stack): 20 per-block launches of ~18 us per encoder step were launch latency, not traffic."""
        blocks = list(self.layers)
        if not blocks:
            return None
        m0 = blocks[0].fwd
        ndir = 2 if blocks[0].bwd is not None else 1
        ps = []
        for blk in blocks:
            for mix in ((blk.fwd, blk.bwd) if ndir == 2 else (blk.fwd,)):
                for t in mix._params():
                    t = t.detach()
                    if t.dtype != torch.float32 or t.device != x.device:
                        return None                      # (unusual parameter dtypes: let each block validate and pack)
                    ps.append(t if t.is_contiguous() else t.contiguous())
        adt = _autocast_dtype() or x.dtype
        L_ = lib()
        act = dtype_code(adt)
        per = int(L_.raw("block_packed_bytes")(m0.d_model, ndir, m0.d_inner, m0.d_state, m0.nheads, act))
        buf = torch.empty((len(blocks), per), dtype=torch.uint8, device=x.device)
        ptrs = (ctypes.c_void_p * len(ps))(*[t.data_ptr() for t in ps])
        L_.call("pack_mixer_stack", ctypes.addressof(ptrs), len(blocks), ndir, m0.d_model, m0.d_inner, m0.d_state, m0.nheads, act,
                buf, stream())
        self._pack_keepalive = ps                        # contiguous copies (if any) must outlive the launch
        return buf
