"""CTC head on the CUDA path — host-side mirror of ``dcasr.decoders.ctc.CTCHead``
(/root/reference/src/dcasr/decoders/ctc.py:85-127; the step AFTER the encoder in training).

Same constructor, ``state_dict`` (``proj.weight``, ``proj.bias``), methods and results:
    forward(features) -> logits           the projection on the tcgen05 GEMM (bias in its epilogue)
    log_probs(features)                   fp32 log-softmax
    loss(features, feat_lengths, targets, target_lengths, reduction)    == F.ctc_loss(log_probs^T, ..., zero_infinity=True)
    frame_argmax / greedy_decode
``loss`` never materialises the fp32 logits or log-probabilities: hnb_ctc_lse reads the logits once, the alpha / beta
recursions gather the 2U+1 values per frame they need, hnb_ctc_grad writes d logits in the logits' dtype.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._lib import HnbError, dtype_code, lib, stream
from .hnet_chunk import _autocast_dtype


def ctc_greedy_collapse(frame_ids, blank_id: int):
    """Collapse consecutive duplicates, then drop blanks (ctc.py:70-83)."""
    out, prev = [], None
    for s in frame_ids:
        if s != prev:
            if s != blank_id:
                out.append(s)
            prev = s
    return out


def _pad_targets(targets: torch.Tensor, target_lengths: torch.Tensor, B: int) -> torch.Tensor:
    """[B, U] int64 on the device; the 1-D concatenated form F.ctc_loss also accepts is unpacked row by row."""
    if targets.dim() == 2:
        return targets.to(torch.int64).contiguous()
    lens = [int(v) for v in target_lengths.tolist()]
    U = max(lens + [1])
    out = targets.new_zeros((B, U), dtype=torch.int64)
    o = 0
    for i, n in enumerate(lens):
        out[i, :n] = targets[o:o + n]
        o += n
    return out


def _ld(v1: int) -> int:
    """Row stride of the logits buffer: V+1 rounded up to 16 elements, so that rows are 32-byte aligned (vector stores in
    the GEMM epilogue) and the gradient matrix is a legal TMA operand (16-byte row strides) for dgrad / wgrad."""
    return (v1 + 15) // 16 * 16


class _HeadProjFn(torch.autograd.Function):
    """logits = features W^T + b as a [B, L, V1] view of a [B*L, ld] buffer (ld = _ld(V1)); the projection, its dgrad and
    wgrad run on the library's GEMMs (tcgen05 under bf16 autocast), the bias gradient is one column-sum kernel."""

    @staticmethod
    def forward(ctx, x, w, b):
        adt = _autocast_dtype() or x.dtype
        B, L, d = x.shape
        V1 = w.shape[0]
        x2 = x.reshape(B * L, d)
        xa = (x2 if x2.dtype == adt else x2.to(adt)).contiguous()
        wa = w.to(adt).contiguous()
        buf = torch.empty((B * L, _ld(V1)), dtype=adt, device=x.device)
        ops.gemm(xa, wa, bias=b.float() if b is not None else None, out=buf[:, :V1])
        ctx.save_for_backward(xa, wa)
        ctx.meta = (x.shape, x.dtype, w.dtype, b.dtype if b is not None else None)
        return buf.view(B, L, -1)[:, :, :V1]

    @staticmethod
    def backward(ctx, dy):
        xa, wa = ctx.saved_tensors
        shp, xdt, wdt, bdt = ctx.meta
        V1 = wa.shape[0]
        d2 = dy.reshape(-1, V1)
        if d2.dtype != xa.dtype or d2.stride(1) != 1 or d2.stride(0) % 16 != 0 or d2.data_ptr() % 32 != 0:
            buf = torch.empty((d2.shape[0], _ld(V1)), dtype=xa.dtype, device=d2.device)
            buf[:, :V1].copy_(d2)
            d2 = buf[:, :V1]
        dx = ops.gemm(d2, wa, trans_b=True, out_dtype=xdt if xa.dtype == torch.bfloat16 else None)
        sk = ops.wgrad_splitk(xa.shape[0], V1, wa.shape[1]) if xa.dtype == torch.bfloat16 else 1
        dw = ops.gemm(d2, xa, trans_a=True, trans_b=True, splitk=sk, out_dtype=torch.float32)
        db = None
        if bdt is not None:
            db = ops.col_sum(d2).to(bdt)
        return dx.to(xdt).view(shp), dw.to(wdt), db


class _CTCLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, feat_lengths, targets, target_lengths, blank, reduction):
        if not logits.is_cuda:
            raise HnbError(f"CTCHead.loss: got a {logits.device} tensor; the CUDA path has no CPU fallback")
        B, T, V1 = logits.shape
        lg = logits                                          # [B, T, V1] rows of stride ldl (the padded projection buffer)
        if lg.stride(2) != 1 or lg.stride(0) != T * lg.stride(1) or lg.stride(1) < V1:
            lg = lg.contiguous()
        ldl = lg.stride(1)
        dev = lg.device
        fl = feat_lengths.to(device=dev, dtype=torch.int64).contiguous()
        tl = target_lengths.to(device=dev, dtype=torch.int64).contiguous()
        tg = _pad_targets(targets.to(dev), tl, B)
        U = tg.shape[1]
        if tg.numel() == 0:                                  # all-empty targets: keep a valid pointer
            tg = torch.zeros((B, 1), dtype=torch.int64, device=dev)
            U = 0
        S = 2 * U + 1
        L_ = lib()
        dt = dtype_code(lg.dtype)
        lse = torch.empty((B, T), dtype=torch.float32, device=dev)
        alpha = torch.empty((B, T, S), dtype=torch.float32, device=dev)
        beta = torch.empty((B, T, S), dtype=torch.float32, device=dev)
        nll = torch.empty((B,), dtype=torch.float32, device=dev)
        L_.call("ctc_lse", lg.data_ptr(), dt, B * T, V1, ldl, lse, None, stream())
        L_.call("ctc_alpha_beta", lg.data_ptr(), dt, lse, tg, fl, tl, B, T, V1, ldl, U, int(blank), alpha, beta, nll, stream())
        losses = torch.where(torch.isinf(nll), torch.zeros_like(nll), nll)        # zero_infinity=True
        ctx.save_for_backward(lg, lse, alpha, beta, tg, fl, tl, nll)
        ctx.meta = (B, T, V1, U, int(blank), reduction, ldl)
        if reduction == "mean":
            return (losses / tl.clamp_min(1).to(losses.dtype)).mean()
        if reduction == "sum":
            return losses.sum()
        if reduction == "none":
            return losses
        raise ValueError(f"{reduction} is not a valid value for reduction")

    @staticmethod
    def backward(ctx, g):
        lg, lse, alpha, beta, tg, fl, tl, nll = ctx.saved_tensors
        B, T, V1, U, blank, reduction, ldl = ctx.meta
        g = g.to(torch.float32)
        if reduction == "mean":
            gscale = g / (tl.clamp_min(1).to(torch.float32) * B)
        elif reduction == "sum":
            gscale = g.expand(B)
        else:
            gscale = g
        gscale = gscale.contiguous()
        ldd = _ld(V1)                                        # d logits in the padded layout the projection's GEMMs take as is
        buf = torch.empty((B * T, ldd), dtype=lg.dtype, device=lg.device)
        lib().call("ctc_grad", lg.data_ptr(), dtype_code(lg.dtype), lse, alpha, beta, tg, fl, tl, nll, gscale, B, T, V1, ldl, U,
                   blank, buf, ldd, stream())
        return buf.view(B, T, ldd)[:, :, :V1], None, None, None, None, None


class CTCHead(nn.Module):
    """Linear d_model -> vocab_size+1 CTC head (blank appended at id vocab_size), ctc.py:85-127."""

    def __init__(self, d_model: int, vocab_size: int, blank_id: int | None = None):
        super().__init__()
        self.vocab_size = vocab_size
        self.blank_id = vocab_size if blank_id is None else blank_id
        self.num_classes = vocab_size + 1
        self.proj = nn.Linear(d_model, self.num_classes)

    def forward(self, features: torch.Tensor) -> torch.Tensor:
        """features [B, L, d_model] -> logits [B, L, vocab_size+1]."""
        if features.dim() != 3:
            features = features.reshape(-1, 1, features.shape[-1]) if features.dim() == 2 else features.reshape(
                features.shape[0], -1, features.shape[-1])
        return _HeadProjFn.apply(features, self.proj.weight, self.proj.bias)

    def log_probs(self, features: torch.Tensor) -> torch.Tensor:
        """Log-softmax over classes in fp32, [B, L, V+1] (decode paths; the training loss does not materialise it)."""
        logits = self.forward(features)
        B, L, V1 = logits.shape
        lse = torch.empty((B, L), dtype=torch.float32, device=logits.device)
        lib().call("ctc_lse", logits.data_ptr(), dtype_code(logits.dtype), B * L, V1, logits.stride(1), lse, None, stream())
        return logits.float() - lse.unsqueeze(-1)

    def loss(self, features: torch.Tensor, feat_lengths: torch.Tensor, targets: torch.Tensor,
             target_lengths: torch.Tensor, reduction: str = "mean") -> torch.Tensor:
        """CTC loss with zero_infinity=True; targets [B, U] (padding ignored) or 1-D concatenated."""
        return _CTCLossFn.apply(self.forward(features), feat_lengths, targets, target_lengths, self.blank_id, reduction)

    @torch.no_grad()
    def frame_argmax(self, features: torch.Tensor) -> torch.Tensor:
        """Per-frame top class incl. blank, [B, L] int64 — the raw CTC spikes."""
        logits = self.forward(features)
        B, L, V1 = logits.shape
        lse = torch.empty((B, L), dtype=torch.float32, device=logits.device)
        am = torch.empty((B, L), dtype=torch.int32, device=logits.device)
        lib().call("ctc_lse", logits.data_ptr(), dtype_code(logits.dtype), B * L, V1, logits.stride(1), lse, am, stream())
        return am.to(torch.int64)

    @torch.no_grad()
    def greedy_decode(self, features: torch.Tensor, feat_lengths: torch.Tensor) -> list[list[int]]:
        preds = self.frame_argmax(features)
        return [ctc_greedy_collapse(preds[i, :n].tolist(), self.blank_id)
                for i, n in enumerate(feat_lengths.tolist())]
