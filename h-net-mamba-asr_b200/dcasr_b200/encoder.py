"""DC-ASR encoder assembly over the CUDA hot path — host-side mirror of ``dcasr.models.encoder``
(/root/reference/src/dcasr/models/encoder.py: EncoderOutput :40-47, ConvSubsampling4 :55-70,
DCASREncoder :77-144, build_chunker :33-37).  Same constructor, module tree / ``state_dict`` keys,
forward signature, output contract and exceptions, so ``build_model`` / ``scripts/train.py`` /
``scripts/decode.py`` of the reference run on it unchanged.

In scope: the Mamba stacks, the chunk stage(s), proj_in/proj_out and the residual.
``ConvSubsampling4`` stays on cuDNN/cuBLAS through torch: north_star does not name it (§8f "next").
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn as nn

from . import ops
from .fixed_pool import FixedPoolChunker
from .hnet_chunk import DynamicChunker, _autocast_dtype
from .mamba_block import MambaStack

_CHUNKERS = {"dynamic": DynamicChunker, "fixed": FixedPoolChunker}


def register_chunker(name: str, cls) -> None:
    """Extension seam for the reference's other chunkers (e.g. its pure-torch FixedPoolChunker control)."""
    _CHUNKERS[str(name).lower()] = cls


def build_chunker(kind: str, d_model: int, N, ema_smoothing: bool = True) -> nn.Module:
    kind = str(kind).lower()
    if kind not in _CHUNKERS:
        raise ValueError(f"unknown chunker {kind!r}; choices: {sorted(_CHUNKERS)}")
    return _CHUNKERS[kind](d_model, N, ema_smoothing=ema_smoothing)


@dataclass
class EncoderOutput:
    features: torch.Tensor
    lengths: torch.Tensor
    ratio_loss: torch.Tensor
    boundaries: list
    chunk_embeddings: list
    kept_fractions: list


def _subsampled_length(lengths: torch.Tensor) -> torch.Tensor:
    return (((lengths - 1) // 2 - 1) // 2).clamp_min(0)


class _Conv1ReluFn(torch.autograd.Function):
    """relu(Conv2d(1 -> C, k3, s2)(x)) of ConvSubsampling4 as ONE kernel each way (csrc/subsample_kernels.cu): bf16 result
    written once, directly in the NHWC layout the tensor-core kernels of the second convolution read; the backward
    masks with the saved output (a1 > 0, as threshold_backward does) and reduces dW / db on the fly.  The input features need no gradient."""

    @staticmethod
    def forward(ctx, feats, w, b):
        f32 = feats.float().contiguous()
        a1 = ops.subsample_conv1_fwd(f32, w.detach().float().contiguous(), b.detach().float().contiguous())
        ctx.save_for_backward(f32, a1)            # a1 is alive anyway: it is the second convolution's saved input
        ctx.meta = (w.dtype, b.dtype)
        return a1

    @staticmethod
    def backward(ctx, dout):
        f32, a1 = ctx.saved_tensors
        d = dout if dout.dtype == torch.bfloat16 else dout.to(torch.bfloat16)
        d = d.contiguous(memory_format=torch.channels_last)
        dw, db = ops.subsample_conv1_bwd(f32, a1, d)
        return None, dw.to(ctx.meta[0]), db.to(ctx.meta[1])


class _BiasReluFn(torch.autograd.Function):
    """relu(x + bias) on the second convolution's NHWC output, in place; backward masks and reduces db in one pass."""

    @staticmethod
    def forward(ctx, x, bias):                       # x: bf16 [B, C, T, F], channels_last, fresh conv output
        B, C, T, F = x.shape
        x2 = x.permute(0, 2, 3, 1).reshape(B * T * F, C)            # a view of the NHWC storage
        ops.bias_relu_fwd_(x2, bias.detach().float().contiguous())
        ctx.mark_dirty(x)
        ctx.save_for_backward(x)
        ctx.bdtype = bias.dtype
        return x

    @staticmethod
    def backward(ctx, dout):
        (out,) = ctx.saved_tensors
        B, C, T, F = out.shape
        d = dout if dout.dtype == torch.bfloat16 else dout.to(torch.bfloat16)
        d = d.contiguous(memory_format=torch.channels_last)
        dpre, db = ops.bias_relu_bwd(d.permute(0, 2, 3, 1).reshape(B * T * F, C), out.permute(0, 2, 3, 1).reshape(B * T * F, C))
        return dpre.view(B, T, F, C).permute(0, 3, 1, 2), db.to(ctx.bdtype)


class ConvSubsampling4(nn.Module):
    """x4 time downsample (two Conv2d k3 s2 + ReLU, then Linear) -- reference encoder.py:55-70; same module tree and
    state_dict.  Under CUDA bf16 autocast (the training configuration) the one-input-channel first convolution + ReLU
    runs as the fused kernels above and everything downstream stays NHWC; the second convolution (803 GFLOP of
    implicit GEMM per step) and the Linear are library kernels (cuDNN / cuBLAS).  Any other setting (fp32 decode, CPU)
    takes the reference's own op sequence."""

    def __init__(self, n_mels: int, d_model: int):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(1, d_model, kernel_size=3, stride=2), nn.ReLU(),
            nn.Conv2d(d_model, d_model, kernel_size=3, stride=2), nn.ReLU())
        self.proj = nn.Linear(d_model * (((n_mels - 1) // 2 - 1) // 2), d_model)
        self.fused_front_end = True          # False: always the reference op sequence (parity tests)

    def forward(self, feats: torch.Tensor, lengths: torch.Tensor):
        c1, c2 = self.conv[0], self.conv[2]
        fused = (self.fused_front_end and feats.is_cuda and _autocast_dtype() == torch.bfloat16 and feats.shape[1] >= 3 and feats.shape[2] >= 3
                 and not feats.requires_grad and ops.subsample_conv1_supported(c1.out_channels))
        if fused:
            a1 = _Conv1ReluFn.apply(feats, c1.weight, c1.bias)             # bf16, channels_last
            w2 = c2.weight.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            x = torch.nn.functional.conv2d(a1, w2, None, stride=2)
            if x.is_contiguous(memory_format=torch.channels_last) and x.shape[1] % 8 == 0 and x.shape[1] <= 2048:
                x = _BiasReluFn.apply(x, c2.bias)
            else:
                x = torch.relu(x + c2.bias.to(x.dtype).view(1, -1, 1, 1))
        elif (self.fused_front_end and feats.is_cuda and feats.dtype == torch.float32 and _autocast_dtype() is None
              and not torch.is_grad_enabled() and feats.shape[1] >= 3 and feats.shape[2] >= 3
              and c1.weight.dtype == torch.float32 and c1.out_channels % 8 == 0 and ops.subsample_conv1_supported(c1.out_channels)):
            # fp32 no_grad forward (decoding, reference tasks/decode_task.py:123-151): the first convolution + ReLU as one kernel
            # writing fp32 NHWC, the second convolution by cuDNN on that memory (bias + ReLU in place by one kernel; TF32 as the process
            # has it set, exactly as for the reference's own call), the Linear on the NHWC memory as it lies through the
            # fp32-accurate tensor-core GEMM: 7.8 -> ~3 ms of a 32 ms decode step (no NCHW <-> NHWC copies, no SIMT SGEMM)
            a1 = ops.subsample_conv1_fwd(feats.contiguous(), c1.weight.detach().contiguous(), c1.bias.detach().contiguous(),
                                         out_dtype=torch.float32)
            w2 = c2.weight.detach().contiguous(memory_format=torch.channels_last)
            x = torch.nn.functional.conv2d(a1, w2, None, stride=2)
            if x.is_contiguous(memory_format=torch.channels_last):
                B, C, T, F = x.shape
                ops.bias_relu_fwd_(x.permute(0, 2, 3, 1).reshape(B * T * F, C), c2.bias.detach().float().contiguous())
                O = self.proj.out_features
                w3 = self.proj.weight.detach().view(O, C, F).transpose(1, 2).reshape(O, F * C).contiguous()
                y = ops.gemm(x.permute(0, 2, 3, 1).reshape(B * T, F * C), w3, bias=self.proj.bias.detach().float().contiguous())
                return y.view(B, T, O), _subsampled_length(lengths)
            x = torch.relu_(x + c2.bias.view(1, -1, 1, 1))
        else:
            x = self.conv(feats.unsqueeze(1))
        B, C, T, F = x.shape
        if fused and x.is_contiguous(memory_format=torch.channels_last):
            # NHWC memory is already [B, T, F, C]: flatten it as it lies and permute the (2.8 M element) weight's
            # columns instead of copying the (116 M element) activation into channel-major order, forward and backward
            O = self.proj.out_features
            w3 = self.proj.weight.view(O, C, F).transpose(1, 2).reshape(O, F * C)
            y = torch.nn.functional.linear(x.permute(0, 2, 3, 1).reshape(B, T, F * C), w3, self.proj.bias)
            return y, _subsampled_length(lengths)
        return self.proj(x.transpose(1, 2).reshape(B, T, C * F)), _subsampled_length(lengths)


class _LinearFn(torch.autograd.Function):
    """nn.Linear on the tcgen05 GEMM (proj_in / proj_out): y = x W^T + b over [B, L, d_in]."""

    @staticmethod
    def forward(ctx, x, w, b):
        adt = _autocast_dtype() or x.dtype
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        xa = (x2 if x2.dtype == adt else x2.to(adt)).contiguous()
        wa = w.to(adt)
        y = ops.gemm(xa, wa, bias=b.float() if b is not None else None)
        ctx.save_for_backward(xa, wa)
        ctx.meta = (shp, x.dtype, w.dtype, b is not None)
        return y.view(*shp[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        xa, wa = ctx.saved_tensors
        shp, xdt, wdt, has_b = ctx.meta
        d2 = dy.reshape(-1, dy.shape[-1])
        da = (d2 if d2.dtype == xa.dtype else d2.to(xa.dtype)).contiguous()
        dx = ops.gemm(da, wa, trans_b=True, out_dtype=xdt if xa.dtype == torch.bfloat16 else None)
        sk = ops.wgrad_splitk(xa.shape[0], wa.shape[0], wa.shape[1]) if xa.dtype == torch.bfloat16 else 1
        dw = ops.gemm(da, xa, trans_a=True, trans_b=True, splitk=sk, out_dtype=torch.float32)
        db = ops.col_sum(da) if has_b else None                     # bias gradient: one column-sum kernel over d y
        return dx.to(xdt).view(shp), dw.to(wdt), (db.to(wdt) if has_b else None)


def _linear(mod: nn.Linear, x: torch.Tensor) -> torch.Tensor:
    return _LinearFn.apply(x, mod.weight, mod.bias)


class DCASREncoder(nn.Module):
    """Type A (1-stage) or Type B (2-stage) Mamba-H-Net encoder (encoder.py:77-144)."""

    def __init__(self, n_mels: int = 80, d_outer: int = 384, d_main: int = 512,
                 n_enc: int = 4, n_main: int = 12, n_dec: int = 4, n_mid: int = 4,
                 arch_type: str = "A", N: int = 1, bidirectional: bool = True,
                 hnet_ema: bool = True, chunker: str = "dynamic"):
        super().__init__()
        if arch_type not in ("A", "B"):
            raise ValueError(f"arch_type must be 'A' or 'B', got {arch_type!r}")
        self.arch_type, self.N, self.chunker = arch_type, N, chunker
        self.subsample = ConvSubsampling4(n_mels, d_outer)
        self.enc = MambaStack(n_enc, d_outer, bidirectional)
        self.dec = MambaStack(n_dec, d_outer, bidirectional)
        if arch_type == "A":
            self.chunk = build_chunker(chunker, d_outer, N, hnet_ema)
            self.proj_in = nn.Linear(d_outer, d_main)
            self.main = MambaStack(n_main, d_main, bidirectional)
            self.proj_out = nn.Linear(d_main, d_outer)
        else:
            nb = math.sqrt(N)
            self.chunk1 = build_chunker(chunker, d_outer, nb, hnet_ema)
            self.proj1_in = nn.Linear(d_outer, d_main)
            self.mid = MambaStack(n_mid, d_main, bidirectional)
            self.chunk2 = build_chunker(chunker, d_main, nb, hnet_ema)
            self.main = MambaStack(n_main, d_main, bidirectional)
            self.mid_dec = MambaStack(n_mid, d_main, bidirectional)
            self.proj1_out = nn.Linear(d_main, d_outer)

    def forward(self, feats: torch.Tensor, feat_lengths: torch.Tensor) -> EncoderOutput:
        x, lengths = self.subsample(feats, feat_lengths)
        return self.forward_hot_path(x, lengths)

    def forward_hot_path(self, x: torch.Tensor, lengths: torch.Tensor) -> EncoderOutput:
        """Everything after ConvSubsampling4: the path north_star names."""
        mask = torch.arange(x.shape[1], device=lengths.device)[None, :] < lengths[:, None]
        l32 = lengths.to(torch.int32)
        x_enc = self.enc(x, l32)
        if self.arch_type == "A":
            return self._forward_A(x_enc, mask, lengths, l32)
        return self._forward_B(x_enc, mask, lengths, l32)

    @staticmethod
    def _dechunk_add(chunker, z, co, resid):
        if isinstance(chunker, (DynamicChunker, FixedPoolChunker)):
            return chunker.dechunk(z, co, residual=resid)        # fused gather (+ STE) + residual
        return resid + chunker.dechunk(z, co)

    def _forward_A(self, x_enc, mask, lengths, l32) -> EncoderOutput:
        co = self.chunk.chunk(x_enc, mask)
        z = _linear(self.proj_in, co.z)
        z = self.main(z, co.z_mask.sum(1))
        z = _linear(self.proj_out, z)
        x_out = self.dec(self._dechunk_add(self.chunk, z, co, x_enc), l32)
        return EncoderOutput(x_out, lengths, co.ratio_loss, [(co.p, co.b)], [co.z], [co.kept_fraction])

    def _forward_B(self, x_enc, mask, lengths, l32) -> EncoderOutput:
        co1 = self.chunk1.chunk(x_enc, mask)
        n1 = co1.z_mask.sum(1)
        z1 = self.mid(_linear(self.proj1_in, co1.z), n1)
        co2 = self.chunk2.chunk(z1, co1.z_mask)
        z2 = self.main(co2.z, co2.z_mask.sum(1))
        z1_dec = self.mid_dec(self._dechunk_add(self.chunk2, z2, co2, z1), n1)
        x_dech_in = _linear(self.proj1_out, z1_dec)
        x_out = self.dec(self._dechunk_add(self.chunk1, x_dech_in, co1, x_enc), l32)
        return EncoderOutput(x_out, lengths, co1.ratio_loss + co2.ratio_loss,
                             [(co1.p, co1.b), (co2.p, co2.b)], [co1.z, co2.z],
                             [co1.kept_fraction, co2.kept_fraction])
