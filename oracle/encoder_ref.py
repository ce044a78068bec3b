"""Oracle: CPU restatement of the reference encoder assembly (test infrastructure).

  reverse_ref        <- reverse_sequences     src/dcasr/models/mamba_block.py:19-28
  MambaBlockRef      <- MambaBlock            src/dcasr/models/mamba_block.py:31-56
  MambaStackRef      <- MambaStack            src/dcasr/models/mamba_block.py:59-73
  DynamicChunkerRef  <- DynamicChunker        src/dcasr/models/hnet_chunk.py:142-252
  EncoderRef         <- DCASREncoder          src/dcasr/models/encoder.py:77-144

Module tree and state_dict keys equal the reference's, so one state_dict drives the
reference (when importable), this oracle, and the CUDA product.  The Mamba-2 mixer is
``Mamba2Ref`` (parity unpinned by the reference, see oracle/__init__.py); everything
else is pinned through tests/golden/ vectors produced by the reference's own modules.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn as nn

from . import hnet_ref
from .mamba2_ref import Mamba2Ref


def reverse_ref(x, lengths=None):
    if lengths is None:
        return x.flip(1)
    out = x.clone()
    for i in range(x.shape[0]):
        n = int(lengths[i])
        out[i, :n] = x[i, :n].flip(0)
    return out


class MambaBlockRef(nn.Module):
    def __init__(self, d_model, bidirectional=True, d_state=128, d_conv=4, expand=2, headdim=64):
        super().__init__()
        assert (expand * d_model) % headdim == 0
        self.norm = nn.LayerNorm(d_model)
        kw = dict(d_model=d_model, d_state=d_state, d_conv=d_conv, expand=expand, headdim=headdim)
        self.fwd = Mamba2Ref(**kw)
        self.bwd = Mamba2Ref(**kw) if bidirectional else None

    def forward(self, x, lengths=None):
        h = self.norm(x)
        y = self.fwd(h)
        if self.bwd is not None:
            y = y + reverse_ref(self.bwd(reverse_ref(h, lengths)), lengths)
        return x + y


class MambaStackRef(nn.Module):
    def __init__(self, n_layers, d_model, bidirectional=True, **kw):
        super().__init__()
        self.layers = nn.ModuleList(MambaBlockRef(d_model, bidirectional, **kw) for _ in range(n_layers))
        self.norm = nn.LayerNorm(d_model)

    def forward(self, x, lengths=None):
        for layer in self.layers:
            x = layer(x, lengths)
        return self.norm(x)


class _Router(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.W_q = nn.Linear(d, d, bias=False)
        self.W_k = nn.Linear(d, d, bias=False)
        nn.init.eye_(self.W_q.weight)
        nn.init.eye_(self.W_k.weight)


class DynamicChunkerRef(nn.Module):
    def __init__(self, d_model, N=1, ema_smoothing=True):
        super().__init__()
        self.N, self.ema_smoothing, self.identity = N, ema_smoothing, (N == 1)
        self.router = None if self.identity else _Router(d_model)

    def chunk(self, x, mask=None):
        Bsz, L, _ = x.shape
        if self.identity:
            ones = x.new_ones(Bsz, L)
            if mask is not None:
                ones = ones * mask.to(x.dtype)
            return hnet_ref.ChunkRef(x, mask if mask is not None else torch.ones(Bsz, L, dtype=torch.bool, device=x.device),
                                     ones, ones, torch.arange(L, device=x.device).expand(Bsz, L).clone(),
                                     x.new_zeros(()), x.new_ones(()))
        return hnet_ref.chunk_vec(x, self.router.W_q.weight, self.router.W_k.weight, self.N, mask)

    def dechunk(self, z_proc, co):
        if self.identity:
            return z_proc
        return hnet_ref.dechunk_vec(z_proc, co, self.ema_smoothing)


class _Subsample(nn.Module):
    """ConvSubsampling4 (src/dcasr/models/encoder.py:55-70): library conv/linear, outside the hot path."""

    def __init__(self, n_mels, d_model):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(1, d_model, 3, 2), nn.ReLU(),
                                  nn.Conv2d(d_model, d_model, 3, 2), nn.ReLU())
        self.proj = nn.Linear(d_model * (((n_mels - 1) // 2 - 1) // 2), d_model)

    def forward(self, feats, lengths):
        x = self.conv(feats.unsqueeze(1))
        Bsz, C, T, Fq = x.shape
        x = self.proj(x.transpose(1, 2).reshape(Bsz, T, C * Fq))
        return x, (((lengths - 1) // 2 - 1) // 2).clamp_min(0)


class FixedPoolChunkerRef(nn.Module):
    """The reference's fixed-stride control (src/dcasr/models/fixed_pool.py:31-110) behind the chunker interface the
    encoder uses; the arithmetic is oracle/fixed_pool_ref.py."""

    def __init__(self, d_model, N=1, ema_smoothing=True):
        super().__init__()
        self.stride = int(round(float(N)))
        if abs(float(N) - self.stride) > 1e-6 or self.stride < 1:
            raise ValueError(f"fixed-stride pooling needs an integer stride >= 1, got N={N!r}")

    def chunk(self, x, mask=None):
        from . import fixed_pool_ref as fp
        if self.stride == 1:                                       # identity passthrough (:56-67)
            B, L, _ = x.shape
            ones = x.new_ones(B, L) if mask is None else mask.to(x.dtype)
            return fp.FixedChunkRef(z=x, z_mask=torch.ones(B, L, dtype=torch.bool) if mask is None else mask, p=ones,
                                    b=ones, membership=torch.arange(L).repeat(B, 1), ratio_loss=x.new_zeros(()),
                                    kept_fraction=x.new_ones(()), cnt=ones)
        return fp.chunk_ref(x, self.stride, mask)

    def dechunk(self, z_proc, co):
        from . import fixed_pool_ref as fp
        return z_proc if self.stride == 1 else fp.dechunk_ref(z_proc, co.membership)


@dataclass
class EncoderOutRef:
    features: torch.Tensor
    lengths: torch.Tensor
    ratio_loss: torch.Tensor
    boundaries: list
    chunk_embeddings: list
    kept_fractions: list


class EncoderRef(nn.Module):
    def __init__(self, n_mels=80, d_outer=384, d_main=512, n_enc=4, n_main=12, n_dec=4, n_mid=4,
                 arch_type="A", N=1, bidirectional=True, hnet_ema=True, chunker="dynamic"):
        super().__init__()
        if arch_type not in ("A", "B"):
            raise ValueError(f"arch_type must be 'A' or 'B', got {arch_type!r}")
        if chunker not in ("dynamic", "fixed"):
            raise ValueError(f"unknown chunker {chunker!r}")                 # src/dcasr/models/encoder.py:35-36
        mk = DynamicChunkerRef if chunker == "dynamic" else FixedPoolChunkerRef
        self.arch_type = arch_type
        self.subsample = _Subsample(n_mels, d_outer)
        self.enc = MambaStackRef(n_enc, d_outer, bidirectional)
        self.dec = MambaStackRef(n_dec, d_outer, bidirectional)
        if arch_type == "A":
            self.chunk = mk(d_outer, N, hnet_ema)
            self.proj_in = nn.Linear(d_outer, d_main)
            self.main = MambaStackRef(n_main, d_main, bidirectional)
            self.proj_out = nn.Linear(d_main, d_outer)
        else:
            nb = math.sqrt(N)
            self.chunk1 = mk(d_outer, nb, hnet_ema)
            self.proj1_in = nn.Linear(d_outer, d_main)
            self.mid = MambaStackRef(n_mid, d_main, bidirectional)
            self.chunk2 = mk(d_main, nb, hnet_ema)
            self.main = MambaStackRef(n_main, d_main, bidirectional)
            self.mid_dec = MambaStackRef(n_mid, d_main, bidirectional)
            self.proj1_out = nn.Linear(d_main, d_outer)

    def forward(self, feats, feat_lengths):
        x, lengths = self.subsample(feats, feat_lengths)
        return self.forward_from_subsampled(x, lengths)

    def forward_from_subsampled(self, x, lengths):
        """The hot path proper: everything after ConvSubsampling4."""
        mask = torch.arange(x.shape[1], device=x.device)[None, :] < lengths[:, None]
        x_enc = self.enc(x, lengths)
        if self.arch_type == "A":
            co = self.chunk.chunk(x_enc, mask)
            z = self.main(self.proj_in(co.z), co.z_mask.sum(1))
            x_dech = self.chunk.dechunk(self.proj_out(z), co)
            out = self.dec(x_enc + x_dech, lengths)
            return EncoderOutRef(out, lengths, co.ratio_loss, [(co.p, co.b)], [co.z], [co.kept_fraction])
        co1 = self.chunk1.chunk(x_enc, mask)
        n1 = co1.z_mask.sum(1)
        z1 = self.mid(self.proj1_in(co1.z), n1)
        co2 = self.chunk2.chunk(z1, co1.z_mask)
        z2 = self.main(co2.z, co2.z_mask.sum(1))
        z1d = self.mid_dec(z1 + self.chunk2.dechunk(z2, co2), n1)
        x_dech = self.chunk1.dechunk(self.proj1_out(z1d), co1)
        out = self.dec(x_enc + x_dech, lengths)
        return EncoderOutRef(out, lengths, co1.ratio_loss + co2.ratio_loss,
                             [(co1.p, co1.b), (co2.p, co2.b)], [co1.z, co2.z],
                             [co1.kept_fraction, co2.kept_fraction])
