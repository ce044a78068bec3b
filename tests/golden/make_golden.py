"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container only (it needs /root/reference, which does not exist on the
GPU box):   python tests/golden/make_golden.py

* hnet_*.npz  — /root/reference/src/dcasr/models/hnet_chunk.py executed unmodified on CPU
  (fp32) on seeded inputs: router / ratio loss / chunk / dechunk values, the integer
  outputs, and autograd gradients.
* ema_*.npz   — DynamicChunker._ema values and gradients, including saturated p.
* enc_*.npz   — the reference's own encoder.py + mamba_block.py (unmodified) with the
  missing third-party ``mamba_ssm`` module supplied by oracle/mamba2_ref.py.  These pin the
  assembly (block wiring, length-aware reversal, Type A / Type B orchestration); they do
  NOT pin the Mamba-2 arithmetic (parity unpinned by the reference, oracle/__init__.py).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _util import fill_weights  # noqa: E402

REF_SRC = "/root/reference/src"
sys.path.insert(0, REF_SRC)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.mamba2_ref import Mamba2Ref  # noqa: E402

shim = types.ModuleType("mamba_ssm")
shim.Mamba2 = Mamba2Ref
sys.modules["mamba_ssm"] = shim

from dcasr.models.hnet_chunk import DynamicChunker  # noqa: E402
from dcasr.models.encoder import DCASREncoder  # noqa: E402
from dcasr.models.mamba_block import MambaStack, reverse_sequences  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def hnet_case(name, B, L, D, N, lengths, seed, ema=True):
    g = torch.Generator().manual_seed(seed)
    ch = DynamicChunker(D, N=N, ema_smoothing=ema)
    fill_weights(ch, seed)
    # smooth-ish frames so that p spreads over (0,1) instead of clustering at 0.5
    base = torch.randn(B, L, D, generator=g)
    x = (base + 1.5 * torch.roll(base, 1, 1) * (torch.rand(B, L, 1, generator=g) > 0.5)).requires_grad_(True)
    mask = None
    if lengths is not None:
        mask = torch.arange(L)[None, :] < torch.tensor(lengths)[:, None]
    co = ch.chunk(x, mask)
    M = co.z.shape[1]
    z_proc = torch.randn(B, M, D, generator=g).requires_grad_(True)
    w = torch.randn(B, L, D, generator=g)
    wz = torch.randn(B, M, D, generator=g)
    y = ch.dechunk(z_proc, co)
    loss = (y * w).sum() + (co.z * wz).sum() + 0.03 * co.ratio_loss
    loss.backward()
    out = dict(x=_np(x), Wq=_np(ch.router.W_q.weight), Wk=_np(ch.router.W_k.weight),
               N=np.float64(N), ema=np.int64(ema), z_proc=_np(z_proc), w=_np(w), wz=_np(wz),
               p=_np(co.p), b=_np(co.b), membership=_np(co.membership), z=_np(co.z),
               z_mask=_np(co.z_mask), ratio_loss=_np(co.ratio_loss), kept_fraction=_np(co.kept_fraction),
               y=_np(y), loss=_np(loss), gx=_np(x.grad), gz=_np(z_proc.grad),
               gWq=_np(ch.router.W_q.weight.grad), gWk=_np(ch.router.W_k.weight.grad))
    if mask is not None:
        out["mask"] = _np(mask)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "M=", M, "kept=", float(co.kept_fraction), "margin=",
          float((co.p - 0.5).abs()[co.p > 0].min()))


def ema_case(name, B, M, D, seed, saturate):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, M, D, generator=g).requires_grad_(True)
    p = torch.rand(B, M, generator=g)
    if saturate:
        p[:, ::7] = 1.0
        p[:, 3::11] = 0.0
    p.requires_grad_(True)
    w = torch.randn(B, M, D, generator=g)
    out = DynamicChunker._ema(x, p)
    (out * w).sum().backward()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x=_np(x), p=_np(p), w=_np(w),
                        out=_np(out), gx=_np(x.grad), gp=_np(p.grad))
    print(name)


def stack_case(name, n_layers, d, B, L, lengths, seed, bidir=True):
    g = torch.Generator().manual_seed(seed)
    st = MambaStack(n_layers, d, bidirectional=bidir)
    fill_weights(st, seed)
    x = torch.randn(B, L, d, generator=g).requires_grad_(True)
    lens = torch.tensor(lengths) if lengths is not None else None
    w = torch.randn(B, L, d, generator=g)
    y = st(x, lens)
    (y * w).sum().backward()
    out = dict(x=_np(x), w=_np(w), y=_np(y), gx=_np(x.grad), seed=np.int64(seed), n_layers=np.int64(n_layers),
               d=np.int64(d), bidir=np.int64(bidir),
               rev=_np(reverse_sequences(x.detach(), lens)))
    for k, p_ in st.named_parameters():
        if k in ("layers.0.fwd.in_proj.weight", "layers.0.fwd.A_log", "layers.0.fwd.conv1d.weight",
                 "layers.0.norm.weight", "layers.0.fwd.dt_bias", "layers.0.fwd.D",
                 "layers.0.fwd.norm.weight", "layers.0.fwd.out_proj.weight", "norm.bias",
                 "layers.0.bwd.conv1d.bias" if bidir else "norm.weight"):
            out["g_" + k] = _np(p_.grad)
    if lengths is not None:
        out["lengths"] = np.asarray(lengths, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name)


def enc_case(name, arch, N, T, seed):
    g = torch.Generator().manual_seed(seed)
    enc = DCASREncoder(n_mels=80, d_outer=64, d_main=128, n_enc=1, n_main=1, n_dec=1, n_mid=1,
                       arch_type=arch, N=N)
    fill_weights(enc, seed)
    lengths = torch.tensor(list(T))
    feats = torch.randn(len(T), max(T), 80, generator=g)
    out = enc(feats, lengths)
    w = torch.randn(out.features.shape, generator=g)
    mask = (torch.arange(out.features.shape[1])[None] < out.lengths[:, None]).unsqueeze(-1)
    loss = (out.features * w * mask).sum() + 0.03 * out.ratio_loss
    loss.backward()
    d = dict(feats=_np(feats), feat_lengths=_np(lengths), w=_np(w), features=_np(out.features),
             lengths=_np(out.lengths), ratio_loss=_np(out.ratio_loss), loss=_np(loss),
             seed=np.int64(seed), N=np.int64(N), arch=np.array(arch))
    for i, (p, b) in enumerate(out.boundaries):
        d[f"p{i}"], d[f"b{i}"] = _np(p), _np(b)
        d[f"z{i}"] = _np(out.chunk_embeddings[i])
        d[f"kept{i}"] = _np(out.kept_fractions[i])
    sd = dict(enc.named_parameters())
    for k in ("subsample.proj.weight", "enc.layers.0.fwd.in_proj.weight", "enc.layers.0.bwd.A_log",
              "main.layers.0.fwd.conv1d.weight", "dec.layers.0.norm.weight", "dec.norm.bias",
              "chunk.router.W_q.weight", "chunk1.router.W_k.weight", "chunk2.router.W_q.weight",
              "proj_in.weight", "proj1_out.bias", "mid.layers.0.fwd.dt_bias"):
        if k in sd and sd[k].grad is not None:
            d["g_" + k] = _np(sd[k].grad)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, "ratio=", float(out.ratio_loss), "kept=", [float(k) for k in out.kept_fractions])


if __name__ == "__main__":
    torch.manual_seed(0)
    hnet_case("hnet_N2_ragged", B=3, L=37, D=16, N=2, lengths=[37, 20, 1], seed=11)
    hnet_case("hnet_N3_nomask", B=2, L=50, D=24, N=3, lengths=None, seed=12)
    hnet_case("hnet_sqrt2_ragged", B=4, L=64, D=32, N=2 ** 0.5, lengths=[64, 63, 33, 2], seed=13)
    hnet_case("hnet_N4_noema", B=2, L=29, D=8, N=4, lengths=[29, 17], seed=14, ema=False)
    hnet_case("hnet_N2_long", B=2, L=700, D=16, N=2, lengths=[700, 512], seed=15)
    ema_case("ema_plain", B=3, M=41, D=8, seed=21, saturate=False)
    ema_case("ema_saturated", B=2, M=60, D=5, seed=22, saturate=True)
    stack_case("stack_bidir_ragged", n_layers=2, d=64, B=3, L=45, lengths=[45, 30, 7], seed=31)
    stack_case("stack_causal", n_layers=1, d=64, B=2, L=70, lengths=None, seed=32, bidir=False)
    enc_case("enc_A_N2", "A", 2, (140, 101), seed=41)
    enc_case("enc_A_N1", "A", 1, (100, 80), seed=42)
    enc_case("enc_B_N4", "B", 4, (180, 150), seed=43)
