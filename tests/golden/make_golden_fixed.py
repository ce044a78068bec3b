"""Golden vectors for the fixed-stride pooling chunker, produced by RUNNING THE REFERENCE ITSELF
(/root/reference/src/dcasr/models/fixed_pool.py, unmodified, CPU fp32).  Build container only:
    python tests/golden/make_golden_fixed.py
Each fixed_*.npz holds inputs, every ChunkOutput field, dechunk's output and autograd gradients of
sum(w * dechunk(z_proc)) + sum(v * z) w.r.t. x and z_proc."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/src")
from dcasr.models.fixed_pool import FixedPoolChunker  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def case(name, B, L, D, N, lengths, seed):
    g = torch.Generator().manual_seed(seed)
    ch = FixedPoolChunker(D, N=N)
    x = torch.randn(B, L, D, generator=g).requires_grad_(True)
    mask = None if lengths is None else torch.arange(L)[None, :] < torch.tensor(lengths)[:, None]
    co = ch.chunk(x, mask)
    M = co.z.shape[1]
    z_proc = torch.randn(B, M, D, generator=g).requires_grad_(True)
    w, v = torch.randn(B, L, D, generator=g), torch.randn(B, M, D, generator=g)
    out = ch.dechunk(z_proc, co)
    ((out * w).sum() + (co.z * v).sum()).backward()
    d = dict(x=_np(x), N=np.float64(N), z=_np(co.z), z_mask=_np(co.z_mask), p=_np(co.p), b=_np(co.b),
             membership=_np(co.membership), ratio_loss=_np(co.ratio_loss), kept_fraction=_np(co.kept_fraction),
             z_proc=_np(z_proc), w=_np(w), v=_np(v), out=_np(out), dx=_np(x.grad), dz_proc=_np(z_proc.grad))
    if mask is not None:
        d["mask"] = _np(mask)
    np.savez_compressed(os.path.join(HERE, f"fixed_{name}.npz"), **d)
    print(name, "M =", M, "kept =", float(co.kept_fraction))


def encoder_case(name, arch, N, T, seed):
    """The reference's own DCASREncoder with `chunker: fixed` (encoder.py:30-37), the missing third-party mamba_ssm
    supplied by oracle/mamba2_ref.py as in make_golden.py.  Named encfixed_* so that the enc_* globs of the dynamic
    chunker's tests do not pick it up."""
    import types
    sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from _util import fill_weights
    from oracle.mamba2_ref import Mamba2Ref
    shim = types.ModuleType("mamba_ssm")
    shim.Mamba2 = Mamba2Ref
    sys.modules["mamba_ssm"] = shim
    from dcasr.models.encoder import DCASREncoder
    g = torch.Generator().manual_seed(seed)
    enc = DCASREncoder(n_mels=80, d_outer=64, d_main=128, n_enc=1, n_main=1, n_dec=1, n_mid=1, arch_type=arch, N=N,
                       chunker="fixed")
    fill_weights(enc, seed)
    lengths = torch.tensor(list(T))
    feats = torch.randn(len(T), max(T), 80, generator=g)
    out = enc(feats, lengths)
    w = torch.randn(out.features.shape, generator=g)
    mask = (torch.arange(out.features.shape[1])[None] < out.lengths[:, None]).unsqueeze(-1)
    loss = (out.features * w * mask).sum() + 0.03 * out.ratio_loss
    loss.backward()
    d = dict(feats=_np(feats), feat_lengths=_np(lengths), w=_np(w), features=_np(out.features), lengths=_np(out.lengths),
             ratio_loss=_np(out.ratio_loss), seed=np.int64(seed), N=np.int64(N), arch=np.array(arch))
    for i, (p, b) in enumerate(out.boundaries):
        d[f"p{i}"], d[f"b{i}"], d[f"z{i}"] = _np(p), _np(b), _np(out.chunk_embeddings[i])
        d[f"kept{i}"] = _np(out.kept_fractions[i])
    sd = dict(enc.named_parameters())
    for k in ("subsample.proj.weight", "enc.layers.0.fwd.in_proj.weight", "main.layers.0.fwd.conv1d.weight",
              "dec.layers.0.norm.weight", "proj_in.weight", "proj1_out.bias", "mid.layers.0.fwd.dt_bias"):
        if k in sd and sd[k].grad is not None:
            d["g_" + k] = _np(sd[k].grad)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, "kept =", [float(k) for k in out.kept_fractions])


if __name__ == "__main__":
    encoder_case("encfixed_A_N2", "A", 2, (330, 211), 21)
    encoder_case("encfixed_B_N4", "B", 4, (300, 97, 250), 22)
    case("N2_nomask", 2, 37, 24, 2, None, 1)
    case("N2_ragged", 3, 50, 40, 2, [50, 31, 1], 2)             # a 1-frame row
    case("N3_padded_tail", 3, 41, 16, 3, [20, 7, 13], 3)        # L far beyond the longest row: tail clamped into the last window
    case("N4_ragged", 4, 64, 32, 4, [64, 33, 5, 60], 4)
    case("N1_identity", 2, 19, 8, 1, [19, 11], 5)
