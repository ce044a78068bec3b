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


if __name__ == "__main__":
    case("N2_nomask", 2, 37, 24, 2, None, 1)
    case("N2_ragged", 3, 50, 40, 2, [50, 31, 1], 2)             # a 1-frame row
    case("N3_padded_tail", 3, 41, 16, 3, [20, 7, 13], 3)        # L far beyond the longest row: tail clamped into the last window
    case("N4_ragged", 4, 64, 32, 4, [64, 33, 5, 60], 4)
    case("N1_identity", 2, 19, 8, 1, [19, 11], 5)
