"""Pin oracle/fixed_pool_ref.py against golden vectors produced by the reference's own fixed_pool.py, and check the
host-side (GPU-free) logic of the CUDA mirror: constructor errors, the N = 1 identity, the registry."""
import glob
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN, max_err
from oracle import fixed_pool_ref

FIXED = sorted(p for p in glob.glob(os.path.join(GOLDEN, "fixed_*.npz")) if "identity" not in p)


def _t(a, grad=False):
    t = torch.from_numpy(np.asarray(a))
    return t.requires_grad_(True) if grad else t


@pytest.mark.parametrize("path", FIXED, ids=[os.path.basename(p)[:-4] for p in FIXED])
def test_fixed_pool_oracle_matches_reference(path):
    g = np.load(path)
    x, z_proc = _t(g["x"], True), _t(g["z_proc"], True)
    mask = _t(g["mask"]) if "mask" in g else None
    co = fixed_pool_ref.chunk_ref(x, int(g["N"]), mask)
    assert torch.equal(co.membership, _t(g["membership"])) and torch.equal(co.z_mask, _t(g["z_mask"]))
    assert torch.equal(co.b, _t(g["b"])) and torch.equal(co.p, _t(g["p"]))
    assert max_err(co.z, _t(g["z"])) < 1e-6
    assert max_err(co.kept_fraction, _t(g["kept_fraction"])) < 1e-7 and float(co.ratio_loss) == 0.0
    out = fixed_pool_ref.dechunk_ref(z_proc, co.membership)
    assert torch.equal(out, _t(g["out"]))
    ((out * _t(g["w"])).sum() + (co.z * _t(g["v"])).sum()).backward()
    assert max_err(x.grad, _t(g["dx"])) < 1e-6 and max_err(z_proc.grad, _t(g["dz_proc"])) < 1e-5


def test_fixed_pool_host_logic():
    import dcasr_b200 as dd
    from dcasr_b200.encoder import build_chunker
    with pytest.raises(ValueError):
        dd.FixedPoolChunker(16, N=2 ** 0.5)                 # Type B at a non-square N: no fractional window
    with pytest.raises(ValueError):
        dd.FixedPoolChunker(16, N=0)
    ch = build_chunker("fixed", 16, 4)
    assert isinstance(ch, dd.FixedPoolChunker) and ch.stride == 4 and ch.N == 4 and not ch.identity
    assert len(list(ch.parameters())) == 0
    g = np.load(os.path.join(GOLDEN, "fixed_N1_identity.npz"))             # stride 1 is pure tensor plumbing: CPU is fine
    ident = dd.FixedPoolChunker(8, N=1)
    x, mask = _t(g["x"]), _t(g["mask"])
    co = ident.chunk(x, mask)
    assert co.z is x and torch.equal(co.membership, _t(g["membership"])) and torch.equal(co.b, _t(g["b"]))
    assert torch.equal(co.z_mask, _t(g["z_mask"])) and float(co.kept_fraction) == 1.0 and float(co.ratio_loss) == 0.0
    z = _t(g["z_proc"])
    assert ident.dechunk(z, co) is z


ALL_FIXED = sorted(glob.glob(os.path.join(GOLDEN, "fixed_*.npz")))


@pytest.mark.parametrize("path", ALL_FIXED, ids=[os.path.basename(p)[:-4] for p in ALL_FIXED])
def test_fixed_pool_host_wiring_with_stand_in_kernels(path, monkeypatch):
    """The Python side of the CUDA FixedPoolChunker (index tensors, ChunkOutput fields, autograd wiring of the two
    kernels) against the reference's golden vectors, with the two kernel entry points replaced by torch stand-ins that
    follow include/hnet_b200.h.  (The kernels themselves are checked on the GPU, tests/test_gpu_hnet.py.)"""
    import dcasr_b200 as dd
    import dcasr_b200.fixed_pool as fp
    from dcasr_b200 import ops

    def windows(L, M, stride):
        return torch.div(torch.arange(L), stride, rounding_mode="floor").clamp(max=M - 1)

    def window_reduce(x, mask_u8, M, stride, normalize, z_dtype, want_cnt=True):
        B, L, D = x.shape
        m = torch.ones(B, L) if mask_u8 is None else mask_u8.float()
        z = torch.zeros(B, M, D).index_add_(1, windows(L, M, stride), x.float() * m[..., None])
        cnt = torch.zeros(B, M).index_add_(1, windows(L, M, stride), m)
        return (z / cnt.clamp_min(1)[..., None] if normalize else z).to(z_dtype), cnt

    def window_broadcast(z, mask_u8, cnt, resid, L, stride, out_dtype):
        w = windows(L, z.shape[1], stride)
        out = z.float()[:, w]
        if cnt is not None:
            out = out / cnt.clamp_min(1)[:, w][..., None]
        if mask_u8 is not None:
            out = out * mask_u8.float()[..., None]
        return (out if resid is None else out + resid.float()).to(out_dtype)

    monkeypatch.setattr(ops, "window_reduce", window_reduce)
    monkeypatch.setattr(ops, "window_broadcast", window_broadcast)
    monkeypatch.setattr(fp, "_mask_u8", lambda m: None if m is None else m.to(torch.uint8))
    g = np.load(path)
    x, z_proc = _t(g["x"], True), _t(g["z_proc"], True)
    mask = _t(g["mask"]) if "mask" in g else None
    ch = dd.FixedPoolChunker(x.shape[-1], N=int(g["N"]))
    co = ch.chunk(x, mask)
    assert torch.equal(co.membership, _t(g["membership"])) and torch.equal(co.z_mask, _t(g["z_mask"]))
    assert torch.equal(co.b, _t(g["b"])) and torch.equal(co.p, _t(g["p"])) and max_err(co.z, _t(g["z"])) < 1e-6
    assert abs(float(co.kept_fraction) - float(g["kept_fraction"])) < 1e-6 and float(co.ratio_loss) == 0.0
    out = ch.dechunk(z_proc, co)
    assert torch.equal(out, _t(g["out"]))
    if int(g["N"]) > 1:
        ((out * _t(g["w"])).sum() + (co.z * _t(g["v"])).sum()).backward()
        assert max_err(x.grad, _t(g["dx"])) < 1e-6 and max_err(z_proc.grad, _t(g["dz_proc"])) < 1e-5
        r = torch.randn_like(out)
        assert max_err(ch.dechunk(z_proc, co, residual=r), r + out) < 1e-6
