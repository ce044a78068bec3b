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
