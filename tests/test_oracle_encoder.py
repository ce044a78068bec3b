"""Pin oracle/encoder_ref.py (assembly) against golden vectors produced by the reference's own
encoder.py / mamba_block.py (with the oracle Mamba2 supplied for the missing mamba_ssm)."""
import glob
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN, fill_weights, max_err, rel_err
from oracle.encoder_ref import EncoderRef, MambaStackRef, reverse_ref

STACK = sorted(glob.glob(os.path.join(GOLDEN, "stack_*.npz")))
ENC = sorted(glob.glob(os.path.join(GOLDEN, "enc_*.npz")))


@pytest.mark.parametrize("path", STACK, ids=[os.path.basename(p)[:-4] for p in STACK])
def test_stack_oracle_matches_reference(path):
    g = np.load(path)
    st = MambaStackRef(int(g["n_layers"]), int(g["d"]), bool(g["bidir"]))
    fill_weights(st, int(g["seed"]))
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    lens = torch.from_numpy(g["lengths"]) if "lengths" in g else None
    assert torch.equal(reverse_ref(x.detach(), lens), torch.from_numpy(g["rev"]))
    y = st(x, lens)
    assert rel_err(y, torch.from_numpy(g["y"])) < 1e-5
    (y * torch.from_numpy(g["w"])).sum().backward()
    assert rel_err(x.grad, torch.from_numpy(g["gx"])) < 1e-4
    sd = dict(st.named_parameters())
    for k in g.files:
        if k.startswith("g_"):
            assert rel_err(sd[k[2:]].grad, torch.from_numpy(g[k])) < 1e-4, k


@pytest.mark.parametrize("path", ENC, ids=[os.path.basename(p)[:-4] for p in ENC])
def test_encoder_oracle_matches_reference(path):
    g = np.load(path)
    enc = EncoderRef(n_mels=80, d_outer=64, d_main=128, n_enc=1, n_main=1, n_dec=1, n_mid=1,
                     arch_type=str(g["arch"]), N=int(g["N"]))
    fill_weights(enc, int(g["seed"]))
    out = enc(torch.from_numpy(g["feats"]), torch.from_numpy(g["feat_lengths"]))
    assert torch.equal(out.lengths, torch.from_numpy(g["lengths"]))
    i = 0
    while f"p{i}" in g:
        p, b = out.boundaries[i]
        assert max_err(p, torch.from_numpy(g[f"p{i}"])) < 1e-5
        assert torch.equal(b, torch.from_numpy(g[f"b{i}"]))
        assert rel_err(out.chunk_embeddings[i], torch.from_numpy(g[f"z{i}"])) < 1e-5
        assert max_err(out.kept_fractions[i], torch.from_numpy(g[f"kept{i}"])) < 1e-7
        i += 1
    mask = (torch.arange(out.features.shape[1])[None] < out.lengths[:, None]).unsqueeze(-1)
    ref = torch.from_numpy(g["features"])
    assert rel_err(out.features * mask, ref * mask) < 1e-4
    assert max_err(out.ratio_loss, torch.from_numpy(g["ratio_loss"])) < 1e-5
    loss = (out.features * torch.from_numpy(g["w"]) * mask).sum() + 0.03 * out.ratio_loss
    loss.backward()
    sd = dict(enc.named_parameters())
    for k in g.files:
        if k.startswith("g_"):
            assert rel_err(sd[k[2:]].grad, torch.from_numpy(g[k])) < 2e-3, k


ENC_FIXED = sorted(glob.glob(os.path.join(GOLDEN, "encfixed_*.npz")))


@pytest.mark.parametrize("path", ENC_FIXED, ids=[os.path.basename(p)[:-4] for p in ENC_FIXED])
def test_encoder_oracle_with_fixed_chunker_matches_reference(path):
    """`chunker: fixed` (src/dcasr/models/encoder.py:30-37, fixed_pool.py): the oracle's assembly against the reference's
    own DCASREncoder run with its FixedPoolChunker (tests/golden/make_golden_fixed.py)."""
    g = np.load(path)
    enc = EncoderRef(n_mels=80, d_outer=64, d_main=128, n_enc=1, n_main=1, n_dec=1, n_mid=1,
                     arch_type=str(g["arch"]), N=int(g["N"]), chunker="fixed")
    fill_weights(enc, int(g["seed"]))
    out = enc(torch.from_numpy(g["feats"]), torch.from_numpy(g["feat_lengths"]))
    assert torch.equal(out.lengths, torch.from_numpy(g["lengths"])) and float(out.ratio_loss) == 0.0
    i = 0
    while f"p{i}" in g:
        p, b = out.boundaries[i]
        assert torch.equal(p, torch.from_numpy(g[f"p{i}"])) and torch.equal(b, torch.from_numpy(g[f"b{i}"]))
        assert rel_err(out.chunk_embeddings[i], torch.from_numpy(g[f"z{i}"])) < 1e-5
        assert max_err(out.kept_fractions[i], torch.from_numpy(g[f"kept{i}"])) < 1e-7
        i += 1
    mask = (torch.arange(out.features.shape[1])[None] < out.lengths[:, None]).unsqueeze(-1)
    assert rel_err(out.features * mask, torch.from_numpy(g["features"]) * mask) < 1e-4
    ((out.features * torch.from_numpy(g["w"]) * mask).sum() + 0.03 * out.ratio_loss).backward()
    sd = dict(enc.named_parameters())
    for k in g.files:
        if k.startswith("g_"):
            assert rel_err(sd[k[2:]].grad, torch.from_numpy(g[k])) < 2e-3, k
