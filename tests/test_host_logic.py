"""CPU: host-side mirror of the reference interface (construction, state_dict layout, registries, errors)."""
import sys

import pytest
import torch

import dcasr_b200 as d
from oracle.encoder_ref import EncoderRef
from oracle.mamba2_ref import Mamba2Ref


def test_state_dict_layout_equals_oracle_and_reference_closed_form():
    kw = dict(d_outer=64, d_main=128, n_enc=1, n_main=2, n_dec=1, n_mid=1)
    for arch, N in (("A", 1), ("A", 2), ("B", 4)):
        a, b = d.DCASREncoder(arch_type=arch, N=N, **kw), EncoderRef(arch_type=arch, N=N, **kw)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb)
        assert all(sa[k].shape == sb[k].shape for k in sa)
        b.load_state_dict(sa)
    m = d.Mamba2(384)
    assert sum(p.numel() for p in m.parameters()) == 993_572
    assert {k: v.shape for k, v in m.state_dict().items()} == {k: v.shape for k, v in Mamba2Ref(384).state_dict().items()}
    assert all(getattr(getattr(m, k), "_no_weight_decay", False) for k in ("A_log", "D", "dt_bias"))
    # Type A Small N=2 encoder parameters (62.01 M)
    n = sum(p.numel() for p in d.DCASREncoder(N=2).parameters())
    assert abs(n - 62.01e6) < 0.01e6, n


def test_registries_and_errors():
    with pytest.raises(ValueError):
        d.DCASREncoder(arch_type="C")
    with pytest.raises(ValueError):
        d.build_chunker("nope", 64, 2)
    with pytest.raises(AssertionError):
        d.MambaBlock(80, headdim=64)
    assert isinstance(d.build_chunker("dynamic", 64, 2), d.DynamicChunker)
    ch = d.DynamicChunker(32, N=1)
    assert ch.router is None and ch.identity
    ch2 = d.DynamicChunker(32, N=2 ** 0.5)
    assert ch2.router is not None and torch.equal(ch2.router.W_q.weight, torch.eye(32))
    names = [n for n, _ in d.DCASREncoder(d_outer=64, d_main=128, n_enc=1, n_main=1, n_dec=1, N=2).named_parameters()]
    assert "chunk.router.W_q.weight" in names and "enc.layers.0.bwd.A_log" in names and "proj_in.bias" in names


def test_identity_chunker_runs_on_cpu_like_the_reference():
    ch = d.DynamicChunker(8, N=1)
    x = torch.randn(2, 5, 8)
    co = ch.chunk(x)
    assert torch.equal(co.z, x) and float(co.ratio_loss) == 0.0 and torch.equal(ch.dechunk(co.z, co), x)
    assert float(d.ratio_loss(torch.rand(2, 5), torch.ones(2, 5), 1)) == 0.0


def test_install_shim_provides_mamba_ssm():
    saved = sys.modules.get("mamba_ssm")
    try:
        d.install(patch_dcasr=False)
        from mamba_ssm import Mamba2
        assert Mamba2 is d.Mamba2
    finally:
        if saved is not None:
            sys.modules["mamba_ssm"] = saved
        else:
            sys.modules.pop("mamba_ssm", None)


def test_ctc_host_helpers():
    """Host-side pieces of the CTC head that need no GPU: the greedy collapse rule (reference decoders/ctc.py:70-83), the
    unpacking of 1-D concatenated targets, the padded row stride of the logits buffer, and the CPU refusal."""
    import pytest
    import torch
    import dcasr_b200 as dd
    from dcasr_b200.ctc import _ld, _pad_targets
    assert dd.ctc_greedy_collapse([5, 5, 9, 9, 5, 9, 3, 3, 9], 9) == [5, 5, 3]          # a blank separates equal labels
    assert dd.ctc_greedy_collapse([], 9) == [] and dd.ctc_greedy_collapse([9, 9], 9) == []
    flat = torch.tensor([1, 2, 3, 7, 8])
    pad = _pad_targets(flat, torch.tensor([3, 0, 2]), 3)
    assert pad.shape == (3, 3) and pad.dtype == torch.int64
    assert pad[0].tolist() == [1, 2, 3] and pad[2, :2].tolist() == [7, 8]
    two_d = torch.tensor([[1, 2], [3, 4]], dtype=torch.int32)
    assert _pad_targets(two_d, torch.tensor([2, 1]), 2).dtype == torch.int64
    assert _ld(501) == 512 and _ld(512) == 512 and _ld(6) == 16                       # 32-byte aligned bf16 rows
    head = dd.CTCHead(8, 5)
    assert head.blank_id == 5 and head.num_classes == 6 and sorted(head.state_dict()) == ["proj.bias", "proj.weight"]
    assert dd.CTCHead(8, 5, blank_id=0).blank_id == 0
    with pytest.raises(dd.HnbError):
        head.loss(torch.randn(1, 3, 8), torch.tensor([3]), torch.tensor([[1]]), torch.tensor([1]))
