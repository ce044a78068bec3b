"""Cross-checks for oracle/mamba2_ref.py (parity unpinned by the reference; see oracle/__init__.py):
(i) chunked SSD == sequential recurrence (fp64), (ii) several chunk sizes agree,
(iii) the whole mixer == transformers' independent Mamba2Mixer.torch_forward with copied weights."""
import pytest
import torch

from _util import fill_weights, max_err, rel_err
from oracle.mamba2_ref import Mamba2Ref, ssd_chunked, ssd_sequential


def _rand_ssd(B=2, L=77, H=3, P=8, N=16, dtype=torch.float64, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, L, H, P, generator=g, dtype=dtype)
    dt = torch.nn.functional.softplus(torch.randn(B, L, H, generator=g, dtype=dtype) - 1.0)
    A = -torch.rand(H, generator=g, dtype=dtype) * 8 - 0.5
    Bm = torch.randn(B, L, N, generator=g, dtype=dtype)
    Cm = torch.randn(B, L, N, generator=g, dtype=dtype)
    D = torch.randn(H, generator=g, dtype=dtype)
    return x, dt, A, Bm, Cm, D


@pytest.mark.parametrize("chunk", [1, 16, 64, 128])
def test_chunked_equals_sequential_fp64(chunk):
    args = _rand_ssd()
    assert max_err(ssd_chunked(*args, chunk=chunk), ssd_sequential(*args)) < 1e-10


def test_chunked_gradients_equal_sequential_fp64():
    a1 = [t.clone().requires_grad_(True) for t in _rand_ssd(L=40)]
    a2 = [t.clone().requires_grad_(True) for t in _rand_ssd(L=40)]
    w = torch.randn_like(a1[0])
    (ssd_chunked(*a1, chunk=16) * w).sum().backward()
    (ssd_sequential(*a2) * w).sum().backward()
    for u, v in zip(a1, a2):
        assert max_err(u.grad, v.grad) < 1e-9


@pytest.mark.parametrize("d_model", [64, 128])
def test_mixer_equals_hf_torch_forward(d_model):
    tr = pytest.importorskip("transformers")
    from transformers.models.mamba2.configuration_mamba2 import Mamba2Config
    from transformers.models.mamba2.modeling_mamba2 import Mamba2Mixer
    cfg = Mamba2Config(hidden_size=d_model, state_size=128, conv_kernel=4, expand=2, head_dim=64,
                       num_heads=2 * d_model // 64, n_groups=1, chunk_size=32, use_bias=False,
                       use_conv_bias=True, rms_norm=True, layer_norm_epsilon=1e-5, hidden_act="silu")
    hf = Mamba2Mixer(cfg, layer_idx=0).double().eval()
    ours = Mamba2Ref(d_model).double()
    fill_weights(ours, 5)
    sd = ours.state_dict()
    hf.load_state_dict({k: v for k, v in sd.items()}, strict=True)      # identical key set & shapes
    u = torch.randn(2, 50, d_model, dtype=torch.float64)
    with torch.no_grad():
        y_hf = hf.torch_forward(u)
        y = ours(u)
        ours.mode = "sequential"
        y_seq = ours(u)
    assert rel_err(y, y_hf) < 1e-6      # HF computes A and dt in fp32 internally
    assert rel_err(y_seq, y_hf) < 1e-6
    assert rel_err(y, y_seq) < 1e-10


def test_param_layout_matches_reference_closed_form():
    # src/dcasr/eval/efficiency.py:49-57 closed form; SURVEY.md App. B: 993 572 @384, 1 719 600 @512
    for d, n in ((384, 993_572), (512, 1_719_600)):
        m = Mamba2Ref(d)
        assert sum(p.numel() for p in m.parameters()) == n
        assert all(getattr(getattr(m, k), "_no_weight_decay", False) for k in ("A_log", "D", "dt_bias"))
        assert m.in_proj.weight.shape == (2 * 2 * d + 2 * 128 + 2 * d // 64, d)
        assert m.conv1d.weight.shape == (2 * d + 256, 1, 4)
