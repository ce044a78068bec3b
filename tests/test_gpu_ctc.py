"""CTC head on the CUDA path (dcasr_b200.CTCHead: projection GEMM + hnb_ctc_lse / hnb_ctc_alpha_beta / hnb_ctc_grad /
hnb_col_sum) against the reference's own CTCHead outputs and gradients (tests/golden/ctc_*.npz), the CPU oracle, and
PyTorch's CUDA F.ctc_loss on the reference's op sequence at the headline shape."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from _util import GOLDEN, rel_err
from test_oracle_ctc import load_head

pytestmark = pytest.mark.gpu
DEV = "cuda"
CASES = sorted(glob.glob(os.path.join(GOLDEN, "ctc_*.npz")))


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[4:-4] for p in CASES])
def test_ctc_head_matches_reference_golden_fp32(path):
    import dcasr_b200 as dd
    g = np.load(path)
    head = load_head(g, dd.CTCHead).to(DEV)
    x = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
    fl, tl, tg = (torch.from_numpy(g[k]).to(DEV) for k in ("feat_lens", "tgt_lens", "targets"))
    dd.reset_launch_count()
    assert rel_err(head.log_probs(x).cpu(), torch.from_numpy(g["log_probs"])) < 1e-5
    for red in ("mean", "sum", "none"):
        got = head.loss(x, fl, tg, tl, reduction=red)
        np.testing.assert_allclose(got.detach().cpu().numpy(), g["loss_" + red], rtol=1e-4, atol=1e-4)
    head.loss(x, fl, tg, tl).backward()
    assert dd.launch_count() > 0
    assert rel_err(x.grad.cpu(), torch.from_numpy(g["dx"])) < 1e-3
    assert rel_err(head.proj.weight.grad.cpu(), torch.from_numpy(g["dweight"])) < 1e-3
    assert rel_err(head.proj.bias.grad.cpu(), torch.from_numpy(g["dbias"])) < 1e-3
    am = head.frame_argmax(x).cpu()
    ref_am = torch.from_numpy(g["frame_argmax"])
    lp = torch.from_numpy(g["log_probs"])
    top2 = lp.topk(2, -1).values
    clear = (top2[..., 0] - top2[..., 1]) > 1e-4             # the frame's winner is not a numerical tie
    assert torch.equal(am[clear], ref_am[clear]) and am.dtype == torch.int64
    dec = head.greedy_decode(x, fl)
    if bool(clear.all()):
        for i, n in enumerate(g["greedy_len"]):
            assert dec[i] == g["greedy"][i, :n].tolist()


def test_ctc_head_bf16_autocast_vs_reference_ops_on_gpu():
    """bf16 autocast at the headline shape (40 x 398 frames, d 384, 500 pieces + blank): our fused path against the
    reference's op sequence (nn.Linear -> .float() -> log_softmax -> F.ctc_loss) with the same weights under the same
    autocast; north_star's 2e-2."""
    import dcasr_b200 as dd
    torch.manual_seed(0)
    B, T, d, V, U = 40, 398, 384, 500, 60
    head = dd.CTCHead(d, V).to(DEV)
    with torch.no_grad():
        head.proj.weight.mul_(3.0)
    x = torch.randn(B, T, d, device=DEV)
    fl = torch.randint(300, T + 1, (B,), device=DEV); fl[0] = T
    tl = torch.randint(20, U + 1, (B,), device=DEV)
    tg = torch.randint(0, V, (B, U), device=DEV)

    def run(mine):
        xx = x.clone().requires_grad_(True)
        for p in head.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if mine:
                loss = head.loss(xx, fl, tg, tl)
            else:
                lp = F.log_softmax(F.linear(xx, head.proj.weight, head.proj.bias).float(), dim=-1).transpose(0, 1)
                loss = F.ctc_loss(lp, tg, fl, tl, blank=V, reduction="mean", zero_infinity=True)
        loss.backward()
        return loss.detach(), xx.grad.clone(), head.proj.weight.grad.clone(), head.proj.bias.grad.clone()

    l1, dx1, dw1, db1 = run(True)
    l0, dx0, dw0, db0 = run(False)
    print(f"ctc bf16: loss {float(l1):.5f} vs {float(l0):.5f}; rel err dx {rel_err(dx1, dx0):.2e}, dW {rel_err(dw1, dw0):.2e}, "
          f"db {rel_err(db1, db0):.2e}")
    assert abs(float(l1) - float(l0)) < 2e-2 * abs(float(l0))
    assert rel_err(dx1, dx0) < 2e-2 and rel_err(dw1, dw0) < 2e-2 and rel_err(db1, db0) < 2e-2


def test_ctc_loss_fp32_vs_torch_cuda_ctc_random_shapes():
    """The loss kernels alone on fp32 logits against PyTorch's CUDA F.ctc_loss: ragged lengths, repeated labels, long targets."""
    from dcasr_b200.ctc import _CTCLossFn
    g = torch.Generator().manual_seed(7)
    for (B, T, V1, U) in ((5, 50, 12, 9), (3, 200, 31, 64), (2, 7, 4, 3), (4, 128, 501, 40)):
        logits = (torch.randn(B, T, V1, generator=g) * 2.0).to(DEV).requires_grad_(True)
        fl = torch.randint(max(1, T // 2), T + 1, (B,), generator=g).to(DEV)
        tl = torch.randint(0, U + 1, (B,), generator=g).to(DEV)
        tg = torch.randint(0, min(V1 - 1, 3), (B, U), generator=g).to(DEV)        # tiny alphabet: many repeated labels
        for red in ("mean", "sum"):
            got = _CTCLossFn.apply(logits, fl, tg, tl, V1 - 1, red)
            g1, = torch.autograd.grad(got, logits)
            ref = F.ctc_loss(F.log_softmax(logits, -1).transpose(0, 1), tg, fl, tl, blank=V1 - 1, reduction=red, zero_infinity=True)
            g0, = torch.autograd.grad(ref, logits)
            assert abs(float(got) - float(ref)) <= 1e-4 * max(1.0, abs(float(ref))), (B, T, V1, U, red, float(got), float(ref))
            assert rel_err(g1, g0) < 1e-3, (B, T, V1, U, red, rel_err(g1, g0))


def test_ctc_head_rejects_cpu_tensors_and_installs_under_the_reference_name():
    import dcasr_b200 as dd
    head = dd.CTCHead(16, 5)
    assert sorted(head.state_dict()) == ["proj.bias", "proj.weight"] and head.blank_id == 5 and head.num_classes == 6
    with pytest.raises(dd.HnbError):
        head.loss(torch.randn(1, 4, 16), torch.tensor([4]), torch.tensor([[1]]), torch.tensor([1]))
