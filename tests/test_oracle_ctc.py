"""oracle/ctc_ref.py pinned against the reference's own CTCHead (tests/golden/ctc_*.npz, made by make_golden_ctc.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN, rel_err
from oracle.ctc_ref import CTCHeadRef

CASES = sorted(glob.glob(os.path.join(GOLDEN, "ctc_*.npz")))


def load_head(g, cls, **kw):
    V = int(g["V"])
    head = cls(g["x"].shape[2], V, **kw)
    head.load_state_dict({"proj.weight": torch.from_numpy(g["weight"]), "proj.bias": torch.from_numpy(g["bias"])})
    return head


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[4:-4] for p in CASES])
def test_oracle_ctc_matches_reference(path):
    g = np.load(path)
    head = load_head(g, CTCHeadRef)
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    fl, tl, tg = torch.from_numpy(g["feat_lens"]), torch.from_numpy(g["tgt_lens"]), torch.from_numpy(g["targets"])
    assert rel_err(head.log_probs(x), torch.from_numpy(g["log_probs"])) < 1e-6
    for red in ("mean", "sum", "none"):
        got = head.loss(x, fl, tg, tl, reduction=red)
        np.testing.assert_allclose(got.detach().numpy(), g["loss_" + red], rtol=2e-5, atol=1e-5)
    head.loss(x, fl, tg, tl).backward()
    assert rel_err(x.grad, torch.from_numpy(g["dx"])) < 2e-4        # fp32 recursions of different summation order
    assert rel_err(head.proj.weight.grad, torch.from_numpy(g["dweight"])) < 2e-4
    assert rel_err(head.proj.bias.grad, torch.from_numpy(g["dbias"])) < 2e-4
    assert torch.equal(head.frame_argmax(x), torch.from_numpy(g["frame_argmax"]))


def test_ctc_cases_cover_the_edge_cases():
    g = np.load(os.path.join(GOLDEN, "ctc_edge.npz"))
    assert g["loss_none"][2] == 0.0            # target longer than its input: infeasible, zeroed by zero_infinity
    assert g["tgt_lens"][1] == 0               # empty target
    assert len(CASES) >= 5
