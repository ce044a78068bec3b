"""Golden vectors for the CTC head, produced by RUNNING THE REFERENCE ITSELF
(/root/reference/src/dcasr/decoders/ctc.py CTCHead, unmodified, CPU fp32).  Build container only:
    python tests/golden/make_golden_ctc.py
Each ctc_*.npz holds the head's parameters, inputs, log_probs, the loss for every reduction and the autograd gradients of
the 'mean' loss w.r.t. features, proj.weight and proj.bias, plus frame_argmax and greedy_decode."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/src")
from dcasr.decoders.ctc import CTCHead  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def case(name, B, T, d, V, feat_lens, targets, tgt_lens, seed, scale=1.0, concat=False):
    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed)
    head = CTCHead(d, V)
    with torch.no_grad():
        head.proj.weight.mul_(scale * 4.0)                     # sharper posteriors than the default init gives
    x = torch.randn(B, T, d, generator=g).requires_grad_(True)
    fl, tl = torch.tensor(feat_lens), torch.tensor(tgt_lens)
    tg = torch.tensor(targets, dtype=torch.long)
    if concat:
        tg = torch.cat([tg[i, :n] for i, n in enumerate(tgt_lens)])
    out = {"weight": _np(head.proj.weight), "bias": _np(head.proj.bias), "x": _np(x), "feat_lens": _np(fl), "tgt_lens": _np(tl),
           "targets": _np(tg), "V": np.int64(V), "log_probs": _np(head.log_probs(x)),
           "frame_argmax": _np(head.frame_argmax(x))}
    for red in ("mean", "sum", "none"):
        out["loss_" + red] = _np(head.loss(x, fl, tg, tl, reduction=red))
    loss = head.loss(x, fl, tg, tl)
    loss.backward()
    out.update(dx=_np(x.grad), dweight=_np(head.proj.weight.grad), dbias=_np(head.proj.bias.grad))
    dec = head.greedy_decode(x, fl)
    out["greedy_len"] = np.array([len(v) for v in dec])
    out["greedy"] = np.array([v + [-1] * (T - len(v)) for v in dec])
    np.savez_compressed(os.path.join(HERE, f"ctc_{name}.npz"), **out)
    print(name, "loss", float(loss), "none", out["loss_none"])


if __name__ == "__main__":
    g = torch.Generator().manual_seed(0)
    r = lambda B, U, V: torch.randint(0, V, (B, U), generator=g).tolist()    # noqa: E731
    case("basic", 3, 40, 32, 20, [40, 33, 25], r(3, 7, 20), [7, 5, 3], 1)
    # repeated labels (need a blank between them), an EMPTY target, and a target longer than its input (infeasible: zero_infinity)
    case("edge", 4, 12, 16, 6, [12, 12, 4, 9], [[1, 1, 1, 2, 2, 0, 0, 0], [3] * 8, [0, 1, 2, 3, 4, 5, 0, 1], [2, 2, 2, 2, 2, 0, 0, 0]],
         [5, 0, 8, 5], 2)
    case("vocab500", 2, 99, 64, 500, [99, 71], r(2, 24, 500), [24, 17], 3)                 # the reference's vocabulary size
    case("concat_targets", 3, 30, 24, 11, [30, 30, 21], r(3, 6, 11), [6, 2, 4], 4, concat=True)
    case("one_frame", 2, 1, 8, 5, [1, 1], [[3], [2]], [1, 0], 5)
