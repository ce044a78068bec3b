"""CPU restatement of the reference's CTC head (TEST INFRASTRUCTURE: imported by tests/ and bench.py's checker only).

Follows /root/reference/src/dcasr/decoders/ctc.py:85-127 (`CTCHead`: Linear d_model -> V+1, blank at id V; `log_probs` =
fp32 log-softmax; `loss` = `F.ctc_loss(lp^T, targets, feat_lengths, target_lengths, blank, reduction, zero_infinity=True)`).
`F.ctc_loss` itself is PyTorch library code; its algorithm (Graves et al. 2006: forward variable over the blank-extended
label sequence) is restated here as plain differentiable torch ops, so that autograd -- not a hand-derived formula --
provides the gradients the CUDA kernels are checked against.  Pinned by tests/golden/ctc_*.npz, which
tests/golden/make_golden_ctc.py produced by running the reference's own CTCHead (values and gradients).
"""
from __future__ import annotations

import torch
import torch.nn as nn

NEG = float("-inf")


def ctc_nll(lp: torch.Tensor, target: torch.Tensor, blank: int) -> torch.Tensor:
    """-log p(target | lp) for ONE utterance: lp [T, V1] log-probabilities, target [U] int.  inf when no alignment exists."""
    T = lp.shape[0]
    U = int(target.shape[0])
    if T == 0:
        return lp.new_zeros(()) if U == 0 else lp.new_full((), float("inf"))
    ext = torch.full((2 * U + 1,), blank, dtype=torch.long)
    ext[1::2] = target.long()
    S = ext.shape[0]
    # transition s-2 -> s allowed for labels that differ from the label two positions back
    skip = torch.zeros(S, dtype=torch.bool)
    if S > 2:
        skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
    em = lp[:, ext]                                                # [T, S] gathered emissions
    alpha = torch.full((S,), NEG, dtype=lp.dtype)
    init = [em[0, 0]] + ([em[0, 1]] if S > 1 else [])
    alpha = torch.cat([torch.stack(init), alpha[len(init):]])
    for t in range(1, T):
        a1 = torch.cat([alpha.new_full((1,), NEG), alpha[:-1]])
        a2 = torch.cat([alpha.new_full((2,), NEG), alpha[:-2]]) if S > 2 else alpha.new_full((S,), NEG)
        a2 = torch.where(skip, a2, a2.new_full((), NEG))
        stacked = torch.stack([alpha, a1, a2])                     # [3, S]
        m = stacked.max(0).values
        ok = m > NEG
        safe = torch.where(ok, m, torch.zeros_like(m))
        alpha = torch.where(ok, safe + torch.log(torch.exp(stacked - safe).sum(0)), m) + em[t]
        alpha = torch.where(ok, alpha, alpha.new_full((), NEG))
    tail = alpha[-2:] if S > 1 else alpha[-1:]
    m = tail.max()
    if not bool(m > NEG):
        return lp.new_full((), float("inf"))
    return -(m + torch.log(torch.exp(tail - m).sum()))


def ctc_loss_ref(logits, feat_lengths, targets, target_lengths, blank, reduction="mean"):
    """logits [B, T, V1] (any float dtype) -> the reference's loss: fp32 log-softmax, per-utterance nll, zero_infinity."""
    lp = torch.log_softmax(logits.float(), dim=-1)
    B = lp.shape[0]
    if targets.dim() == 1:                                         # concatenated form
        offs, rows = 0, []
        for n in target_lengths.tolist():
            rows.append(targets[offs:offs + n]); offs += n
    else:
        rows = [targets[i, : int(target_lengths[i])] for i in range(B)]
    nll = []
    for i in range(B):
        v = ctc_nll(lp[i, : int(feat_lengths[i])], rows[i], blank)
        nll.append(torch.zeros_like(v) if torch.isinf(v) else v)   # zero_infinity=True
    nll = torch.stack(nll)
    if reduction == "mean":
        return (nll / target_lengths.clamp_min(1).to(nll.dtype)).mean()
    if reduction == "sum":
        return nll.sum()
    return nll


class CTCHeadRef(nn.Module):
    """Same constructor / state_dict / methods as the reference's CTCHead."""

    def __init__(self, d_model: int, vocab_size: int, blank_id: int | None = None):
        super().__init__()
        self.vocab_size = vocab_size
        self.blank_id = vocab_size if blank_id is None else blank_id
        self.num_classes = vocab_size + 1
        self.proj = nn.Linear(d_model, self.num_classes)

    def forward(self, features):
        return self.proj(features)

    def log_probs(self, features):
        return torch.log_softmax(self.forward(features).float(), dim=-1)

    def loss(self, features, feat_lengths, targets, target_lengths, reduction="mean"):
        return ctc_loss_ref(self.forward(features), feat_lengths, targets, target_lengths, self.blank_id, reduction)

    @torch.no_grad()
    def frame_argmax(self, features):
        return self.forward(features).argmax(dim=-1)
