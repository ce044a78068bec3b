"""How the seeds of tests/test_gpu_baseline_configs.py were chosen (run once, on CPU; nothing here runs in the tests).

north_star asks for bit-exact boundary masks wherever |p - 0.5| > 1e-4.  A boundary that flips INSIDE that band is
allowed, but it re-indexes every chunk after it, so activations can no longer be compared frame by frame.  At a
realistic operating point (keep fraction ~ 1/N) p is ~N(0.5, 0.025) and among ~1000 valid frames a few always sit within
1e-4 of the threshold; so for every full-size configuration this script searches router seeds (weights and inputs stay fixed) until the CPU
oracle's smallest |p - 0.5| over all valid frames of all chunk stages is at least MARGIN (> the band), which makes the
whole forward comparable without any escape hatch.  The tests re-assert the margin on the oracle at run time.

    python tests/golden/find_margin_seeds.py A_small_N2 | B_small_N4 | A_large_N3
"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]
from _util import fill_weights, set_router_operating_point  # noqa: E402
from baseline_cases import CASES, make_inputs  # noqa: E402
from oracle.encoder_ref import EncoderRef  # noqa: E402


@torch.no_grad()
def search(name, want, first=0):
    """Weights and inputs are fixed by the case's `seed`; only the routers' seed is searched (the enc stack's output is
    computed once, a candidate costs one router evaluation -- plus the mid stack for Type B's second stage)."""
    case = CASES[name]
    kw, shift, seed = case["kw"], case["shift"], case["seed"]
    ref = EncoderRef(**kw)
    fill_weights(ref, seed, router_identity=True)
    feats, lens = make_inputs(case, seed)
    x, l = ref.subsample(feats, lens)
    mask = torch.arange(x.shape[1])[None, :] < l[:, None]
    x_enc = ref.enc(x, l)
    t0 = time.time()
    for rs in range(first, 1000000):
        set_router_operating_point(ref, rs, shift)
        if kw["arch_type"] == "A":
            co = ref.chunk.chunk(x_enc, mask)
            m, kept = float((co.p[mask] - 0.5).abs().min()), [float(co.kept_fraction)]
        else:
            co1 = ref.chunk1.chunk(x_enc, mask)
            m, kept = float((co1.p[mask] - 0.5).abs().min()), [float(co1.kept_fraction)]
            if m >= want:
                z1 = ref.mid(ref.proj1_in(co1.z), co1.z_mask.sum(1))
                co2 = ref.chunk2.chunk(z1, co1.z_mask)
                m = min(m, float((co2.p[co1.z_mask] - 0.5).abs().min()))
                kept.append(float(co2.kept_fraction))
        target = 1.0 / (kw["N"] if kw["arch_type"] == "A" else kw["N"] ** 0.5)      # keep fraction the ratio loss steers to
        if any(abs(k - target) > 0.07 for k in kept):
            m = 0.0
        if rs % 200 == 0 or m >= want:
            print(f"{name} router_seed {rs}: min |p-0.5| = {m:.2e} kept {kept}  ({time.time() - t0:.0f} s)", flush=True)
        if m >= want:
            return rs


if __name__ == "__main__":
    name = sys.argv[1]
    search(name, float(sys.argv[2]) if len(sys.argv) > 2 else CASES[name]["margin"], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
