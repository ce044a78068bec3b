"""The full-size BASELINE.json configurations as parity-test cases (shared by tests/test_gpu_baseline_configs.py, the
seed search tests/golden/find_margin_seeds.py and bench.py's in-bench parity check).

kw      : DCASREncoder constructor arguments (reference configs/typeA_small_N2.yaml; Type B = the same with arch_type B,
          N 4, n_mid 4, tasks/asr_task.py:34; Large = docs/experimental_plan.md:123)
seconds : utterance durations (ragged); T = 1 + (16000 s - 400) // 160 frames of 80-dim log-mel (data/librispeech.py:30-32)
shift   : identity share of the router's W_k (tests/_util.py:set_router_operating_point): moves the keep fraction to ~1/N
seed    : weights (tests/_util.py:fill_weights) and inputs
router_seed : chosen by find_margin_seeds.py so that the oracle's smallest |p - 0.5| is >= margin (> north_star's 1e-4 band)
"""
import torch

CASES = {
    "A_small_N2": dict(kw=dict(n_mels=80, d_outer=384, d_main=512, n_enc=4, n_main=12, n_dec=4, arch_type="A", N=2),
                       seconds=[16.0, 12.1, 6.07], shift=0.0, margin=3e-4, seed=31, router_seed=7646),
    "B_small_N4": dict(kw=dict(n_mels=80, d_outer=384, d_main=512, n_enc=4, n_main=12, n_dec=4, n_mid=4, arch_type="B", N=4),
                       seconds=[16.0, 9.3], shift=0.0, margin=2e-4, seed=31, router_seed=11684),
    "A_large_N3": dict(kw=dict(n_mels=80, d_outer=512, d_main=768, n_enc=6, n_main=18, n_dec=6, arch_type="A", N=3),
                       seconds=[60.0], shift=0.12, margin=1.3e-4, seed=31, router_seed=2322),
}


def n_frames_100hz(seconds: float) -> int:
    return 1 + (int(16000 * seconds) - 400) // 160


def make_inputs(case, seed):
    lens = torch.tensor([n_frames_100hz(s) for s in case["seconds"]])
    g = torch.Generator().manual_seed(1000 + seed)
    feats = torch.randn(len(lens), int(lens.max()), 80, generator=g)
    for i, n in enumerate(lens.tolist()):
        feats[i, n:] = 0.0                                    # right padding, as the reference's collate pads (zeros)
    return feats, lens
