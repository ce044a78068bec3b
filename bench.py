#!/usr/bin/env python
"""Headline benchmark: encoder frames/sec, fwd+bwd, Type A Small N=2 (BASELINE.json), on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload W] [--mode train|decode] [--N n]

A step = one DCASREncoder forward + backward (loss = mean(features^2) + 0.03*ratio_loss)
over one synthetic batch of 80-dim log-mel (B utterances x 16 s; 40 = the reference's per-GPU
batch_bins=64000 budget), bf16 autocast, random-init weights.  A frame = one valid 25 Hz encoder frame.
  value  : inputs resident in HBM.     e2e : pinned-host feats copied H2D every step + loss read back.
N > 1 (torchrun): the utterance batch is sharded (weak scaling), gradients are all-reduced over NCCL
inside the timed step, time = max over ranks.
--impl reference times the CPU restatement of the reference path (oracle/, kind "port") on host cores.
--workload selects the other BASELINE.json configurations (B_small_N4, A_large_N3_60s, ragged); --mode decode times the
fp32 no_grad forward.  After the timed region utterance 0 goes through the CPU oracle: the line's `parity` object.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
for _p in (REPO, os.path.join(REPO, "h-net-mamba-asr_b200"), os.path.join(REPO, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "encoder frames/sec fwd+bwd, Type A Small N=2"
UNIT = "frames/s"
SMALL = dict(n_mels=80, d_outer=384, d_main=512, n_enc=4, n_main=12, n_dec=4, arch_type="A", N=2)
# BASELINE.json configs: [1] the headline (default), [2] Type B N=4, [3] Large Type A N=3 on 60 s, [4] ragged 2-35 s batches
WORKLOADS = {
    "A_small_N2": dict(kw=SMALL, seconds=16.0, batch=40, what="Type A Small N=2"),
    "B_small_N4": dict(kw=dict(n_mels=80, d_outer=384, d_main=512, n_enc=4, n_main=12, n_dec=4, n_mid=4, arch_type="B", N=4),
                       seconds=16.0, batch=40, what="Type B Small N=4 (two sqrt(N)=2 chunk stages)"),
    "A_large_N3_60s": dict(kw=dict(n_mels=80, d_outer=512, d_main=768, n_enc=6, n_main=18, n_dec=6, arch_type="A", N=3),
                           seconds=60.0, batch=10, what="Type A Large N=3, 60 s utterances (docs/experimental_plan.md:123)"),
    "ragged": dict(kw=SMALL, seconds=None, batch=None, what="Type A Small, ragged 2-35 s batches formed by the reference's bucketing rule"),
}
BATCH_BINS = 64000          # padded 100 Hz frames per GPU batch (reference configs/typeA_small_N2.yaml:71)


def n_frames_100hz(seconds: float) -> int:
    return 1 + (int(16000 * seconds) - 400) // 160            # reference data/librispeech.py:30-32


def sub_len(t: int) -> int:
    return ((t - 1) // 2 - 1) // 2


def peaks():
    try:
        return json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML, 10 ms period; nvidia-smi as fallback)."""
    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index, self.sm, self.max_mhz, self.reasons, self.stop = index, [], None, set(), False
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop:
                self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.reasons |= {n for bit, n in self.NAMES.items() if mask & bit}
                time.sleep(0.01)
        except Exception:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            while not self.stop:
                try:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    r = [c.strip() for c in out.strip().split(",")]
                    if len(r) >= 6 and r[0].isdigit():
                        self.sm.append(int(r[0])); self.max_mhz = int(r[1])
                        self.reasons |= {n for n, v in zip(names, r[2:6]) if v.lower().startswith("active")}
                except Exception:
                    pass
                time.sleep(0.05)

    def __enter__(self):
        self.t.start()
        time.sleep(0.05)
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(sm)}


def synth_batch(B: int, seconds: float, seed: int):
    g = torch.Generator().manual_seed(seed)
    T = n_frames_100hz(seconds)
    return torch.randn(B, T, 80, generator=g), torch.full((B,), T, dtype=torch.int64)


def bucket_batches(lengths, max_frames):
    """Length-bucketed batches under a padded-frame budget: sort by length, then greedily extend the current batch while
    (count + 1) * longest <= max_frames -- the rule of the reference's DistributedBucketBatchSampler
    (src/dcasr/data/librispeech.py:174-187), restated."""
    order = sorted(range(len(lengths)), key=lambda i: lengths[i])
    out, cur, longest = [], [], 0
    for i in order:
        cand = max(longest, lengths[i])
        if cur and (len(cur) + 1) * cand > max_frames:
            out.append(cur)
            cur, longest = [i], lengths[i]
        else:
            cur.append(i)
            longest = cand
    if cur:
        out.append(cur)
    return out


def ragged_batches(rank: int, n_utts: int = 4000, n_pick: int = 12):
    """BASELINE config 5: durations 35 * Beta(2.2, 4.0) clipped to [2, 35] s (LibriSpeech-like, mean ~12.4 s), bucketed
    with the reference's rule under batch_bins = 64000; every rank gets its own `n_pick` batches, evenly spread over
    the sorted batch list (so short- and long-utterance batches are both in the sample).  -> [(feats, lens)], stats."""
    g = torch.Generator().manual_seed(1234)
    # (torch.distributions has no generator argument: draw Beta(a, b) = Ga / (Ga + Gb) from seeded Gamma variates)
    ga = torch._standard_gamma(torch.full((n_utts,), 2.2), generator=g)
    gb = torch._standard_gamma(torch.full((n_utts,), 4.0), generator=g)
    dur = (35.0 * ga / (ga + gb)).clamp(2.0, 35.0)
    lens = [n_frames_100hz(float(d)) for d in dur]
    batches = bucket_batches(lens, BATCH_BINS)
    idx = [int(round(k * (len(batches) - 1) / (n_pick - 1))) for k in range(n_pick)]
    idx = [(i + rank) % len(batches) for i in idx]
    out = []
    for bi in idx:
        ls = torch.tensor([lens[i] for i in batches[bi]], dtype=torch.int64)
        gg = torch.Generator().manual_seed(100 + bi)
        f = torch.randn(len(ls), int(ls.max()), 80, generator=gg)
        for r, n in enumerate(ls.tolist()):
            f[r, n:] = 0.0
        out.append((f, ls))
    valid = sum(int(sub_len_t(l).sum()) for _, l in out)
    padded = sum(f.shape[0] * sub_len(f.shape[1]) for f, _ in out)
    return out, {"utterance_pool": n_utts, "mean_seconds": round(float(dur.mean()), 2), "batches_in_pool": len(batches),
                 "batches_sampled": n_pick, "batch_shapes": [[f.shape[0], f.shape[1]] for f, _ in out],
                 "valid_frames": valid, "padded_frames": padded, "padding_share": round(1 - valid / padded, 4)}


def sub_len_t(t):
    return (((t - 1) // 2 - 1) // 2).clamp_min(0)


def set_routers(enc, keep_N):
    """Untrained identity routers keep ~0.3 % of the frames (the main stack would run on M ~ 1).  The metric's config is the
    TRAINED operating point, keep fraction ~ 1/N per stage: W_k = shift I + seeded Gaussian / sqrt(d) puts
    cos(q_t, k_{t-1}) ~ N(c, 1/d); shift 0 -> ~0.5, 0.12 -> ~1/3, 0.2 -> ~1/4 (SURVEY.md §8d, config 5 note)."""
    shift = {1: 0.0, 2: 0.0, 3: 0.12, 4: 0.2}.get(int(round(keep_N)), 0.0)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for name in ("chunk", "chunk1", "chunk2"):
            ch = getattr(enc, name, None)
            if ch is not None and getattr(ch, "router", None) is not None:
                d = ch.router.W_k.weight.shape[0]
                w = shift * torch.eye(d) + torch.randn(d, d, generator=g) / d ** 0.5
                ch.router.W_k.weight.copy_(w.to(ch.router.W_k.weight.device))


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU restatement of the reference path, fwd+bwd, on host cores
# ---------------------------------------------------------------------------------------------------
def metric_name(workload: str, mode: str) -> str:
    if workload == "A_small_N2" and mode == "train":
        return METRIC
    what = {"A_small_N2": "Type A Small N=2", "B_small_N4": "Type B Small N=4", "A_large_N3_60s": "Type A Large N=3 60 s",
            "ragged": "Type A Small ragged 2-35 s"}[workload]
    return f"encoder frames/sec {'fwd+bwd' if mode == 'train' else 'fwd (fp32 decode)'}, {what}"


def cpu_reference_step_fn(kw, batch: int, seconds: float, keep_N, mode: str = "train"):
    """One step of the CPU restatement of the reference path (oracle/encoder_ref.py: the reference's encoder.py /
    mamba_block.py / hnet_chunk.py restated as vectorised torch code + the Mamba2 restatement), fp32, all host threads."""
    from oracle.encoder_ref import EncoderRef
    torch.manual_seed(1)
    enc = EncoderRef(**kw)
    set_routers(enc, keep_N)                                # same keep-fraction operating point as the GPU arm
    feats, lens = synth_batch(batch, seconds, 1)

    def step():
        if mode != "train":
            with torch.no_grad():
                return float(enc(feats, lens).features.float().pow(2).mean())
        enc.zero_grad(set_to_none=True)
        out = enc(feats, lens)
        loss = out.features.float().pow(2).mean() + 0.03 * out.ratio_loss
        loss.backward()
        return float(loss.detach())

    return step, batch * sub_len(feats.shape[1])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1: use every host core
    wl = WORKLOADS[args.workload]
    kw = dict(wl["kw"]); kw["N"] = args.N if args.N else kw["N"]
    seconds = args.seconds or wl["seconds"] or 12.4      # ragged: one utterance of the distribution's mean length
    step, frames = cpu_reference_step_fn(kw, 1, seconds, kw["N"] if kw["arch_type"] == "A" else kw["N"] ** 0.5, args.mode)
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = frames * args.steps / dt
    what = "fwd+bwd" if args.mode == "train" else "fwd (no_grad)"
    line = {"metric": metric_name(args.workload, args.mode), "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{wl['what']} encoder {what}, 1 x {seconds:g} s utterance per step (bounded CPU sample)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{args.steps} steps x 1 utterance x {seconds:g} s, oracle/encoder_ref.py (vectorised "
                                       f"torch restatement of the reference modules; Mamba-2 arithmetic as chunk-64 einsums)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# roofline of the dominant kernel from a profiled step
# ---------------------------------------------------------------------------------------------------
def algorithmic_work(name: str, a: tuple):
    """(flops, bytes) of ONE launch of a C-ABI entry point: flops in the reference's convention, bytes = every input
    read once + every output written once (saved-for-backward by-products such as the SSD chunk
    states are NOT counted).  `a` holds the scalar arguments of the C call in declaration order (include/hnet_b200.h).
    Which roofline binds is decided from these two numbers and the measured peaks, not asserted per kernel."""
    if name == "gemm_bf16":       # lda, transA, ldb, transB, M, N, K, ldr, ldc, c_dtype, splitk
        M, N, K = a[4], a[5], a[6]
        csz = 4 if a[9] == 0 else 2
        return 2.0 * M * N * K, 2.0 * (M * K + N * K) + csz * M * N
    if name in ("ssd_fwd", "ssd_bwd"):   # dtype, ndir, B, L, di, N, H, impl
        ndir, B, L, di, N, H = a[1], a[2], a[3], a[4], a[5], a[6]
        T, C = ndir * B * L, di + 2 * N
        f = 4.0 * di * N * T                     # linear recurrence, reference convention (eval/efficiency.py:135)
        if name == "ssd_fwd":
            return f, T * (C * 2 + di * 2 + H * 4)                    # xconv, dt -> y
        return 2.5 * f, T * (di * 2 + C * 2 + H * 4 + di * 2 + 2 * N * 2 + H * 4)   # dy, xconv, dt -> dxc, dBC, ddt
    if name == "conv_fwd":        # dtype, ldz, dstride, ndir, B, L, di, N, H
        ndir, B, L, di, N, H = a[3:9]
        return None, ndir * B * L * ((di + 2 * N) * 4 + H * 6)
    if name == "conv_bwd":
        ndir, B, L, di, N, H = a[3:9]
        return None, ndir * B * L * ((di + 2 * N) * 4 + di * 2 + 2 * N * 4 + H * 10)
    if name == "gated_norm_fwd":  # dtype, ldz, dstride, ndir, B, L, di, eps
        ndir, B, L, di = a[3:7]
        return None, ndir * B * L * di * 6
    if name == "gated_norm_bwd":
        ndir, B, L, di = a[3:7]
        return None, ndir * B * L * di * 10
    if name == "layernorm_fwd":   # x_dtype, rows, d, eps, y_dtype
        return None, a[1] * a[2] * 6
    if name == "layernorm_bwd":   # dy_dtype, x_dtype, rows, d, dx_dtype
        return None, a[2] * a[3] * 12
    return None, None


def profile_step(step, pk):
    from dcasr_b200 import _lib
    _lib.profile_start()
    step()
    rec = _lib.profile_stop()
    agg = {}
    pf, pb = pk["bf16_tflops_sustained"] * 1e12, pk["hbm_gbs"] * 1e9
    for name, a, ms in rec:
        e = agg.setdefault(name, {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0, "ok": True})
        f, by = algorithmic_work(name, a)
        e["ms"] += ms; e["n"] += 1
        if by is None:
            e["ok"] = False
        else:
            e["flops"] += f or 0.0; e["bytes"] += by
    total = sum(e["ms"] for e in agg.values())
    table = []
    for k, e in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        row = {"kernel": k, "launches": e["n"], "ms": round(e["ms"], 3), "share": round(e["ms"] / total, 4)}
        if e["ok"]:
            row["bytes_per_call"] = round(e["bytes"] / e["n"])
        if e["ok"] and e["ms"] > 0:
            sec = e["ms"] * 1e-3
            t_f, t_b = e["flops"] / pf, e["bytes"] / pb          # time each roofline alone would allow
            if t_f > t_b:
                ach, peak = e["flops"] / sec / 1e12, pk["bf16_tflops_sustained"]
                row.update(bound="tensor", achieved=round(ach, 2), peak=peak, unit="TFLOP/s", frac=round(ach / peak, 4))
            else:
                ach, peak = e["bytes"] / sec / 1e9, pk["hbm_gbs"]
                row.update(bound="hbm", achieved=round(ach, 1), peak=peak, unit="GB/s", frac=round(ach / peak, 4))
        table.append(row)
    return table, total


# ---------------------------------------------------------------------------------------------------
def ctc_head_leg(dev, B=40, T=398, d=384, V=500, U=60, reps=10):
    """The step after the path: CTC head fwd+bwd on the encoder's output shape, bf16 autocast -- the drop-in
    (projection GEMM + hnb_ctc_lse / alpha_beta / grad / col_sum: no fp32 logits, no [B, L, V+1] fp32 log-probabilities) timed
    beside the reference's own op sequence on the same GPU (nn.Linear -> .float() -> log_softmax -> F.ctc_loss, ctc.py:100-115)."""
    import torch.nn.functional as F
    import dcasr_b200 as dd
    torch.manual_seed(3)
    head = dd.CTCHead(d, V).to(dev)
    x = torch.randn(B, T, d, device=dev)
    fl = torch.full((B,), T, device=dev, dtype=torch.int64)
    tl = torch.randint(U // 2, U + 1, (B,), device=dev)
    tg = torch.randint(0, V, (B, U), device=dev)

    def ours():
        xx = x.requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = head.loss(xx, fl, tg, tl)
        loss.backward()
        return loss

    def ref_ops():
        xx = x.requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            lp = F.log_softmax(F.linear(xx, head.proj.weight, head.proj.bias).float(), dim=-1).transpose(0, 1)
            loss = F.ctc_loss(lp, tg, fl, tl, blank=V, reduction="mean", zero_infinity=True)
        loss.backward()
        return loss

    out = {}
    for name, fn in (("ours_ms", ours), ("reference_ops_same_gpu_ms", ref_ops)):
        for _ in range(3):
            l = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            l = fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = round(e0.elapsed_time(e1) / reps, 4)
        out[name.replace("_ms", "_loss")] = round(float(l), 5)
    out["shape"] = f"{B} x {T} frames, d {d}, {V} pieces + blank, targets <= {U}"
    return out


def parity_check(enc, kw, feats1, lens1, dev, mode):
    """After the timed region: utterance 0 of the bench batch through the CPU oracle (fp32) and through the product, once
    in fp32 (exact kernels; features at north_star's 1e-3, boundaries bit-exact) and once as timed (bf16 autocast; loss
    against the fp32 oracle, boundary decisions compared outside the bf16 band)."""
    from oracle.encoder_ref import EncoderRef
    ref = EncoderRef(**kw)
    ref.load_state_dict({k: v.detach().float().cpu() for k, v in enc.state_dict().items()})
    with torch.no_grad():
        r = ref(feats1, lens1)
        loss_r = float(r.features.pow(2).mean() + 0.03 * r.ratio_loss)
        tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        o32 = enc(feats1.to(dev), lens1.to(dev))
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ob = enc(feats1.to(dev), lens1.to(dev))
        loss_b = float(ob.features.float().pow(2).mean() + 0.03 * ob.ratio_loss)
    den = float(r.features.double().norm())
    same32 = all(torch.equal(b.cpu(), br) for (_, b), (_, br) in zip(o32.boundaries, r.boundaries))
    margin = min(float((p[p > 0] - 0.5).abs().min()) for p, _ in r.boundaries) if r.boundaries else None
    out = {"oracle": "oracle/encoder_ref.py fp32 on host, utterance 0 of the batch, forward",
           "fp32_boundaries_equal": same32, "oracle_min_abs_p_minus_half": margin,
           "fp32_feature_rel_err": (float((o32.features.cpu().double() - r.features.double()).norm()) / den) if same32 else None,
           "bf16_loss": loss_b, "oracle_loss": loss_r, "bf16_loss_rel_err": abs(loss_b - loss_r) / abs(loss_r)}
    if all(bb.shape == br.shape and torch.equal(bb.cpu(), br) for (_, bb), (_, br) in zip(ob.boundaries, r.boundaries)):
        # the loss sits behind a LayerNorm (~1 by construction): the feature error is the informative number.  This is
        # bf16 against FP32 truth after 20 blocks; the bf16-vs-bf16 bar of north_star (2e-2) is asserted in
        # tests/test_gpu_baseline_configs.py against the oracle run under the same autocast.
        out["bf16_feature_rel_err_vs_fp32_oracle"] = float((ob.features.float().cpu().double() - r.features.double()).norm()) / den
    else:
        out["bf16_feature_rel_err_vs_fp32_oracle"] = None      # a boundary inside the bf16 band flipped: frames not comparable
    bad = 0
    for (pb, bb), (pr, br) in zip(ob.boundaries, r.boundaries):
        if pb.shape == pr.shape:
            outside = (pr - 0.5).abs() > 2e-2
            bad += int((bb.cpu()[outside] != br[outside]).sum())
        else:
            bad = None
            break
    out["bf16_boundary_mismatches_outside_2e-2_band"] = bad
    return out


def run_ours(args):
    import torch.distributed as dist
    import dcasr_b200 as dd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")               # its banner goes to stdout: keep stdout to the one JSON line
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    wl = WORKLOADS[args.workload]
    kw = dict(wl["kw"])
    if args.N:
        kw["N"] = args.N
    train = args.mode == "train"
    ragged = args.workload == "ragged"
    torch.manual_seed(1)
    torch.backends.cudnn.benchmark = not ragged     # ConvSubsampling4 is library code: let cuDNN pick its kernels (fixed shapes only)
    enc = dd.DCASREncoder(**kw).to(dev)
    set_routers(enc, kw["N"] if kw["arch_type"] == "A" else kw["N"] ** 0.5)
    if world > 1:                                   # identical replicas, as DDP's initial broadcast would make them
        for p in enc.parameters():
            dist.broadcast(p.data, 0)
    params = [p for p in enc.parameters()]
    from dcasr_b200.distributed import GradAllReducer
    reducer = (GradAllReducer(params, bucket_mb=float(os.environ.get("HNB_BUCKET_MB", "32")),
                              overlap=os.environ.get("HNB_REDUCER_OVERLAP", "1") != "0") if (world > 1 and train) else None)
    # ---- the batches: one fixed-length batch (each rank its own shard of utterances), or a cycle of ragged batches
    if ragged:
        host_batches, ragged_stats = ragged_batches(rank)
    else:
        batch = args.batch or wl["batch"]
        seconds = args.seconds or wl["seconds"]
        host_batches, ragged_stats = [synth_batch(batch, seconds, 1 + rank)], None
    pinned = [(f.pin_memory(), l.pin_memory()) for f, l in host_batches]
    resident = [(f.to(dev), l.to(dev)) for f, l in host_batches]
    frames_of = [int(sub_len_t(l).sum()) for _, l in host_batches]
    nb = len(host_batches)
    kept_log, exposed = [], []
    cursor = {"resident": 0, "e2e": 0, "hot": 0}

    def loss_of(out, lens_sub):
        if ragged:                                  # masked mean: padded frames carry no loss (as CTC / the ratio loss do)
            m = (torch.arange(out.features.shape[1], device=dev)[None, :] < lens_sub[:, None]).unsqueeze(-1)
            return (out.features.float().pow(2) * m).sum() / (m.sum() * out.features.shape[2]) + 0.03 * out.ratio_loss
        return out.features.float().pow(2).mean() + 0.03 * out.ratio_loss

    def fwd_bwd(feats, lens):
        if not train:                               # decode: fp32, no_grad forward (reference tasks/decode_task.py:123-151)
            with torch.no_grad():
                out = enc(feats, lens)
                kept_log.append(out.kept_fractions)
                return out.features.float().pow(2).mean()
        for p in params:
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = enc(feats, lens)
        loss = loss_of(out, out.lengths)
        loss.backward()
        if reducer is not None:                     # the one exchange of the path: gradient all-reduce (DDP semantics)
            if len(exposed) < 64:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); reducer(); e1.record()
                exposed.append((e0, e1))
            else:
                reducer()
        kept_log.append(out.kept_fractions)
        return loss

    def step_resident():
        i = cursor["resident"] % nb
        cursor["resident"] += 1
        return fwd_bwd(*resident[i])

    from dcasr_b200.distributed import HostBatchPrefetcher
    pref = HostBatchPrefetcher(dev)
    e2e_left = [0]                                  # steps still to run in the current e2e loop
    loss_pin = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(4)]
    e2e_pending, e2e_losses, e2e_seq, e2e_wall = [], [], [0], []

    def e2e_push():
        i = cursor["e2e"] % nb
        cursor["e2e"] += 1
        pref.push(*pinned[i])

    def step_e2e():
        # every step copies its own batch from pinned host memory (exactly one H2D copy per step, all of them inside
        # the timed region); the copy for step i+1 is started on a side stream before step i's kernels are launched
        if pref.empty():
            e2e_push()
        f, l = pref.pop()
        e2e_left[0] -= 1
        if e2e_left[0] > 0:
            e2e_push()
        loss = fwd_bwd(f, l)
        # device -> host read of the step's result, every step: an asynchronous copy into pinned memory that the host
        # picks up as soon as it has landed (polled once per step, never waited for; whatever is still in flight is
        # drained in e2e_finish, inside the timed region), so the read never drains the launch queue -- the reference's
        # trainer keeps its loss on the device for the same reason (src/dcasr/training/trainer.py:250,301)
        slot = loss_pin[e2e_seq[0] % len(loss_pin)]
        e2e_seq[0] += 1
        slot.copy_(loss.detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        e2e_pending.append((slot, ev))
        while e2e_pending and (e2e_pending[0][1].query() or len(e2e_pending) >= len(loss_pin)):
            s0, ev0 = e2e_pending.pop(0)
            ev0.synchronize()
            e2e_losses.append(float(s0))
        e2e_wall.append(time.perf_counter())

    def e2e_finish():
        while e2e_pending:
            s0, ev0 = e2e_pending.pop(0)
            ev0.synchronize()
            e2e_losses.append(float(s0))

    hot_inputs = []
    if train:
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            for f, l in resident:
                xs, ls = enc.subsample(f, l)
                hot_inputs.append((xs.detach(), ls))

    def step_hot_path():                            # the path north_star names: everything after ConvSubsampling4
        i = cursor["hot"] % nb
        cursor["hot"] += 1
        xs, ls = hot_inputs[i]
        for p in params:
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = enc.forward_hot_path(xs, ls)
        loss = loss_of(out, ls)
        loss.backward()
        if reducer is not None:
            reducer()
        return loss

    def timed(step, steps, finish=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dd.reset_launch_count()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            step()
        if finish is not None:
            finish()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / 1e3, wall, dd.launch_count()

    steps = args.steps if not ragged else max(args.steps, nb) // nb * nb      # ragged: whole cycles over the sampled batches
    warm = max(args.warmup, 3) if not ragged else max(args.warmup, nb)        # ragged: every batch shape once before timing
    frames_timed = lambda k0: sum(frames_of[(k0 + i) % nb] for i in range(steps)) * world      # noqa: E731
    for _ in range(warm):
        step_resident()
    del exposed[:], kept_log[:]
    k0 = cursor["resident"]
    with ClockSampler(local) as cs:
        sec, wall, launches = timed(step_resident, steps)
    clocks = cs.summary()
    frames_value = frames_timed(k0)
    exposed_ms = [a.elapsed_time(b) for a, b in exposed]
    kept = [[round(float(v), 4) for v in ks] for ks in kept_log[:nb]]
    import gc
    e2e_left[0] = 3
    for _ in range(3):
        step_e2e()
    e2e_finish()
    gc.collect()
    e2e_left[0] = steps
    del e2e_losses[:], e2e_wall[:]
    cursor["e2e"] = 0
    sec_e2e, _, _ = timed(step_e2e, steps, finish=e2e_finish)
    assert len(e2e_losses) == steps and all(math.isfinite(v) for v in e2e_losses), e2e_losses
    frames_e2e = frames_timed(0)
    if train:
        for _ in range(2 if not ragged else nb):
            step_hot_path()
        kh = cursor["hot"]
        sec_hot, _, launches_hot = timed(step_hot_path, steps)
        frames_hot = frames_timed(kh)

    f0, l0 = host_batches[0]
    value = frames_value / sec
    e2e_v = frames_e2e / sec_e2e
    h2d = sum(f.numel() * 4 + l.numel() * 8 for f, l in host_batches) / nb * world
    shape = (f"{f0.shape[0]} x {(args.seconds or wl['seconds']):g} s utterances per GPU (L0={sub_len(f0.shape[1])} frames each)"
             if not ragged else f"{nb} ragged batches per GPU cycled (2-35 s, batch_bins {BATCH_BINS})")
    line = {"metric": metric_name(args.workload, args.mode), "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 * sec / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if train else "f32", "data": "synthetic",
            "config": {"workload": f"{wl['what']}{'' if not args.N else f' (N={args.N})'} DCASREncoder "
                                   f"{'fwd+bwd, bf16 autocast' if train else 'forward, fp32 no_grad (decode)'}, {shape}, random-init weights",
                       "frames_per_step": frames_value / steps, "kept_fraction": kept[0] if kept else None,
                       "includes_conv_subsample": "yes (cuDNN/cuBLAS via torch; outside the hand-written hot path)",
                       "optimizer": "none (metric is encoder fwd+bwd); N>1 adds the NCCL gradient all-reduce",
                       "l2": "no flush: the step's working set (saved activations, several GB) is far larger than the 126 MB L2",
                       "parallelism": f"dp{world} (utterance batch sharded, replicas)"},
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4 * world,
                    "host_ms_between_steps": [round(1e3 * (b - a), 2) for a, b in zip(e2e_wall, e2e_wall[1:])][:16],
                    "how": "per step: H2D of that step's pinned batch (side stream, started one step ahead) and an async "
                           "D2H copy of its loss into pinned memory, picked up by the host when it has landed (all of "
                           "them before the closing event)"},
            "gpu_launches": launches, "clocks": clocks, "wall_s": round(wall, 3)}
    if train:
        line["hot_path"] = {"value": frames_hot / sec_hot, "unit": UNIT, "ms_per_step": 1e3 * sec_hot / steps,
                            "what": "forward_hot_path + backward from the subsampled features (ConvSubsampling4 excluded)",
                            "gpu_launches": launches_hot}
    if ragged:
        line["config"]["ragged"] = ragged_stats
        line["config"]["kept_fraction_per_batch"] = kept
    if exposed_ms:
        line["allreduce_exposed_ms"] = round(sum(exposed_ms) / len(exposed_ms), 3)
    # the profiled step contains the gradient all-reduce when world > 1: EVERY rank must run it
    pk, pk_src = peaks()
    table, total = profile_step(step_resident, pk)
    if rank == 0:
        top = next((r for r in table if "frac" in r), None)
        if top:
            traffic = None      # DRAM read+write bytes per call of this entry point, from the committed ncu pass
            import glob
            for tf in sorted(glob.glob(os.path.join(REPO, "profiles", "r*_traffic.json")), reverse=True):
                try:
                    tj = json.load(open(tf))
                    if tj.get("workload", "A_small_N2") == args.workload and args.mode == "train":
                        traffic = round(tj["entries"][top["kernel"]]["dram_bytes_per_call"])
                        break
                except Exception:
                    pass
            line["roofline"] = {"kernel": "hnb_" + top["kernel"], "bound": top["bound"], "achieved": top["achieved"],
                                "peak": top["peak"], "unit": top["unit"], "frac": top["frac"], "traffic": traffic,
                                "algorithmic_bytes_per_call": top.get("bytes_per_call"),
                                "peak_source": pk_src + (" (sustained)" if top["bound"] == "tensor" else ""),
                                "share_of_step": top["share"], "launches_per_step": top["launches"]}
        line["kernel_table"] = table[:12]
        line["kernel_ms_sum"] = round(total, 3)
        if world == 1 and not args.no_parity:
            try:
                line["parity"] = parity_check(enc, kw, f0[:1, : int(l0[0])].contiguous(), l0[:1], dev, args.mode)
            except Exception as e:            # a parity failure must be visible in the line, never hide the measurement
                line["parity"] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1 and train and not args.no_parity:
            try:
                line["ctc_head"] = ctc_head_leg(dev)
            except Exception as e:
                line["ctc_head"] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1 and not args.no_cpu:
            torch.set_num_threads(os.cpu_count() or 1)
            secs = args.seconds or wl["seconds"] or 12.4
            step, frames = cpu_reference_step_fn(kw, 1, secs, kw["N"] if kw["arch_type"] == "A" else kw["N"] ** 0.5, args.mode)
            step()
            t0 = time.perf_counter()
            n = 0
            while n < 2 or time.perf_counter() - t0 < 12:
                step(); n += 1
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": frames * n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"{n} steps x 1 utterance x {secs:g} s {'fwd+bwd' if train else 'fwd'}, oracle/encoder_ref.py "
                                              f"(vectorised torch restatement of the reference modules, fp32, {os.cpu_count()} host cpus)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="A_small_N2", choices=sorted(WORKLOADS),
                    help="BASELINE.json config: A_small_N2 (headline, default), B_small_N4, A_large_N3_60s, ragged")
    ap.add_argument("--mode", default="train", choices=["train", "decode"],
                    help="train: fwd+bwd under bf16 autocast (the metric); decode: fp32 no_grad forward")
    ap.add_argument("--N", type=float, default=0, help="override the compression ratio (ragged sweep: 1, 2, 3, 4)")
    ap.add_argument("--batch", type=int, default=0, help="utterances per GPU (default: batch_bins 64000 / frames per utterance)")
    ap.add_argument("--seconds", type=float, default=0.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-bench oracle check of utterance 0")
    args = ap.parse_args()
    if args.N and args.N == int(args.N):
        args.N = int(args.N)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
