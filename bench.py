#!/usr/bin/env python
"""Headline benchmark: encoder frames/sec, fwd+bwd, Type A Small N=2 (BASELINE.json), on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--seconds S]

A step = one DCASREncoder forward + backward (loss = mean(features^2) + 0.03*ratio_loss)
over one synthetic batch of 80-dim log-mel (B utterances x 16 s; 40 = the reference's per-GPU
batch_bins=64000 budget), bf16 autocast, random-init weights.  A frame = one valid 25 Hz encoder frame.
  value  : inputs resident in HBM.     e2e : pinned-host feats copied H2D every step + loss read back.
N > 1 (torchrun): the utterance batch is sharded (weak scaling), gradients are all-reduced over NCCL
inside the timed step, time = max over ranks.
--impl reference times the CPU restatement of the reference path (oracle/, kind "port") on host cores.
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


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU restatement of the reference path, fwd+bwd, on host cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference_step_fn(batch: int, seconds: float):
    from oracle.encoder_ref import EncoderRef
    torch.manual_seed(1)
    enc = EncoderRef(**SMALL)
    with torch.no_grad():                                  # same keep-fraction ~0.5 operating point as the GPU arm
        g = torch.Generator().manual_seed(7)
        enc.chunk.router.W_k.weight.copy_(torch.randn(384, 384, generator=g) / 384 ** 0.5)
    feats, lens = synth_batch(batch, seconds, 1)

    def step():
        enc.zero_grad(set_to_none=True)
        out = enc(feats, lens)
        loss = out.features.float().pow(2).mean() + 0.03 * out.ratio_loss
        loss.backward()
        return float(loss)

    return step, batch * sub_len(feats.shape[1])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1: use every host core
    step, frames = cpu_reference_step_fn(1, args.seconds)
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = frames * args.steps / dt
    line = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"Type A Small N=2 encoder fwd+bwd, 1 x {args.seconds:g} s utterance per step (bounded CPU sample)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{args.steps} steps x 1 utterance x {args.seconds:g} s, oracle/encoder_ref.py"},
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
def run_ours(args):
    import torch.distributed as dist
    import dcasr_b200 as dd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    torch.manual_seed(1)
    torch.backends.cudnn.benchmark = True           # ConvSubsampling4 is library code: let cuDNN pick its kernels
    enc = dd.DCASREncoder(**SMALL).to(dev)
    with torch.no_grad():
        # Untrained identity routers keep ~0.3 % of the frames (the main stack would run on M ~ 1).  The metric's
        # config is the TRAINED operating point, keep-fraction ~ 1/N = 0.5, so W_k is set to a seeded random matrix:
        # cos(q_t, k_{t-1}) is then ~N(0, 1/D) and half of the frames cross p >= 0.5 (SURVEY.md §8d, config 5 note).
        g = torch.Generator().manual_seed(7)
        enc.chunk.router.W_k.weight.copy_(torch.randn(384, 384, generator=g) / 384 ** 0.5)
    if world > 1:                                   # identical replicas, as DDP's initial broadcast would make them
        for p in enc.parameters():
            dist.broadcast(p.data, 0)
    params = [p for p in enc.parameters()]
    from dcasr_b200.distributed import GradAllReducer
    reducer = GradAllReducer(params, bucket_mb=32.0, overlap=True) if world > 1 else None
    feats_h, lens_h = synth_batch(args.batch, args.seconds, 1 + rank)      # each rank its own shard of utterances
    feats_pin, lens_pin = feats_h.pin_memory(), lens_h.pin_memory()
    feats_d, lens_d = feats_h.to(dev), lens_h.to(dev)
    frames_per_step = args.batch * sub_len(feats_h.shape[1]) * world
    kept = [0.0]

    def fwd_bwd(feats, lens):
        for p in params:
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = enc(feats, lens)
        loss = out.features.float().pow(2).mean() + 0.03 * out.ratio_loss
        loss.backward()
        if reducer is not None:                     # the one exchange of the path: gradient all-reduce (DDP semantics)
            reducer()
        kept[0] = out.kept_fractions[0]
        return loss

    def step_resident():
        return fwd_bwd(feats_d, lens_d)

    from dcasr_b200.distributed import HostBatchPrefetcher
    pref = HostBatchPrefetcher(dev)
    e2e_left = [0]                                  # steps still to run in the current e2e loop
    loss_pin = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(4)]
    e2e_pending, e2e_losses, e2e_seq, e2e_wall = [], [], [0], []

    def step_e2e():
        # every step copies its own batch from pinned host memory (exactly one H2D copy per step, all of them inside
        # the timed region); the copy for step i+1 is started on a side stream before step i's kernels are launched
        mode = int(os.environ.get("HNB_BENCH_E2E_MODE", "2"))     # diagnosis: 0 same-stream copy + blocking read, 1 prefetch + blocking read
        if mode == 0:
            f, l = feats_pin.to(dev, non_blocking=True), lens_pin.to(dev, non_blocking=True)
            e2e_losses.append(float(fwd_bwd(f, l)))
            return
        if pref.empty():
            pref.push(feats_pin, lens_pin)
        f, l = pref.pop()
        e2e_left[0] -= 1
        if e2e_left[0] > 0:
            pref.push(feats_pin, lens_pin)
        loss = fwd_bwd(f, l)
        if mode == 1:
            e2e_losses.append(float(loss))
            return
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

    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        xsub_d, lsub_d = enc.subsample(feats_d, lens_d)
    xsub_d = xsub_d.detach()

    def step_hot_path():                            # the path north_star names: everything after ConvSubsampling4
        for p in params:
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = enc.forward_hot_path(xsub_d, lsub_d)
        loss = out.features.float().pow(2).mean() + 0.03 * out.ratio_loss
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

    for _ in range(max(args.warmup, 3)):
        step_resident()
    with ClockSampler(local) as cs:
        sec, wall, launches = timed(step_resident, args.steps)
    clocks = cs.summary()
    import gc
    e2e_left[0] = 3
    for _ in range(3):
        step_e2e()
    e2e_finish()
    gc.collect()
    e2e_left[0] = args.steps
    del e2e_losses[:], e2e_wall[:]
    sec_e2e, _, _ = timed(step_e2e, args.steps, finish=e2e_finish)
    assert len(e2e_losses) == args.steps and all(math.isfinite(v) for v in e2e_losses), e2e_losses

    for _ in range(2):
        step_hot_path()
    sec_hot, _, launches_hot = timed(step_hot_path, args.steps)

    value = frames_per_step * args.steps / sec
    e2e_v = frames_per_step * args.steps / sec_e2e
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"Type A Small N=2 DCASREncoder fwd+bwd, {args.batch} x {args.seconds:g} s utterances per GPU "
                                   f"(L0={sub_len(feats_h.shape[1])} frames each), bf16 autocast, random-init weights",
                       "frames_per_step": frames_per_step, "kept_fraction": round(float(kept[0]), 4),
                       "includes_conv_subsample": "yes (cuDNN/cuBLAS via torch; outside the hand-written hot path)",
                       "optimizer": "none (metric is encoder fwd+bwd); N>1 adds the NCCL gradient all-reduce",
                       "l2": "no flush: the step's working set (saved activations, several GB) is far larger than the 126 MB L2",
                       "parallelism": f"dp{world} (utterance batch sharded, replicas)"},
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": (feats_h.numel() * 4 + lens_h.numel() * 8) * world,
                    "d2h_bytes_per_step": 4 * world,
                    "host_ms_between_steps": [round(1e3 * (b - a), 2) for a, b in zip(e2e_wall, e2e_wall[1:])],
                    "how": "per step: H2D of that step's pinned batch (side stream, started one step ahead) and an async "
                           "D2H copy of its loss into pinned memory, picked up by the host when it has landed (all of "
                           "them before the closing event)"},
            "hot_path": {"value": frames_per_step * args.steps / sec_hot, "unit": UNIT, "ms_per_step": 1e3 * sec_hot / args.steps,
                         "what": "forward_hot_path + backward from the subsampled features (ConvSubsampling4 excluded)",
                         "gpu_launches": launches_hot},
            "gpu_launches": launches, "clocks": clocks, "wall_s": round(wall, 3)}
    # the profiled step contains the gradient all-reduce when world > 1: EVERY rank must run it
    pk, pk_src = peaks()
    table, total = profile_step(step_resident, pk)
    if rank == 0:
        top = next((r for r in table if "frac" in r), None)
        if top:
            traffic = None      # DRAM read+write bytes per call of this entry point, from the committed ncu pass
            try:
                tj = json.load(open(os.path.join(REPO, "profiles", "r01_traffic.json")))
                traffic = round(tj["entries"][top["kernel"]]["dram_bytes_per_call"])
            except Exception:
                pass
            line["roofline"] = {"kernel": "hnb_" + top["kernel"], "bound": top["bound"], "achieved": top["achieved"],
                                "peak": top["peak"], "unit": top["unit"], "frac": top["frac"], "traffic": traffic,
                                "algorithmic_bytes_per_call": top.get("bytes_per_call"),
                                "peak_source": pk_src + (" (sustained)" if top["bound"] == "tensor" else ""),
                                "share_of_step": top["share"], "launches_per_step": top["launches"]}
        line["kernel_table"] = table[:12]
        if world == 1 and not args.no_cpu:
            torch.set_num_threads(os.cpu_count() or 1)
            step, frames = cpu_reference_step_fn(1, args.seconds)
            step()
            t0 = time.perf_counter()
            n = 0
            while n < 2 or time.perf_counter() - t0 < 12:
                step(); n += 1
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": frames * n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"{n} steps x 1 utterance x {args.seconds:g} s fwd+bwd, oracle/encoder_ref.py "
                                              f"(fp32, {os.cpu_count()} host cpus)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=40, help="utterances per GPU (40 x 16 s = batch_bins 64000)")
    ap.add_argument("--seconds", type=float, default=16.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
