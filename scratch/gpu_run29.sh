#!/bin/bash
# packed fp32 math (FFMA2) in conv1d / gated norm / ConvSubsampling4 front end: targeted parity tests, isolated kernel times, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mamba.py -q -x -k "conv or norm or subsample or block or stack or encoder" 2>&1 | tail -3
timeout 300 python scratch/conv_bench.py 2>/dev/null | grep -v layernorm
timeout 400 python bench.py --steps 10 --warmup 5 > gpurun_out/r2u_bench_A.json 2> gpurun_out/r2u_bench_A.err
HNB_SUB_BWD_MINB=3 timeout 400 python bench.py --steps 10 --warmup 5 --no-cpu > gpurun_out/r2u_bench_A_minb3.json 2> gpurun_out/r2u_bench_A_minb3.err
python - <<'PY'
import json
for f in ("r2u_bench_A", "r2u_bench_A_minb3"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        kt = {r["kernel"]: r["ms"] for r in d["kernel_table"]}
        print(f, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["hot_path"]["ms_per_step"], {k: kt.get(k) for k in ("conv_bwd", "conv_fwd", "gated_norm_bwd", "gated_norm_fwd", "subsample_conv1_fwd", "subsample_conv1_bwd")}, d["parity"]["fp32_feature_rel_err"], d["parity"]["bf16_loss_rel_err"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r2u_bench_A.err
