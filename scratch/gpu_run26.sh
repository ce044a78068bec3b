#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python scratch/kbench.py 2>/dev/null | grep "conv" 
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2t_bench_A.json 2> gpurun_out/r2t_bench_A.err
python - <<'PY'
import json
for f in ("r2t_bench_A",):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        kt = {r["kernel"]: r["ms"] for r in d["kernel_table"]}
        print(f, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["hot_path"]["ms_per_step"], "ssd_bwd", kt.get("ssd_bwd"), "conv_bwd", kt.get("conv_bwd"), "conv_fwd", kt.get("conv_fwd"), d["parity"]["fp32_feature_rel_err"], d["parity"]["bf16_loss_rel_err"])
    except Exception as e:
        print(f, "ERR", e)
PY
