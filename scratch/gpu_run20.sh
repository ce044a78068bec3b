#!/bin/bash
# programmatic dependent launch: full GPU suite, then the headline bench with / without the launch attribute
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2q_bench_A_pdl.json 2> gpurun_out/r2q_bench_A_pdl.err
HNB_PDL=0 timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2q_bench_A_nopdl.json 2> gpurun_out/r2q_bench_A_nopdl.err
python - <<'PY'
import json
for f in ("r2q_bench_A_pdl", "r2q_bench_A_nopdl"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["hot_path"]["ms_per_step"], d["kernel_ms_sum"], d["parity"]["fp32_feature_rel_err"], d["parity"]["bf16_loss_rel_err"])
    except Exception as e:
        print(f, "ERR", e)
PY
