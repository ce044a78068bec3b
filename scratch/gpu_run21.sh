#!/bin/bash
# span schedule of the fused SSD backward: kernel tests, timings, bench
timeout 1200 python -m pytest tests/test_gpu_mamba.py -q -x -k "ssd or conv or composite or block" 2>&1 | tail -4
timeout 300 python scratch/ssd_time.py 2>&1 | tail -4
HNB_SSD_SPAN=0 timeout 300 python scratch/ssd_time.py 2>&1 | tail -4
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2s_bench_A_span.json 2> gpurun_out/r2s_bench_A_span.err
HNB_SSD_SPAN=0 timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2s_bench_A_nospan.json 2> gpurun_out/r2s_bench_A_nospan.err
python - <<'PY'
import json
for f in ("r2s_bench_A_span", "r2s_bench_A_nospan"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        kt = {r["kernel"]: r["ms"] for r in d["kernel_table"]}
        print(f, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["hot_path"]["ms_per_step"], "ssd_bwd", kt.get("ssd_bwd"), "conv_bwd", kt.get("conv_bwd"), d["parity"]["fp32_feature_rel_err"], d["parity"]["bf16_loss_rel_err"])
    except Exception as e:
        print(f, "ERR", e)
PY
