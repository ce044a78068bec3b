#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_mamba.py -q -rf -k "repeatable or composite_equals or ssd_tcgen05_forward" 2>&1 | tail -2; done
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -5
timeout 600 python bench.py --mode decode --steps 5 --warmup 3 --no-cpu > gpurun_out/r2f_decode.json 2> gpurun_out/r2f_decode.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r2f_decode.json"))
print("decode", round(j["value"]), round(j["ms_per_step"],3), j["dtype"])
for r in j["kernel_table"][:4]: print("   ", r["kernel"], r["launches"], r["ms"], r.get("frac"), r.get("bound"))
print(j["parity"]["fp32_feature_rel_err"], j["parity"]["fp32_boundaries_equal"])
PY
python scratch/ssd_time.py | head -4
