#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 5 > gpurun_out/r2y_bench_A.json 2> gpurun_out/r2y_bench_A.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r2y_bench_A.json"))
print("A", round(j["value"]), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"]), "hot", round(j["hot_path"]["value"]))
for r in j["kernel_table"][:4]: print("   ", r["kernel"], r["launches"], r["ms"], r.get("frac"))
print(j["roofline"]["frac"], j["roofline"]["traffic"], j["ctc_head"])
PY
