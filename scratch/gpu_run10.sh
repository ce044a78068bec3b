#!/bin/bash
mkdir -p gpurun_out
for v in 3 2; do
  HNB_SSD_FWD=$v timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu --no-parity > gpurun_out/r2d_bench_fwd$v.json 2> gpurun_out/r2d_bench_fwd$v.err
  python - <<PY
import json
j=json.load(open("gpurun_out/r2d_bench_fwd$v.json"))
print("HNB_SSD_FWD=$v", round(j["value"]), round(j["ms_per_step"],3), "hot", round(j["hot_path"]["ms_per_step"],3))
for r in j["kernel_table"][:4]: print("   ", r["kernel"], r["launches"], r["ms"], r.get("frac"))
PY
done
