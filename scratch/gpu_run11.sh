#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --mode decode --steps 5 --warmup 3 --no-cpu > gpurun_out/r2e_decode.json 2> gpurun_out/r2e_decode.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r2e_decode.json"))
print("decode", round(j["value"]), round(j["ms_per_step"],3), j["dtype"])
for r in j["kernel_table"][:10]: print("   ", r["kernel"], r["launches"], r["ms"], r.get("frac"), r.get("bound"))
print(j.get("parity"))
PY
tail -3 gpurun_out/r2e_decode.err
for n in 1 2 3 4; do
  timeout 600 python bench.py --workload ragged --N $n --steps 12 --warmup 3 --no-cpu --no-parity > gpurun_out/r2e_ragged_N$n.json 2> gpurun_out/r2e_ragged_N$n.err
  python - <<PY
import json
j=json.load(open("gpurun_out/r2e_ragged_N$n.json"))
print("ragged N=$n", round(j["value"]), round(j["ms_per_step"],3), "kept", j["config"]["kept_fraction"], "pad", j["config"]["ragged"]["padding_share"], "e2e", round(j["e2e"]["value"]))
PY
done
