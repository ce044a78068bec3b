#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
HNB_SUB_FWD_CPT=8 HNB_SUB_BWD_BLOCKS_PER_SM=6 timeout 400 python bench.py --steps 10 --warmup 5 --no-cpu > gpurun_out/r2z_old_$i.json 2>/dev/null
timeout 400 python bench.py --steps 10 --warmup 5 --no-cpu > gpurun_out/r2z_new_$i.json 2>/dev/null
done
python - <<'PY'
import json
for f in ("old_1", "new_1", "old_2", "new_2"):
    d = json.loads(open(f"gpurun_out/r2z_{f}.json").read().strip().splitlines()[-1])
    kt = {r["kernel"]: r["ms"] for r in d["kernel_table"]}
    print(f, round(d["value"]), round(d["ms_per_step"], 3), round(d["hot_path"]["ms_per_step"], 3), kt.get("subsample_conv1_fwd"), kt.get("subsample_conv1_bwd"), d["clocks"])
PY
