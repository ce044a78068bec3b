#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mamba.py -q -x -k "ssd or golden or block_at or composite" 2>&1 | tail -3
timeout 200 python scratch/ssd_exact_time.py
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/ssd_exact.csv python scratch/ssd_exact_time.py prof > /dev/null 2>&1
python -c "
import csv
rows=[r for r in csv.reader(open('gpurun_out/ssd_exact.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    if 'time' in r[h.index('Metric Name')]: print(r[h.index('Kernel Name')][:40], r[h.index('Metric Value')], r[h.index('Metric Unit')])
"
timeout 400 python bench.py --mode decode --steps 5 --warmup 3 --no-cpu > gpurun_out/r2w_bench_D.json 2> gpurun_out/r2w_bench_D.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2w_bench_D.json").read().strip().splitlines()[-1])
kt = {r["kernel"]: r["ms"] for r in d["kernel_table"]}
print("decode", round(d["value"]), d["ms_per_step"], kt.get("ssd_fwd"), kt.get("gemm_f32_tc"), d["parity"])
PY
