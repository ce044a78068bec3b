#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mamba.py -q -x -k "ssd_tcgen05_forward" 2>&1 | tail -2
python scratch/ssd_fwd_prof.py 1 > gpurun_out/r2c_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c_fwd_launches.csv python scratch/ssd_fwd_prof.py 1 > gpurun_out/r2c_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2c_fwd_launches.csv")) if len(r)>5]
h=next(r for r in rows if "Kernel Name" in r); i0=rows.index(h)
kn,mn,mv,idc=h.index("Kernel Name"),h.index("Metric Name"),h.index("Metric Value"),h.index("ID")
cur={}
for r in rows[i0+1:]: cur.setdefault((int(r[idc]), r[kn].split("::")[-1][:24]),{})[r[mn].split("__")[-1][:14]]=r[mv]
for k,m in sorted(cur.items()): print(k, m)
PY
