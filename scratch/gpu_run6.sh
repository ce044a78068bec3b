#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mamba.py -q -rf -x -k "ssd_tcgen05_forward or repeatable" > gpurun_out/r2c_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_tests.log
grep -n "passed\|failed\|FAILED\|^E  \|rc=\|timed out" gpurun_out/r2c_tests.log | head -30
timeout 300 python scratch/ssd_time.py 2>&1 | tee gpurun_out/r2c_ssd_time.log
python scratch/ssd_fwd_prof.py 1 > gpurun_out/r2c_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c_fwd_launches.csv python scratch/ssd_fwd_prof.py 1 > gpurun_out/r2c_ncu.log 2>&1
grep -o '"ssd_[a-z_]*\|"void.*ssd_[a-z_]*\|gpu__time_duration.sum","[a-z]*","[0-9.,]*"\|issue_active[^"]*","%","[0-9.]*"\|inst_executed.sum","inst","[0-9,]*"' gpurun_out/r2c_fwd_launches.csv | paste - - - - | head -12
