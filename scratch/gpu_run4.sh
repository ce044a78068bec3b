#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mamba.py -q -rf -k "ssd_tcgen05" -x > gpurun_out/r2_tests3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests3.log
grep -n "passed\|failed\|FAILED\|^E  \|rc=" gpurun_out/r2_tests3.log | head -20
timeout 300 python scratch/ssd_time.py 2>&1 | tee gpurun_out/r2_ssd_time.log
timeout 300 python scratch/ssd_dbg2.py 2>&1 | tee gpurun_out/r2_ssd_dbg.log
