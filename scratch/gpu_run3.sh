#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mamba.py -q -rf -k "ssd_tcgen05" > gpurun_out/r2_tests3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests3.log
grep -n "passed\|failed\|FAILED\|^E  " gpurun_out/r2_tests3.log | head -40
timeout 300 python scratch/ssd_time.py 2>&1 | tee gpurun_out/r2_ssd_time.log
