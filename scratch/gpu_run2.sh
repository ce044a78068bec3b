#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py "tests/test_gpu_mamba.py::test_gemm_bf16_cta_pair_path" tests/test_gpu_reference_properties.py -q -rf -s > gpurun_out/r2_tests2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests2.log
grep -n "ours-vs\|forced\|rel err\|passed\|failed\|FAILED\|^E " gpurun_out/r2_tests2.log | head -60
