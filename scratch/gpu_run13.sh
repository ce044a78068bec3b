#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mamba.py -q -rf -x -k "gemm" 2>&1 | tail -3
echo "== A panel on"; timeout 300 python scratch/kbench.py 2>/dev/null | grep "gemm"
echo "== A panel off"; HNB_GEMM_APANEL=0 timeout 300 python scratch/kbench.py 2>/dev/null | grep "gemm"
