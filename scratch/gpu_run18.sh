#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_mamba.py -q -x -k "gemm or umma or ssd_tcgen05" 2>&1 | tail -2
timeout 300 python scratch/ssd_time.py 2>&1 | head -4
timeout 300 python scratch/kbench.py 2>/dev/null | grep "gemm" | head -6
