#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_mamba.py -q -x -k "gemm or umma" 2>&1 | tail -2
timeout 300 python scratch/kbench.py 2>/dev/null | grep "gemm"
