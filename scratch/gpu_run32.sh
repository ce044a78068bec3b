#!/bin/bash
mkdir -p gpurun_out
for g in 16 8 4 2 1; do echo "heads per group $g"; HNB_SSD_EXACT_HPG=$g timeout 200 python scratch/ssd_exact_time.py; done
timeout 600 python -m pytest tests/test_gpu_mamba.py -q -x -k "ssd or golden or block_at" 2>&1 | tail -2
