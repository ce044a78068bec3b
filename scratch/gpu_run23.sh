#!/bin/bash
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_mamba.py -q -x -k "subsample" 2>&1 | grep -E "assert|Error|passed|failed" | head -8; done
echo "--- PDL off"; for i in 1 2; do HNB_PDL=0 timeout 300 python -m pytest tests/test_gpu_mamba.py -q -x -k "subsample" 2>&1 | grep -E "assert|Error|passed|failed" | head -8; done
for n in 148 111 74; do echo "CTAS=$n"; HNB_SSD_SPAN_CTAS=$n timeout 300 python scratch/ssd_time.py 2>&1 | tail -4 | sed 's/fwd impl 5 [0-9.]*, fwd impl 4 [0-9.]*, //; s/, bwd impl 3.*//'; done
timeout 300 python scratch/ssd_bwd_dbg.py 2>&1 | grep -i "fused" | head -12
