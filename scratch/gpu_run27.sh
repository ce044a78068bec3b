#!/bin/bash
for n in 3 2 4; do echo "dstate per SM = $n"; HNB_SSD_DSTATE_PER_SM=$n timeout 300 python scratch/ssd_time.py 2>&1 | tail -4 | sed 's/fwd impl 5 [0-9.]*, fwd impl 4 [0-9.]*, //; s/, bwd impl 3.*//'; done
timeout 600 python -m pytest tests/test_gpu_mamba.py -q -x -k "ssd" 2>&1 | tail -2
