#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/ncu_subsample.py > gpurun_out/r2v_sub_plain.log 2>&1 && tail -1 gpurun_out/r2v_sub_plain.log &&
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"sub_conv1|bias_relu" -f -o gpurun_out/r2v_sub_full python scratch/ncu_subsample.py > gpurun_out/r2v_ncu.log 2>&1
tail -2 gpurun_out/r2v_ncu.log
python scratch/ncu_summary.py gpurun_out/r2v_sub_full.ncu-rep gpurun_out/r2v_sub_full.md | tail -8
