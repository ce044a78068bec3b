#!/bin/bash
mkdir -p gpurun_out
python scratch/ssd_fwd_prof.py > gpurun_out/r2c_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c_fwd_launches.csv python scratch/ssd_fwd_prof.py > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
