#!/bin/bash
mkdir -p gpurun_out
python scratch/ssd_fwd_prof.py 1 > gpurun_out/r2c_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ssd_states|ssd_scan" -c 2 -f -o gpurun_out/r2c_split python scratch/ssd_fwd_prof.py 1 > gpurun_out/r2c_ncu2.log 2>&1
tail -3 gpurun_out/r2c_ncu2.log
for h in 1 2 4; do echo "state heads $h"; HNB_SSD_STATE_HEADS=$h python scratch/ssd_time.py 2>&1 | head -2; done
for h in 3 6; do echo "scan heads $h"; HNB_SSD_SCAN_HEADS=$h python scratch/ssd_time.py 2>&1 | head -2; done
