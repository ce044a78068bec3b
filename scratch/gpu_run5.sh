#!/bin/bash
# round 2, re-entry: whole GPU suite at HEAD, smoke, headline bench, launch list of two steps
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf --durations=10 > gpurun_out/r2b_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2b_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2b_smoke.log
timeout 600 python bench.py --steps 10 --warmup 5 > gpurun_out/r2b_bench_A.json 2> gpurun_out/r2b_bench_A.err
timeout 300 python scratch/ncu_step.py > gpurun_out/r2b_step_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2b_launches.csv python scratch/ncu_step.py > gpurun_out/r2b_ncu.log 2>&1
tail -n 14 gpurun_out/r2b_tests.log
tail -n 4 gpurun_out/r2b_smoke.log
head -c 1500 gpurun_out/r2b_bench_A.json; echo; tail -n 3 gpurun_out/r2b_bench_A.err
tail -n 3 gpurun_out/r2b_ncu.log
