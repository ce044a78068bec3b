#!/bin/bash
# round 2, first GPU call: the whole GPU test suite, smoke, and the bench lines of every BASELINE workload
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
nproc >> gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -rf -s --durations=15 > gpurun_out/r2_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 10 --warmup 5 > gpurun_out/r2_bench_A.json 2> gpurun_out/r2_bench_A.err
timeout 600 python bench.py --workload B_small_N4 --steps 10 --warmup 5 --no-cpu > gpurun_out/r2_bench_B.json 2> gpurun_out/r2_bench_B.err
timeout 600 python bench.py --workload A_large_N3_60s --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_L.json 2> gpurun_out/r2_bench_L.err
timeout 600 python bench.py --workload ragged --steps 12 --warmup 3 --no-cpu > gpurun_out/r2_bench_R.json 2> gpurun_out/r2_bench_R.err
tail -n 25 gpurun_out/r2_tests.log
cat gpurun_out/r2_smoke.log | tail -n 5
for f in A B L R; do echo "== bench $f"; head -c 600 gpurun_out/r2_bench_$f.json; echo; tail -n 3 gpurun_out/r2_bench_$f.err; done
