#!/bin/bash
# round-2 final evidence: bench lines of the BASELINE workloads, launch list + DRAM traffic of two headline steps, full capture of one block
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 5 > gpurun_out/r2f_bench_A.json 2> gpurun_out/r2f_bench_A.err
timeout 600 python bench.py --workload B_small_N4 --steps 10 --warmup 5 --no-cpu > gpurun_out/r2f_bench_B.json 2> gpurun_out/r2f_bench_B.err
timeout 600 python bench.py --workload A_large_N3_60s --steps 5 --warmup 3 --no-cpu > gpurun_out/r2f_bench_L.json 2> gpurun_out/r2f_bench_L.err
timeout 600 python bench.py --workload ragged --steps 12 --warmup 3 --no-cpu > gpurun_out/r2f_bench_R.json 2> gpurun_out/r2f_bench_R.err
timeout 600 python bench.py --mode decode --steps 5 --warmup 3 --no-cpu > gpurun_out/r2f_bench_D.json 2> gpurun_out/r2f_bench_D.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
for f in A B L R D ref; do echo "== $f"; head -c 260 gpurun_out/r2f_bench_$f.json; echo; tail -n 2 gpurun_out/r2f_bench_$f.err; done
timeout 300 python scratch/ncu_step.py > gpurun_out/r2f_step_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f_launches.csv python scratch/ncu_step.py > gpurun_out/r2f_ncu.log 2>&1
tail -2 gpurun_out/r2f_ncu.log
timeout 300 python scratch/ncu_target.py > gpurun_out/r2f_target_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ssd_|conv_|gemm_bf16|norm_|pack_mixer" -f -o gpurun_out/r2f_block_full python scratch/ncu_target.py > gpurun_out/r2f_ncu_full.log 2>&1
tail -2 gpurun_out/r2f_ncu_full.log
ls -la gpurun_out/r2f_block_full.ncu-rep
