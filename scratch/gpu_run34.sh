#!/bin/bash
# session-5 evidence: smoke, bench lines of every workload, launch list + DRAM traffic of two headline steps, full capture of one block + front end
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 5 > gpurun_out/s5_bench_A.json 2> gpurun_out/s5_bench_A.err
timeout 600 python bench.py --workload B_small_N4 --steps 10 --warmup 5 --no-cpu > gpurun_out/s5_bench_B.json 2> gpurun_out/s5_bench_B.err
timeout 600 python bench.py --workload A_large_N3_60s --steps 5 --warmup 3 --no-cpu > gpurun_out/s5_bench_L.json 2> gpurun_out/s5_bench_L.err
timeout 600 python bench.py --workload ragged --steps 12 --warmup 3 --no-cpu > gpurun_out/s5_bench_R.json 2> gpurun_out/s5_bench_R.err
timeout 600 python bench.py --mode decode --steps 5 --warmup 3 --no-cpu > gpurun_out/s5_bench_D.json 2> gpurun_out/s5_bench_D.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s5_bench_ref.json 2> gpurun_out/s5_bench_ref.err
for f in A B L R D ref; do echo "== $f"; head -c 260 gpurun_out/s5_bench_$f.json; echo; tail -n 1 gpurun_out/s5_bench_$f.err; done
timeout 300 python scratch/ncu_step.py > gpurun_out/s5_step_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/s5_launches.csv python scratch/ncu_step.py > gpurun_out/s5_ncu.log 2>&1
tail -2 gpurun_out/s5_ncu.log
timeout 300 python scratch/ncu_target.py > gpurun_out/s5_target_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ssd_|conv_|gemm_bf16|norm_|pack_mixer" -f -o gpurun_out/s5_block_full python scratch/ncu_target.py > gpurun_out/s5_ncu_full.log 2>&1
tail -2 gpurun_out/s5_ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"sub_conv1|bias_relu" -f -o gpurun_out/s5_sub_full python scratch/ncu_subsample.py > gpurun_out/s5_ncu_sub.log 2>&1
tail -1 gpurun_out/s5_ncu_sub.log
ls -la gpurun_out/*.ncu-rep
