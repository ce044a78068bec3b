#!/bin/bash
# 8-GPU box: headline at N=8, Type B Small N=4 at N=2/4/8 (BASELINE config 3), the trainer-sync micro-benchmark at N=2
mkdir -p gpurun_out
run() { # n workload tag
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + $1)) \
    bench.py --gpus $1 --workload $2 --steps 10 --warmup 5 --no-cpu --no-parity > gpurun_out/r2m_$3_n$1.json 2> gpurun_out/r2m_$3_n$1.err
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2m_$3_n$1.json"))
    print("$2 N=$1", round(j["value"]), "frames/s", round(j["ms_per_step"],3), "ms  exposed allreduce", j.get("allreduce_exposed_ms"), "e2e", round(j["e2e"]["value"]))
except Exception as e:
    print("$2 N=$1 FAILED", e); print(open("gpurun_out/r2m_$3_n$1.err").read()[-1500:])
PY
}
timeout 600 python bench.py --workload B_small_N4 --steps 10 --warmup 5 --no-cpu --no-parity > gpurun_out/r2m_B_n1.json 2> gpurun_out/r2m_B_n1.err
python -c "import json; j=json.load(open('gpurun_out/r2m_B_n1.json')); print('B_small_N4 N=1', round(j['value']), round(j['ms_per_step'],3))"
run 2 B_small_N4 B
run 4 B_small_N4 B
run 8 B_small_N4 B
timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu --no-parity > gpurun_out/r2m_A_n1.json 2> gpurun_out/r2m_A_n1.err
python -c "import json; j=json.load(open('gpurun_out/r2m_A_n1.json')); print('A_small_N2 N=1', round(j['value']), round(j['ms_per_step'],3))"
run 2 A_small_N2 A
run 8 A_small_N2 A
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 scratch/oom_sync_bench.py 2>&1 | tail -4
