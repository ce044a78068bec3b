#!/bin/bash
# 8-GPU box, final code: headline workload at N = 8 (and N = 4) under torchrun
mkdir -p gpurun_out
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
    bench.py --gpus $n --steps 10 --warmup 5 --no-cpu --no-parity > gpurun_out/s5_A_n$n.json 2> gpurun_out/s5_A_n$n.err
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/s5_A_n$n.json").read().strip().splitlines()[-1])
    print("A N=$n", round(j["value"]), "frames/s", round(j["ms_per_step"], 3), "ms  exposed allreduce", j.get("allreduce_exposed_ms"), "e2e", round(j["e2e"]["value"]))
except Exception as e:
    print("A N=$n FAILED", e); print(open("gpurun_out/s5_A_n$n.err").read()[-1500:])
PY
done
