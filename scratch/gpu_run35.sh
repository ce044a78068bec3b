#!/bin/bash
# 2-GPU sanity at the final code: headline workload under torchrun (NCCL gradient all-reduce inside the step) + the 2-GPU tests
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu --no-parity > gpurun_out/s5_A_n1.json 2> gpurun_out/s5_A_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 \
    bench.py --gpus 2 --steps 10 --warmup 5 --no-cpu --no-parity > gpurun_out/s5_A_n2.json 2> gpurun_out/s5_A_n2.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        j = json.loads(open(f"gpurun_out/s5_A_n{n}.json").read().strip().splitlines()[-1])
        print(f"A N={n}", round(j["value"]), "frames/s", round(j["ms_per_step"], 3), "ms  exposed allreduce", j.get("allreduce_exposed_ms"), "e2e", round(j["e2e"]["value"]))
    except Exception as e:
        print(f"A N={n} FAILED", e); print(open(f"gpurun_out/s5_A_n{n}.err").read()[-1500:])
PY
timeout 600 python -m pytest tests -m gpu -q -x -k "two_gpu or 2gpu or non_current or multi" 2>&1 | tail -2
