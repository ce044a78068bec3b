#!/bin/bash
mkdir -p gpurun_out
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus 2 --steps 10 --warmup 5 --no-cpu --no-parity > gpurun_out/r2n_$tag.json 2> gpurun_out/r2n_$tag.err
  python - <<PY
import json
try:
    txt=open("gpurun_out/r2n_$tag.json").read(); j=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print("$tag", round(j["value"]), round(j["ms_per_step"],3), "exposed", j.get("allreduce_exposed_ms"))
except Exception as e:
    print("$tag FAILED", e); print(open("gpurun_out/r2n_$tag.err").read()[-800:])
PY
}
timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu --no-parity > gpurun_out/r2n_n1.json 2>/dev/null
python -c "import json; j=json.load(open('gpurun_out/r2n_n1.json')); print('N=1', round(j['value']), round(j['ms_per_step'],3))"
run default A=1
run maxctas8 NCCL_MAX_CTAS=8
run maxctas4 NCCL_MAX_CTAS=4
run maxctas2 NCCL_MAX_CTAS=2
run nooverlap HNB_REDUCER_OVERLAP=0
run bucket128 HNB_BUCKET_MB=128
run bucket8 HNB_BUCKET_MB=8
