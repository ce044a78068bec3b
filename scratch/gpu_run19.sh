#!/bin/bash
# stack-level weight packing: tests of the block / stack / encoder paths, then the headline bench with and without it
timeout 900 python -m pytest tests/test_gpu_mamba.py tests/test_gpu_hnet.py -q -x -k "block or stack or encoder or composite" 2>&1 | tail -3
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2p_bench_A.json 2> gpurun_out/r2p_bench_A.err; tail -c 600 gpurun_out/r2p_bench_A.json | head -c 300
HNB_STACK_PACK=0 timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2p_bench_A_nopack.json 2> gpurun_out/r2p_bench_A_nopack.err
python - <<'PY'
import json
for f in ("r2p_bench_A", "r2p_bench_A_nopack"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("hot_path"))
    except Exception as e:
        print(f, "ERR", e)
PY
