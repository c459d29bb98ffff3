#!/bin/bash
# 2 GPUs: slab parity with the totals reduction on the side stream + weak bench line
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T="timeout -k 5"
$T 600 python -m pytest tests/test_gpu_multi.py -x -q -k "slabs_match or campaign" > gpurun_out/m2b_tests.log 2>&1
tail -n 3 gpurun_out/m2b_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
$T 300 $TR bench.py --gpus 2 --no-cpu-baseline --no-e2e --no-api-e2e > gpurun_out/m2b_weak.json 2> gpurun_out/m2b_weak.err
$T 300 $TR bench.py --gpus 2 --scaling strong --no-cpu-baseline --no-e2e --no-api-e2e > gpurun_out/m2b_strong.json 2> gpurun_out/m2b_strong.err
for f in weak strong; do python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/m2b_$f.json").read().strip().splitlines()[-1])
    print("$f: value", d["value"], "ms/step", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items() if v})
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/m2b_$f.err").read()[-1500:])
PY
done
