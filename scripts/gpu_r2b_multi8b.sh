#!/bin/bash
# 8 GPUs, final build (totals reduction on the side stream): weak line
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T="timeout -k 5"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
$T 300 $TR bench.py --gpus 8 --no-cpu-baseline --no-e2e --no-api-e2e > gpurun_out/r02_bench_weak512_8gpu.json 2> gpurun_out/r02_bench_weak512_8gpu.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_weak512_8gpu.json").read().strip().splitlines()[-1])
    print("weak512: value", d["value"], "ms/step", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items() if v})
except Exception as e:
    print("failed", e); print(open("gpurun_out/r02_bench_weak512_8gpu.err").read()[-2000:])
PY
