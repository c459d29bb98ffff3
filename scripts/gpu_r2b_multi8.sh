#!/bin/bash
# 8 GPUs: bench lines — weak (512 planes per GPU), strong 1024^3 (BASELINE configs[3]), strong 512^3 (the metric's lattice)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T="timeout -k 5"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
$T 300 $TR bench.py --gpus 8 --no-cpu-baseline --no-e2e --no-api-e2e > gpurun_out/r02_bench_weak512_8gpu.json 2> gpurun_out/r02_bench_weak512_8gpu.err
$T 300 $TR bench.py --gpus 8 --scaling strong --L 1024 --no-cpu-baseline --no-e2e --no-api-e2e > gpurun_out/r02_bench_strong1024_8gpu.json 2> gpurun_out/r02_bench_strong1024_8gpu.err
$T 300 $TR bench.py --gpus 8 --scaling strong --no-cpu-baseline --no-e2e --no-api-e2e > gpurun_out/r02_bench_strong512_8gpu.json 2> gpurun_out/r02_bench_strong512_8gpu.err
for f in weak512 strong1024 strong512; do python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_${f}_8gpu.json").read().strip().splitlines()[-1])
    print("$f: value", d["value"], "ms/step", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items() if v})
except Exception as e:
    print("$f failed", e)
PY
done
