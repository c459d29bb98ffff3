#!/bin/bash
# 2 GPUs: slab parity tests + bench lines (weak: 512 planes per GPU; strong: the 512^3 lattice over 2 slabs)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T="timeout -k 5"
$T 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r02_multi_gpu_tests_2gpu.log 2>&1
tail -n 3 gpurun_out/r02_multi_gpu_tests_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
$T 300 $TR bench.py --gpus 2 --no-cpu-baseline --no-api-e2e > gpurun_out/r02_bench_weak512_2gpu.json 2> gpurun_out/r02_bench_weak512_2gpu.err
$T 300 $TR bench.py --gpus 2 --scaling strong --no-cpu-baseline --no-e2e --no-api-e2e > gpurun_out/r02_bench_strong512_2gpu.json 2> gpurun_out/r02_bench_strong512_2gpu.err
for f in weak512 strong512; do tail -c 600 gpurun_out/r02_bench_${f}_2gpu.json | head -c 300; echo; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_${f}_2gpu.json").read().strip().splitlines()[-1])
    print("$f: value", d["value"], "ms/step", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items() if v})
except Exception as e:
    print("$f failed", e)
PY
done
