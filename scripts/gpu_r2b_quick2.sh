#!/bin/bash
# refresh chunk size A/B: 512 (default), 256 (flag 1048576), 1024 (flag 2097152)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T="timeout -k 5"
B="python bench.py --no-cpu-baseline --no-e2e --no-api-e2e"
$T 600 python -m pytest tests/test_gpu_sweep.py -x -q -k "refresh_variants or resident_rates" > gpurun_out/q_pytest.log 2>&1
tail -n 2 gpurun_out/q_pytest.log
for f in 0 1048576 2097152; do
$T 200 $B --debug-flags $f > gpurun_out/q_bench_$f.json 2> gpurun_out/q_bench_$f.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/q_bench_$f.json").read().strip().splitlines()[-1])
    print("flags $f ms/step", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items() if v})
except Exception as e:
    print("failed", e); print(open("gpurun_out/q_bench_$f.err").read()[-2000:])
PY
done
