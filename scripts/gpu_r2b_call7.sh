#!/bin/bash
# call 7: dense kernel with counting-sorted pass B vs unsorted (flag 131072)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T="timeout -k 5"
B="python bench.py --no-cpu-baseline --no-e2e --no-api-e2e"
$T 600 python -m pytest tests/test_gpu_sweep.py -x -q -k "refresh_variants or resident_rates or dense_rebuild or deterministic or primed" > gpurun_out/c7_pytest.log 2>&1
tail -n 3 gpurun_out/c7_pytest.log
for f in 0 131072; do
  $T 200 $B --debug-flags $f > gpurun_out/c7_bench_flags$f.json 2> gpurun_out/c7_bench_flags$f.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/c7_bench_flags$f.json").read().strip().splitlines()[-1])
    print("flags $f: ms/step", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items()})
except Exception as e:
    print("flags $f: failed", e)
PY
done
