#!/bin/bash
# quick A/B: sweep tests + one bench line
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
T="timeout -k 5"
B="python bench.py --no-cpu-baseline --no-e2e --no-api-e2e"
$T 600 python -m pytest tests/test_gpu_sweep.py -x -q -k "refresh_variants or resident_rates or dense_rebuild or deterministic or primed" > gpurun_out/q_pytest.log 2>&1
tail -n 2 gpurun_out/q_pytest.log
$T 200 $B > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/q_bench.json").read().strip().splitlines()[-1])
    print("ms/step", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["kernel_ms_per_step"].items() if v}, "events", d["executed_events_per_s"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/q_bench.err").read()[-2000:])
PY
