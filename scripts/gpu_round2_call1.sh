set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/r1_smi.txt
timeout 900 python -m pytest tests/test_gpu_sweep.py -q --timeout 180 > gpurun_out/r1_sweep_tests.log 2>&1
echo "sweep tests rc=$?" >> gpurun_out/r1_sweep_tests.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1_bench_fused.json 2> gpurun_out/r1_bench_fused.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-tma > gpurun_out/r1_bench_notma.json 2> gpurun_out/r1_bench_notma.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --gather > gpurun_out/r1_bench_gather.json 2> gpurun_out/r1_bench_gather.err
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_gpu_sweep.py > gpurun_out/r1_all_tests.log 2>&1
echo "all tests rc=$?" >> gpurun_out/r1_all_tests.log
