cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
P=29811
timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 > gpurun_out/r10_multi_tests_$N.log 2>&1
echo "rc=$?" >> gpurun_out/r10_multi_tests_$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N > gpurun_out/r10_strong512_$N.json 2> gpurun_out/r10_strong512_$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus $N --L 1024 > gpurun_out/r10_strong1024_$N.json 2> gpurun_out/r10_strong1024_$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --gpus $N --scaling weak > gpurun_out/r10_weak512_$N.json 2> gpurun_out/r10_weak512_$N.err
