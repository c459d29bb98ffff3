cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
rm -rf outputs
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29841 scripts/run_gr_sweep.py --L 96 --sweeps 401 > gpurun_out/r12_gr_sweep_$N.log 2> gpurun_out/r12_gr_sweep_$N.err
cp outputs/gr_sweep/cet_map.csv gpurun_out/r12_cet_map.csv 2>/dev/null
ls -R outputs/gr_sweep/plot_cet | head -40 > gpurun_out/r12_tree.txt
