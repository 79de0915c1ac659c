cd $GRAFT_REPO_ROOT
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02p_bench_E_n$N.json 2> gpurun_out/r02p_bench_E_n$N.err; tail -2 gpurun_out/r02p_bench_E_n$N.err; cut -c1-400 gpurun_out/r02p_bench_E_n$N.json
