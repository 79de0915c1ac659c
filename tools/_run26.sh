cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02d_bench_E_n$N.json 2> gpurun_out/r02d_bench_E_n$N.err; tail -3 gpurun_out/r02d_bench_E_n$N.err; cat gpurun_out/r02d_bench_E_n$N.json
