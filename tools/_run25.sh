cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02d_bench_E_n2.json 2> gpurun_out/r02d_bench_E_n2.err; tail -3 gpurun_out/r02d_bench_E_n2.err; cat gpurun_out/r02d_bench_E_n2.json
nproc; free -g | head -2
