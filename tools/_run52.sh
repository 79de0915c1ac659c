cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02h_bench_E_n$N.json 2> gpurun_out/r02h_bench_E_n$N.err; tail -2 gpurun_out/r02h_bench_E_n$N.err; cut -c1-300 gpurun_out/r02h_bench_E_n$N.json
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02h_bench_ref_n$N.json 2> gpurun_out/r02h_bench_ref_n$N.err ) 2>&1 | grep real; tail -2 gpurun_out/r02h_bench_ref_n$N.err; cut -c1-600 gpurun_out/r02h_bench_ref_n$N.json
