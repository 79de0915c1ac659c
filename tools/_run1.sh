set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/r2_gputest1.log
timeout 300 python tools/kernel_times.py 1000000 128 10 1 > gpurun_out/r2_kt_D.log 2>&1
timeout 300 python tools/kernel_times.py 10000 64 40 1 > gpurun_out/r2_kt_A.log 2>&1
timeout 300 python tools/kernel_times.py 10000 64 40 64 > gpurun_out/r2_kt_E64.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -5 gpurun_out/r2_gputest1.log; cat gpurun_out/r2_kt_*.log; cat gpurun_out/r2_bench1.json
