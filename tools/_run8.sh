cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r2_gputest8.log
tail -30 gpurun_out/r2_gputest8.log
timeout 900 python bench.py --steps 5 > gpurun_out/r2_bench_D.json 2> gpurun_out/r2_bench_D.err; tail -5 gpurun_out/r2_bench_D.err; cat gpurun_out/r2_bench_D.json
timeout 300 python tools/kernel_times.py 1000000 128 10 1 2>&1 | tee gpurun_out/r2_kt_D.log
