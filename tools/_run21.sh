cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02_gputest_s4.log
timeout 300 python tools/kernel_times.py 1000000 128 10 1 2>&1 | tail -2 | tee gpurun_out/r02_kt_s4.log
timeout 300 python tools/kernel_times.py 10000 64 40 64 2>&1 | tail -1 | tee -a gpurun_out/r02_kt_s4.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02c_bench_D.json 2> gpurun_out/r02c_bench_D.err; tail -3 gpurun_out/r02c_bench_D.err; cat gpurun_out/r02c_bench_D.json
timeout 900 python bench.py --workload E --steps 2 --warmup 3 > gpurun_out/r02c_bench_E_n1.json 2> gpurun_out/r02c_bench_E_n1.err; tail -3 gpurun_out/r02c_bench_E_n1.err; cat gpurun_out/r02c_bench_E_n1.json
