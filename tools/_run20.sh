cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02_gputest_s3.log
timeout 600 python tools/e2e_gradmodel_parts.py D 2>&1 | tail -4 | tee gpurun_out/r02_e2e_parts_D.log
timeout 900 python bench.py --workload E --steps 2 --warmup 3 > gpurun_out/r02b_bench_E_n1.json 2> gpurun_out/r02b_bench_E_n1.err; tail -3 gpurun_out/r02b_bench_E_n1.err; cat gpurun_out/r02b_bench_E_n1.json
DD_BENCH_SUBBATCH=64 timeout 900 python bench.py --workload E --steps 2 --warmup 3 > gpurun_out/r02b_bench_E_n1_sub64.json 2> gpurun_out/r02b_bench_E_n1.err; tail -3 gpurun_out/r02b_bench_E_n1.err; cat gpurun_out/r02b_bench_E_n1_sub64.json
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02b_bench_D.json 2> gpurun_out/r02b_bench_D.err; tail -3 gpurun_out/r02b_bench_D.err; cat gpurun_out/r02b_bench_D.json
