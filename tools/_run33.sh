cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02e_bench_D.json 2> gpurun_out/r02e_bench_D.err; tail -3 gpurun_out/r02e_bench_D.err; cat gpurun_out/r02e_bench_D.json
