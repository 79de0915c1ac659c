cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q 2>&1 | tail -5
for w in A B C; do timeout 600 python bench.py --workload $w --steps 5 --no-cpu-baseline > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; tail -3 gpurun_out/r2_bench_$w.err; cat gpurun_out/r2_bench_$w.json; done
DD_BENCH_SUBBATCH=64 timeout 900 python bench.py --workload E --steps 1 --warmup 3 > gpurun_out/r2_bench_E64.json 2> gpurun_out/r2_bench_E64.err; tail -5 gpurun_out/r2_bench_E64.err; cat gpurun_out/r2_bench_E64.json
timeout 900 python bench.py --workload E --steps 1 --warmup 3 > gpurun_out/r2_bench_E128.json 2> gpurun_out/r2_bench_E128.err; tail -5 gpurun_out/r2_bench_E128.err; cat gpurun_out/r2_bench_E128.json
timeout 900 python bench.py --steps 5 > gpurun_out/r2_bench_D.json 2> gpurun_out/r2_bench_D.err; tail -5 gpurun_out/r2_bench_D.err; cat gpurun_out/r2_bench_D.json
