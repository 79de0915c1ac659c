cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02f_gputest.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02f_bench_D.json 2> gpurun_out/r02f_bench_D.err; tail -3 gpurun_out/r02f_bench_D.err; cat gpurun_out/r02f_bench_D.json
timeout 900 python bench.py --workload E --steps 2 --warmup 3 > gpurun_out/r02f_bench_E_n1.json 2> gpurun_out/r02f_bench_E_n1.err; tail -3 gpurun_out/r02f_bench_E_n1.err; cat gpurun_out/r02f_bench_E_n1.json
for w in A B C; do timeout 600 python bench.py --workload $w --steps 5 --no-cpu-baseline > gpurun_out/r02f_bench_$w.json 2> gpurun_out/r02f_bench_$w.err; tail -3 gpurun_out/r02f_bench_$w.err; cat gpurun_out/r02f_bench_$w.json | cut -c1-400; done
