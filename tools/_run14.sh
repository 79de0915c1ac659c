cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_D.json 2> gpurun_out/r02_bench_D.err; tail -3 gpurun_out/r02_bench_D.err; cat gpurun_out/r02_bench_D.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; tail -3 gpurun_out/r02_bench_ref.err; cat gpurun_out/r02_bench_reference_arm.json
for w in A B C; do timeout 600 python bench.py --workload $w --steps 5 --no-cpu-baseline > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; tail -3 gpurun_out/r02_bench_$w.err; cat gpurun_out/r02_bench_$w.json; done
timeout 900 python bench.py --workload E --steps 2 --warmup 3 > gpurun_out/r02_bench_E_n1.json 2> gpurun_out/r02_bench_E_n1.err; tail -3 gpurun_out/r02_bench_E_n1.err; cat gpurun_out/r02_bench_E_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_p2g_tile|k_g2p_tile|k_grid_b" -s 15 -c 3 -o gpurun_out/r02_prof_fwd -f python tools/kernel_times.py 1000000 128 4 1 > gpurun_out/r02_prof_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_g2p_grad_tile|k_p2g_grad_tile|k_grid_grad_b" -s 12 -c 3 -o gpurun_out/r02_prof_bwd -f python tools/kernel_times.py 1000000 128 4 1 > gpurun_out/r02_prof_bwd.log 2>&1
ls -la gpurun_out | tail -20
