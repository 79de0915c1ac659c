cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02i_e2e_launches_D.csv python tools/e2e_step_once.py D 2>&1 | tail -1
python tools/launch_summary.py gpurun_out/r02i_e2e_launches_D.csv 12
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02i_bench_D.json 2> gpurun_out/r02i_bench_D.err; tail -2 gpurun_out/r02i_bench_D.err; python -c "
import json; d=json.loads(open('gpurun_out/r02i_bench_D.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['path']['frac'])"
timeout 900 python bench.py --workload E --steps 2 --warmup 3 > gpurun_out/r02i_bench_E_n1.json 2> gpurun_out/r02i_bench_E_n1.err; tail -2 gpurun_out/r02i_bench_E_n1.err; cut -c1-160 gpurun_out/r02i_bench_E_n1.json
