cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/r02j_bench_default.json 2> gpurun_out/r02j_bench_default.err; tail -2 gpurun_out/r02j_bench_default.err; python -c "
import json; d=json.loads(open('gpurun_out/r02j_bench_default.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['path']['frac'], d['roofline']['kernels_us'])"
