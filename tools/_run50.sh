cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1200 python bench.py > gpurun_out/r02h_bench_default.json 2> gpurun_out/r02h_bench_default.err ) 2>&1 | grep real
tail -2 gpurun_out/r02h_bench_default.err; python -c "
import json; d=json.loads(open('gpurun_out/r02h_bench_default.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['path']['frac'], d.get('cpu_baseline'))"
( time timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02h_bench_reference_arm.json 2> gpurun_out/r02h_bench_ref.err ) 2>&1 | grep real
tail -2 gpurun_out/r02h_bench_ref.err; cut -c1-700 gpurun_out/r02h_bench_reference_arm.json
