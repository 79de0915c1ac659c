cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider 2>&1 | grep -E "^/|^E |^FAILED|passed|failed" | cut -c1-300 | head -8
timeout 600 python bench.py --workload E --no-cpu-baseline > $O/r02v_bench_E.json 2> $O/r02v_bench_E.err
timeout 600 python bench.py --no-cpu-baseline > $O/r02v_bench_default.json 2> $O/r02v_bench_default.err
timeout 600 python bench.py --workload C --no-cpu-baseline > $O/r02v_bench_C.json 2> $O/r02v_bench_C.err
for f in E default C; do python - $O/r02v_bench_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('_bench_')[1], d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), d.get('loss_check'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
