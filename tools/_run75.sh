cd $GRAFT_REPO_ROOT
O=gpurun_out
for i in 1 2 3; do timeout 600 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider 2>&1 | grep -E "^/|^E |^FAILED|passed|failed" | cut -c1-300 | head -8; done
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py > $O/r02s_bench_default.json 2> $O/r02s_bench_default.err
for w in A B E; do timeout 600 python bench.py --workload $w --no-cpu-baseline > $O/r02s_bench_$w.json 2> $O/r02s_bench_$w.err; done
for f in default A B E; do python - $O/r02s_bench_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('_bench_')[1], d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), d.get('loss_check'), (d.get('roofline') or {}).get('kernels_us'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
