cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r02m_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 > $O/r02m_smoke.txt
timeout 600 python bench.py > $O/r02m_bench_default.json 2> $O/r02m_bench_default.err
for w in A B C E; do timeout 600 python bench.py --workload $w --no-cpu-baseline > $O/r02m_bench_$w.json 2> $O/r02m_bench_$w.err; done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02m_bench_reference_arm.json 2> $O/r02m_bench_ref.err
cat $O/r02m_pytest_gpu.txt $O/r02m_smoke.txt
for f in default A B C E reference_arm; do python - $O/r02m_bench_$f.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('_bench_')[1], d.get('value'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), d.get('loss_check'), (d.get('roofline') or {}).get('path'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
