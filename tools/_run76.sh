cd $GRAFT_REPO_ROOT
O=gpurun_out
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default', d['value'], d['e2e']['value'])"
DEXDEFORM_B200_LIB=$PWD/build_variants/lib_unionfwd.so timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('unionfwd', d['value'], d['e2e']['value'])"
done
