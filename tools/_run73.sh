cd $GRAFT_REPO_ROOT
KT="timeout 300 python tools/kernel_times.py"
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_parity_large_gpu.py tests/test_gradmodel_gpu.py tests/test_parity_gpu.py tests/test_hand.py -m gpu -q --tb=line -p no:cacheprovider 2>&1 | grep -E "^/|^E |^FAILED|passed|failed" | cut -c1-300 | head -8
for cfgs in "10000 64 40 1" "10000 64 40 64" "50000 64 40 8" "1000000 128 10 1"; do
$KT $cfgs 2>&1 | tail -1
for v in unionfwd ss6; do DEXDEFORM_B200_LIB=$PWD/build_variants/lib_$v.so $KT $cfgs 2>&1 | tail -1 | sed "s/^\[[^]]*\]/[$v]/"; done
done
timeout 300 python tools/e_pass_once.py 64 2>&1 | tail -2
DEXDEFORM_B200_LIB=$PWD/build_variants/lib_unionfwd.so timeout 300 python tools/e_pass_once.py 64 2>&1 | tail -2
