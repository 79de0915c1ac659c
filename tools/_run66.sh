cd $GRAFT_REPO_ROOT
KT="timeout 300 python tools/kernel_times.py"
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_parity_large_gpu.py tests/test_gradmodel_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
for v in bpb1 bpb2; do DEXDEFORM_B200_LIB=$PWD/build_variants/lib_$v.so timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_parity_large_gpu.py -m gpu -x -q 2>&1 | tail -1; done
for cfgs in "10000 64 40 1" "10000 64 40 64" "50000 64 40 8" "1000000 128 10 1"; do
$KT $cfgs 2>&1 | tail -1
for v in noexit bpb1 bpb2; do DEXDEFORM_B200_LIB=$PWD/build_variants/lib_$v.so $KT $cfgs 2>&1 | tail -1 | sed "s/^\[[^]]*\]/[$v]/"; done
done
