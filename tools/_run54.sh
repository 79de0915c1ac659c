cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gradmodel_gpu.py tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -2
for cfgs in "1000000 128 1" "10000 64 64"; do
timeout 300 python tools/time_obs.py $cfgs 2>&1 | tail -1
DEXDEFORM_B200_LIB=$PWD/build_variants/lib_obsold.so timeout 300 python tools/time_obs.py $cfgs 2>&1 | tail -1
done
KT="timeout 300 python tools/kernel_times.py"
for cfgs in "1000000 128 10 1" "10000 64 40 64"; do
$KT $cfgs 2>&1 | tail -1
DD_PDL=1 $KT $cfgs 2>&1 | tail -1 | sed "s/^\[[^]]*\]/[pdl]/"
done
