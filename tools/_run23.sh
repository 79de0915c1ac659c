cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
KT="timeout 300 python tools/kernel_times.py"
for cfgs in "1000000 128 10 1" "10000 64 40 64"; do
$KT $cfgs 2>&1 | tail -1
DD_BRICK_BLOCKS=592 $KT $cfgs 2>&1 | tail -1 | sed 's/^\[[^]]*\]/[bb592]/'
DD_BRICK_BLOCKS=1184 $KT $cfgs 2>&1 | tail -1 | sed 's/^\[[^]]*\]/[bb1184]/'
DD_BRICK_BLOCKS=2368 $KT $cfgs 2>&1 | tail -1 | sed 's/^\[[^]]*\]/[bb2368]/'
DEXDEFORM_B200_LIB=$PWD/build_variants/lib_noposeatom.so $KT $cfgs 2>&1 | tail -1
done | tee gpurun_out/r02d_kt_grid.log
