cd $GRAFT_REPO_ROOT
KT="timeout 300 python tools/kernel_times.py"
for cfgs in "10000 64 40 1" "10000 64 40 64" "1000000 128 10 1"; do
$KT $cfgs 2>&1 | tail -1
for v in noinl; do DEXDEFORM_B200_LIB=$PWD/build_variants/lib_$v.so $KT $cfgs 2>&1 | tail -1 | sed "s/^\[[^]]*\]/[$v]/"; done
done
