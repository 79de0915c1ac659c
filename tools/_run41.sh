cd $GRAFT_REPO_ROOT
KT="timeout 300 python tools/kernel_times.py"
for cfgs in "1000000 128 10 1" "10000 64 40 64" "50000 64 40 8"; do
DEXDEFORM_B200_LIB=$PWD/build_variants/lib_drop.so $KT $cfgs 2>&1 | tail -1 | sed "s/^\[[^]]*\]/[drop]/"
done
