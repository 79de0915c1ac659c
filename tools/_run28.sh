cd $GRAFT_REPO_ROOT
KT="timeout 300 python tools/kernel_times.py"
for v in v5 v6 v7; do
DEXDEFORM_B200_LIB=$PWD/build_variants/lib_$v.so $KT 1000000 128 10 1 2>&1 | tail -1 | sed "s/^\[[^]]*\]/[$v]/"
done
for v in v5 v7; do
DEXDEFORM_B200_LIB=$PWD/build_variants/lib_$v.so $KT 10000 64 40 64 2>&1 | tail -1 | sed "s/^\[[^]]*\]/[$v]/"
done
