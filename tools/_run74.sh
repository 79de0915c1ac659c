cd $GRAFT_REPO_ROOT
KT="timeout 300 python tools/kernel_times.py"
$KT 10000 64 40 1 scene_nb=0 2>&1 | tail -1
$KT 10000 64 40 1 scene_nb=4 2>&1 | tail -1
for cfgs in "10000 64 40 1" "10000 64 40 64" "1000000 128 10 1"; do
for v in ss16 ss20; do DEXDEFORM_B200_LIB=$PWD/build_variants/lib_$v.so $KT $cfgs 2>&1 | tail -1 | sed "s/^\[[^]]*\]/[$v]/"; done
done
timeout 300 python tools/e_pass_once.py 64 2>&1 | grep "pass ms"
DEXDEFORM_B200_LIB=$PWD/build_variants/lib_unionfwd.so timeout 300 python tools/e_pass_once.py 64 2>&1 | grep "pass ms"
DEXDEFORM_B200_LIB=$PWD/build_variants/lib_ss20.so timeout 300 python tools/e_pass_once.py 64 2>&1 | grep "pass ms"
