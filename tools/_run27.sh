cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_grid_grad_b|k_grid_b" -s 24 -c 2 -o gpurun_out/r02d_prof_gridE -f python tools/kernel_times.py 10000 64 8 64 > gpurun_out/r02d_prof_gridE.log 2>&1
tail -2 gpurun_out/r02d_prof_gridE.log
KT="timeout 300 python tools/kernel_times.py"
for v in v0 v1 v2 v3 v4; do
DEXDEFORM_B200_LIB=$PWD/build_variants/lib_$v.so $KT 1000000 128 10 1 2>&1 | tail -1 | sed "s/^\[[^]]*\]/[$v]/"
done | tee gpurun_out/r02d_kt_v.log
