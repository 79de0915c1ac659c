cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_p2g_tile" -s 6 -c 1 -o gpurun_out/r2_prof_p3 -f python tools/kernel_times.py 1000000 128 4 1 > gpurun_out/r2_prof_p3.log 2>&1
for i in 1 2 3 4 5 6; do timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q -x 2>&1 | grep "^E  \|passed\|failed" | head -8; done
