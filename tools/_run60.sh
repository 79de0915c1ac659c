cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_grid_grad_b|k_grid_b" -s 60 -c 2 -o gpurun_out/r02k_prof_gridA -f python tools/kernel_times.py 10000 64 40 1 > gpurun_out/r02k_prof_gridA.log 2>&1
tail -1 gpurun_out/r02k_prof_gridA.log | cut -c1-200
