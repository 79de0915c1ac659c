cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_p2g_tile|k_g2p_tile|k_grid_b" -s 15 -c 3 -o gpurun_out/r02d_prof_fwd -f python tools/kernel_times.py 1000000 128 4 1 > gpurun_out/r02d_prof_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_g2p_grad_tile|k_p2g_grad_tile|k_grid_grad_b" -s 12 -c 3 -o gpurun_out/r02d_prof_bwd -f python tools/kernel_times.py 1000000 128 4 1 > gpurun_out/r02d_prof_bwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_grid_grad_b|k_grid_b" -s 24 -c 2 -o gpurun_out/r02d_prof_gridE -f python tools/kernel_times.py 10000 64 8 64 > gpurun_out/r02d_prof_gridE.log 2>&1
tail -2 gpurun_out/r02d_prof_gridE.log
ls -la gpurun_out | tail
