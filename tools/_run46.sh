cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tool in racecheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/san_scenes.py > gpurun_out/r02g_sanitizer_$tool.log 2>&1
  echo "== $tool"; tail -4 gpurun_out/r02g_sanitizer_$tool.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_p2g_tile|k_g2p_tile|k_grid_b" -s 15 -c 3 -o gpurun_out/r02g_prof_fwd -f python tools/kernel_times.py 1000000 128 4 1 > gpurun_out/r02g_prof_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_g2p_grad_tile|k_p2g_grad_tile|k_grid_grad_b" -s 12 -c 3 -o gpurun_out/r02g_prof_bwd -f python tools/kernel_times.py 1000000 128 4 1 > gpurun_out/r02g_prof_bwd.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02g_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_launches.log 2>&1
tail -1 gpurun_out/r02g_prof_bwd.log | cut -c1-200
ls -la gpurun_out | tail -8
