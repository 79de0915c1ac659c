cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_g2p_grad_tile|k_p2g_grad_tile|k_grid_grad_b" -s 12 -c 3 -o $O/r02u_prof_bwd -f python tools/kernel_times.py 1000000 128 4 1 > $O/r02u_prof_bwd.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02u_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r02u_launches.log 2>&1
for i in 1 2 3; do timeout 600 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider 2>&1 | grep -E "^/|^E |^FAILED|passed|failed" | cut -c1-300 | head -8; done
ls -la $O | grep r02u
