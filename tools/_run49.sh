cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02g_e2e_launches_D.csv python tools/e2e_step_once.py D 2>&1 | tail -3
python tools/launch_summary.py gpurun_out/r02g_e2e_launches_D.csv 40 | tee gpurun_out/r02g_e2e_launch_summary_D.txt
timeout 600 python tools/e2e_gradmodel_parts.py D 2>&1 | tail -4
