cd $GRAFT_REPO_ROOT
timeout 600 python tools/e2e_gradmodel_parts.py A 2>&1 | tail -3
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02j_e2e_launches_A.csv python tools/e2e_step_once.py A 2>&1 | tail -1
python tools/launch_summary.py gpurun_out/r02j_e2e_launches_A.csv 45
