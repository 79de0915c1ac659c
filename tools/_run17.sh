cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_e2e_launches_D.csv python tools/e2e_step_once.py D 2>&1 | tail -3
wc -l gpurun_out/r02_e2e_launches_D.csv
