cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_E_pass_launches.csv python tools/e_pass_once.py 64 2>&1 | tail -3
python tools/launch_summary.py gpurun_out/r02_E_pass_launches.csv 45 | tee gpurun_out/r02_E_pass_summary.txt
