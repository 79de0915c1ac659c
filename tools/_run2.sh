cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > gpurun_out/r2_gputest2.log
tail -40 gpurun_out/r2_gputest2.log
