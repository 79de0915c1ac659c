cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/e2e_gradmodel_parts.py D 2>&1 | tail -5 | tee gpurun_out/r02_e2e_parts_D.log
nproc; free -g | head -2
