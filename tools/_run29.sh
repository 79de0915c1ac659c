cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
KT="timeout 300 python tools/kernel_times.py"
$KT 1000000 128 10 1 2>&1 | tail -1
$KT 10000 64 40 64 2>&1 | tail -1
$KT 10000 64 40 1 2>&1 | tail -1
$KT 50000 64 40 8 2>&1 | tail -1
