cd $GRAFT_REPO_ROOT
KT="timeout 300 python tools/kernel_times.py"
$KT 1000000 128 80 1 2>&1 | tail -1
$KT 1000000 128 80 1 resort_interval=40 2>&1 | tail -1
$KT 1000000 128 80 1 resort_interval=20 2>&1 | tail -1
