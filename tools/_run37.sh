cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for cfgs in "1000000 128 10 1" "1000000 128 80 1" "10000 64 40 64" "50000 64 40 8" "10000 64 40 1"; do timeout 300 python tools/defer_count.py $cfgs 2>&1 | tail -1; done
