cd $GRAFT_REPO_ROOT
for cfgs in "1000000 128 10 1" "1000000 128 80 1" "10000 64 40 64" "50000 64 40 8"; do timeout 300 python tools/defer_count.py $cfgs 2>&1 | tail -1; done
