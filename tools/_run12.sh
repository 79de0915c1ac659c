cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gradmodel_gpu.py -m gpu -q -x -k "two_level" 2>&1 | grep -v "^$" | tail -30
