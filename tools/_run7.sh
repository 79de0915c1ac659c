cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q -k "runaway" 2>&1 | grep "^E " | cut -c1-400 | head -30
