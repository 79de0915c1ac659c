cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_hand.py -m gpu -q -x 2>&1 | grep -v "^$" | tail -30
