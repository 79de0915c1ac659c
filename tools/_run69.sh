cd $GRAFT_REPO_ROOT
for i in 1 2 3 4 5 6 7 8; do timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | grep -E "^FAILED|passed|failed" | head -6; done
