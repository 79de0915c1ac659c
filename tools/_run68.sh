cd $GRAFT_REPO_ROOT
timeout 600 python tools/stress_test.py tests/test_engine_gpu.py test_resort_in_pieces_and_stale_slots 60 2>&1 | tail -8
