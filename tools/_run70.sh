cd $GRAFT_REPO_ROOT
timeout 600 python tools/stress_test.py tests/test_engine_gpu.py test_tiled_path_follows_runaway_particles 80 2>&1 | tail -8
timeout 600 python tools/stress_test.py tests/test_engine_gpu.py test_resort_in_pieces_and_stale_slots 80 2>&1 | tail -8
for i in 1 2 3 4 5; do timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_gradmodel_gpu.py tests/test_parity_gpu.py tests/test_parity_large_gpu.py -m gpu -q --tb=line -p no:cacheprovider 2>&1 | grep -E "^/|^E |^FAILED|passed|failed" | cut -c1-300 | head -8; done
