cd $GRAFT_REPO_ROOT
for i in 1 2 3 4; do timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "resort_in_pieces" 2>&1 | grep -E "^E |passed|failed" | head -8; done
echo == noexit
for i in 1 2 3; do DEXDEFORM_B200_LIB=$PWD/build_variants/lib_noexit.so timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "resort_in_pieces" 2>&1 | grep -E "^E |passed|failed" | head -8; done
echo == full file default
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q 2>&1 | grep -E "^E |passed|failed" | head -20
