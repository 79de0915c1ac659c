cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_parity_large_gpu.py tests/test_parity_gpu.py -m gpu -q 2>&1 | grep -v "^$" | tail -8
for cfgs in "1000000 128 10 1" "10000 64 40 64" "50000 64 40 8"; do
timeout 300 python tools/kernel_times.py $cfgs 2>&1 | sed 's/^\[[^]]*\]/[ffma2]/'
done | tee gpurun_out/r2_kt_ffma2.log
