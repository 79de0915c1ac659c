cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q 2>&1 | tail -3
for r in 0 40 20 10; do timeout 300 python tools/kernel_times.py 1000000 128 80 1 resort_interval=$r; done 2>&1 | tee gpurun_out/r2_kt_resort.log
