cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q 2>&1 | grep -v "^$" | tail -30
for p in 0 1; do DD_PDL=$p timeout 300 python tools/kernel_times.py 1000000 128 40 1; DD_PDL=$p timeout 300 python tools/kernel_times.py 10000 64 40 1; DD_PDL=$p timeout 300 python tools/kernel_times.py 10000 64 40 64; done 2>&1 | tee gpurun_out/r2_kt_pdl.log
