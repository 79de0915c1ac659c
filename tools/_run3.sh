cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -80 > gpurun_out/r2_gputest3.log
grep -n "parity\|horizon\|final position\|passed\|failed\|FAILED" gpurun_out/r2_gputest3.log
for m in 1 2; do DD_G2PG_MODE=$m timeout 300 python tools/kernel_times.py 1000000 128 10 1; done 2>&1 | tee gpurun_out/r2_kt_g2pg.log
for m in 1 2; do DD_G2PG_MODE=$m timeout 300 python tools/kernel_times.py 10000 64 40 64; done 2>&1 | tee -a gpurun_out/r2_kt_g2pg.log
