cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
DD_PAIR_G2PG=1 timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_parity_large_gpu.py -m gpu -q 2>&1 | grep -v "^$" | tail -8
for cfgs in "1000000 128 10 1" "10000 64 40 64" "50000 64 40 8" "10000 64 40 1"; do
DD_PAIR_G2PG=0 timeout 300 python tools/kernel_times.py $cfgs 2>&1 | sed 's/^\[[^]]*\]/[single]/'
DD_PAIR_G2PG=1 timeout 300 python tools/kernel_times.py $cfgs 2>&1 | sed 's/^\[[^]]*\]/[pair]/'
done | tee gpurun_out/r2_kt_pair_g2pg.log
