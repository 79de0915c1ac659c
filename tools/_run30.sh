cd $GRAFT_REPO_ROOT
KT="timeout 300 python tools/kernel_times.py"
for cm in 64 96 128 160 192 256; do $KT 10000 64 40 64 chunk_max=$cm 2>&1 | tail -1; done
for cm in 64 128 256; do $KT 10000 64 40 128 chunk_max=$cm 2>&1 | tail -1; done
for cm in 192 224 256 320; do $KT 1000000 128 10 1 chunk_max=$cm 2>&1 | tail -1; done
