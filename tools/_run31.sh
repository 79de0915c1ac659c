cd $GRAFT_REPO_ROOT
KT="timeout 300 python tools/kernel_times.py"
for cm in 272 288 320 352 416 512; do $KT 10000 64 40 64 chunk_max=$cm 2>&1 | tail -1; done
for cm in 320 416 512; do $KT 10000 64 40 128 chunk_max=$cm 2>&1 | tail -1; done
