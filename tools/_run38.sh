cd $GRAFT_REPO_ROOT
KT="timeout 300 python tools/kernel_times.py"
for pct in 100 85 75 65 50; do DD_PB_SCALE=$pct $KT 10000 64 40 64 2>&1 | tail -1 | sed "s/^\[[^]]*\]/[pb $pct]/"; done
for pct in 85 75; do DD_PB_SCALE=$pct $KT 1000000 128 10 1 2>&1 | tail -1 | sed "s/^\[[^]]*\]/[pb $pct]/"; done
for pct in 100 75 50; do DD_PB_SCALE=$pct $KT 10000 64 40 32 2>&1 | tail -1 | sed "s/^\[[^]]*\]/[pb $pct]/"; done
