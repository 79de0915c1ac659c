cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
KT="timeout 300 python tools/kernel_times.py 1000000 128 10 1"
for v in 1 2 3 5 7; do DD_WPB_P2GG=$v $KT 2>&1 | tail -1 | sed "s/^\[[^]]*\]/[p2gg wpb=$v]/"; done
for v in 1 2 3 6 7; do DD_WPB_G2P=$v $KT 2>&1 | tail -1 | sed "s/^\[[^]]*\]/[g2p wpb=$v]/"; done
for v in 1 2 3 5; do DD_WPB_P2G=$v $KT 2>&1 | tail -1 | sed "s/^\[[^]]*\]/[p2g wpb=$v]/"; done
for v in 2 3 4; do DD_WPB_G2PG=$v $KT 2>&1 | tail -1 | sed "s/^\[[^]]*\]/[g2pg wpb=$v]/"; done
