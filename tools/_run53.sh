cd $GRAFT_REPO_ROOT
KT="timeout 300 python tools/kernel_times.py"
for cfgs in "1000000 128 10 1" "10000 64 40 64" "10000 64 40 1"; do
$KT $cfgs 2>&1 | tail -1
DD_PDL=1 $KT $cfgs 2>&1 | tail -1 | sed "s/^\[[^]]*\]/[pdl]/"
done
