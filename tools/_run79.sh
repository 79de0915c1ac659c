cd $GRAFT_REPO_ROOT
for E in 128 64; do
for cm in 256 288 320 352; do echo -n "E=$E chunk_max=$cm: "; DD_CHUNK_MAX=$cm timeout 300 python tools/e_pass_once.py $E 2>&1 | grep "pass ms"; done
done
