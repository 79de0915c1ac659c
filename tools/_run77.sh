cd $GRAFT_REPO_ROOT
for tool in memcheck racecheck; do
  ( time timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/san_scenes.py > gpurun_out/r02t_sanitizer_$tool.log 2>&1 ) 2>&1 | grep real
  echo "== $tool"; tail -4 gpurun_out/r02t_sanitizer_$tool.log
done
