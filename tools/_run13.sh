cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -30
for cfgs in "1000000 128 10 1" "10000 64 40 64"; do
timeout 300 python tools/kernel_times.py $cfgs 2>&1 | sed 's/^\[[^]]*\]/[default = syncwarp]/'
DEXDEFORM_B200_LIB=$PWD/build_variants/lib_nosyncwarp.so timeout 300 python tools/kernel_times.py $cfgs 2>&1 | sed 's/^\[[^]]*\]/[no syncwarp]/'
done | tee gpurun_out/r2_kt_syncwarp.log
cat > /tmp/san.py <<'PY'
import sys, numpy as np
sys.path[:0] = ["/root/repo", "/root/repo/tests"]
from dexdeform_b200.engine import FusedSim
from dexdeform_b200.scenes import make_scene
for n, E, cm in ((36, 2, 32), (1500, 2, 32), (4000, 1, 96)):
    S = 3
    w = 0.05 + 0.1 * min(1.0, n / 5000.0)
    sc = make_scene(n, 32, box_center=(0.5, 0.3, 0.5), box_width=(w, w, w), steps=S, perturb=0.02, vel_scale=0.5, on_floor=True, nb=4, seed=77)
    sim = FusedSim.from_scene(sc, n_envs=E, max_steps=S, chunk_max=cm, use_graphs=False)
    sim.forward(0, S)
    sim.zero_grad(S)
    g = np.zeros((E, n, 3), np.float32); g[..., 1] = -1.0 / n
    sim.add_state_grad(S, g)
    sim.backward(0, S)
    x = sim.get_state_grad(0)["x"]
    print("ok", n, E, cm, float(np.abs(x).sum()))
    sim.close()
PY
for lib in default nosyncwarp; do
  if [ $lib = nosyncwarp ]; then export DEXDEFORM_B200_LIB=$PWD/build_variants/lib_nosyncwarp.so; else unset DEXDEFORM_B200_LIB; fi
  for tool in racecheck memcheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/r2_sanitizer_${lib}_${tool}.log 2>&1
    echo "== $lib $tool"; tail -4 gpurun_out/r2_sanitizer_${lib}_${tool}.log
  done
done
