"""Run one GPU test function repeatedly in-process and report which assertion fails how often.
Usage: python tools/stress_test.py tests/test_engine_gpu.py test_resort_in_pieces_and_stale_slots 40"""
import importlib.util, os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
path, name, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
spec = importlib.util.spec_from_file_location("t", os.path.join(ROOT, path))
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
fn = getattr(mod, name)
fails = {}
for i in range(n):
    try:
        fn()
    except Exception as e:  # noqa: BLE001
        tb = traceback.extract_tb(e.__traceback__)
        key = "; ".join(f"{os.path.basename(f.filename)}:{f.lineno}" for f in tb[-3:]) + " | " + str(e)[:300].replace("\n", " ")
        fails[key] = fails.get(key, 0) + 1
print(f"{name}: {n - sum(fails.values())} / {n} passed")
for k, v in fails.items():
    print(f"  {v} x {k}")
