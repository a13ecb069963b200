"""One-off robustness sweep: the randomised oracle comparison of tests/test_gpu_parity.py over many more seeded cases
(sizes around every tile / slice / chunk / latency-path boundary).  Not collected by pytest (no test_ prefix):
  python tests/manual_random_sweep.py [count] [seed]"""
import os, sys, json, traceback
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import test_gpu_parity as T

count = int(sys.argv[1]) if len(sys.argv) > 1 else 150
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 7
rng = np.random.default_rng(seed)
edge_n = [1, 2, 3, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 383, 384, 385, 511, 512, 513, 1023, 1024, 1025, 1151]
edge_N = [1, 2, 3, 4, 5, 7, 8, 9, 16, 127, 128, 129, 255, 256, 257, 300, 513, 700]
fails = []
for i in range(count):
    n = int(rng.choice(edge_n)) if i % 2 == 0 else int(rng.integers(1, 1200))
    case = (n, int(rng.integers(1, 17)), int(rng.choice(edge_N)), ['rbf', 'matern52', 'matern32'][i % 3], bool(rng.integers(0, 2)),
            int(rng.choice([1, 1, 1, 2, 3, 5])), float(rng.choice([1e-1, 1e-2, 1e-3, 1e-4])), int(rng.integers(0, 1 << 30)))
    try:
        T.test_randomised_shapes_against_oracle(*case)
    except Exception as e:  # noqa: BLE001
        fails.append({"case": case, "error": traceback.format_exc().splitlines()[-1][:300]})
print(json.dumps({"cases": count, "failed": len(fails), "failures": fails[:10]}))
