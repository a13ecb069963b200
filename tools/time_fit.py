"""Refit latency (gpo_fit wall time incl. its one stream sync) for a few n; hyper-parameters change every call so the
(X, Y) upload is skipped as in ML-II / HMC."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import Engine
d = 6
rng = np.random.default_rng(0)
eng = Engine(0)
for n in [int(a) for a in sys.argv[1:]] or [20, 64, 128, 512, 1024, 2048, 4096]:
    X = 1 + 9 * rng.random((n, d)); Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
    ts = []
    for it in range(12):
        t0 = time.perf_counter()
        eng.fit('rbf', X, Y, [1.0 + 0.01 * it], [[2.0]], [1e-2])
        ts.append(time.perf_counter() - t0)
    print(json.dumps({"n": n, "fit_ms_min": min(ts[2:]) * 1e3, "fit_ms_median": float(np.median(ts[2:])) * 1e3,
                      "loglik": float(eng.get_loglik()[0][0])}))
