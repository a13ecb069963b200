"""Wall-clock latency of EI + gradient calls with 9 ... 256 candidates (the k-split tensor-core path) for n = 2048, 4096."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import Engine
d = 6
rng = np.random.default_rng(0)
eng = Engine(0)
out = {}
for n in (2048, 4096):
    X = 1 + 9 * rng.random((n, d)); Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
    eng.fit('rbf', X, Y, [1.0], [[2.0]], [1e-2])
    res = {}
    for N in (8, 9, 16, 64, 128, 256, 257, 1000):
        x = 1 + 9 * rng.random((N, d))
        for _ in range(10): eng.acq('EI', 0.01, x, grad=True)
        t0 = time.perf_counter()
        for _ in range(50): eng.acq('EI', 0.01, x, grad=True)
        res[N] = round((time.perf_counter() - t0) / 50 * 1e3, 4)
    out[n] = res
print(json.dumps(out))
