"""Wall-clock latency of one EI + gradient evaluation at a single point (raw engine call, host arrays) for several n."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import Engine
d = 6
rng = np.random.default_rng(0)
eng = Engine(0)
out = {}
for n in [int(a) for a in sys.argv[1:]] or [64, 512, 1024, 2048, 4096]:
    X = 1 + 9 * rng.random((n, d)); Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
    eng.fit('rbf', X, Y, [1.0], [[2.0]], [1e-2])
    x = 1 + 9 * rng.random((1, d)); x5 = 1 + 9 * rng.random((5, d))
    res = {}
    for name, fn in (("f_df_N1", lambda: eng.acq('EI', 0.01, x, grad=True)), ("f_N1", lambda: eng.acq('EI', 0.01, x)),
                     ("f_df_N5", lambda: eng.acq('EI', 0.01, x5, grad=True))):
        for _ in range(30): fn()
        best = 1e9
        for rep in range(5):
            t0 = time.perf_counter()
            for _ in range(200): fn()
            best = min(best, (time.perf_counter() - t0) / 200)
        res[name] = round(best * 1e3, 4)
    out[n] = res
print(json.dumps(out))
