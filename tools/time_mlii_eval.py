"""One ML-II objective evaluation = refit + log-likelihood + its gradient (what L-BFGS-B calls per step)."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import Engine
d = 6
rng = np.random.default_rng(0)
eng = Engine(0)
for n in [int(a) for a in sys.argv[1:]] or [64, 512, 2048, 4096]:
    X = 1 + 9 * rng.random((n, d)); Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
    tf, tl, tg = [], [], []
    for it in range(10):
        t0 = time.perf_counter(); eng.fit('rbf', X, Y, [1.0 + 0.01 * it], [[2.0]], [1e-2]); t1 = time.perf_counter()
        eng.get_loglik(); t2 = time.perf_counter(); eng.get_loglik_grad(); t3 = time.perf_counter()
        tf.append(t1 - t0); tl.append(t2 - t1); tg.append(t3 - t2)
    print(json.dumps({"n": n, "fit_ms": min(tf[2:]) * 1e3, "loglik_ms": min(tl[2:]) * 1e3, "loglik_grad_ms": min(tg[2:]) * 1e3}))
