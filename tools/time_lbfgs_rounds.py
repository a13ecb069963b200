"""Device L-BFGS: time per lock-step round net of the call's fixed cost (two calls with different maxiter, same anchors)."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import Engine
d, M = 10, 5
for n in [int(a) for a in sys.argv[1:]] or [300, 1024, 2048, 4096]:
    rng = np.random.default_rng(0)
    b = np.array([[-4., 6.]] * d)
    X = -4 + 10 * rng.random((n, d)); Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
    eng = Engine(0); eng.fit('rbf', X, Y, [1.0], [[3.0] * d], [1e-2])
    Xs = -4 + 10 * rng.random((1000, d))
    _, idx = eng.acq_topk('EI', 0.01, Xs, M)
    eng.lp_set(Xs[:15], np.full(15, 0.5), np.full(15, 0.3), 'none')
    res = {}
    for mi in (4, 24):
        best = 1e9
        for rep in range(5):
            t0 = time.perf_counter()
            x, f, it, nf = eng.lbfgs('EI', 0.01, Xs[idx], b, maxiter=mi, pgtol=1e-12, factr=0.0)
            best = min(best, time.perf_counter() - t0)
        res[mi] = (best, int(nf.max()))
    (t4, r4), (t24, r24) = res[4], res[24]
    print(json.dumps({"n": n, "rounds": [r4, r24], "call_ms": [round(t4 * 1e3, 3), round(t24 * 1e3, 3)],
                      "us_per_round": round((t24 - t4) / max(1, r24 - r4) * 1e6, 1)}), flush=True)
    eng.close()
