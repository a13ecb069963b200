"""Refit latency with and without the CUDA-graph replay (GPO_OPT_FIT_GRAPH), plus bitwise equality of the two states."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import Engine
d = 6
rng = np.random.default_rng(0)
for n in [int(a) for a in sys.argv[1:]] or [256, 512, 1024, 2048, 3072]:
    X = 1 + 9 * rng.random((n, d)); Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
    out = {"n": n}
    st = {}
    for g in (0, 1):
        eng = Engine(0)
        eng.set_option('fit_graph', g)
        ts = []
        for it in range(14):
            t0 = time.perf_counter()
            eng.fit('rbf', X, Y, [1.0 + 0.01 * it], [[2.0]], [1e-2])
            ts.append(time.perf_counter() - t0)
        out["fit_ms_graph%d" % g] = round(min(ts[3:]) * 1e3, 4)
        out["fit_ms_median_graph%d" % g] = round(float(np.median(ts[3:])) * 1e3, 4)
        eng.fit('rbf', X, Y, [1.0], [[2.0]], [1e-2])
        st[g] = eng.get_state() + [eng.get_loglik()[0]]
        eng.close()
    out["bitwise_equal"] = bool(all(np.array_equal(a, b) for a, b in zip(st[0], st[1])))
    print(json.dumps(out), flush=True)
