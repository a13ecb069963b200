"""Refit for n >= 4096 on both schedules (fit_chain_max 31 -> round-1 look-ahead scheme, 64 -> short chain): latency, error message
if any, and the factor against numpy.linalg.cholesky.  (This is the script that exposed the in-place panel race of the look-ahead
scheme from 35 block columns on: DESIGN.md section 5, K4.)"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import Engine
d = 6
rng = np.random.default_rng(0)
for n in [int(a) for a in sys.argv[1:]] or [4096, 4224, 4608, 5120, 6144]:
    X = 1 + 9 * rng.random((n, d)); Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
    Z = X / 2.0; sq = (Z * Z).sum(1)
    Ky = np.exp(-0.5 * np.maximum(sq[:, None] + sq[None, :] - 2 * Z @ Z.T, 0.0)) + (1e-2 + 1e-8) * np.eye(n)
    Lref = np.linalg.cholesky(Ky)
    for cm in (31, 64):
        eng = Engine(0); eng.set_option('fit_chain_max', cm)
        out = {"n": n, "chain_max": cm}
        try:
            eng.fit('rbf', X, Y, [1.0], [[2.0]], [1e-2])
            L = eng.get_state()[0][0]
            out["dL"] = float(np.max(np.abs(np.tril(L) - Lref)))
            ts = []
            for it in range(8):
                t0 = time.perf_counter(); eng.fit('rbf', X, Y, [1.0 + 0.01 * it], [[2.0]], [1e-2]); ts.append(time.perf_counter() - t0)
            out["fit_ms_min"] = min(ts[2:]) * 1e3
        except Exception as e:
            out["error"] = repr(e) + ' | ' + eng._lib.gpo_last_error(eng._h).decode()
        print(json.dumps(out), flush=True)
        eng.close()
