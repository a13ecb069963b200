"""Stress test of the latency path (N <= 32 candidates): the same call repeated thousands of times must return the same bits.
Reports the number of calls whose f / df differ from the first call's and the largest difference."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import Engine
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
for (n, d, N, kernel, S, noise) in ((684, 5, 3, 'matern32', 5, 1e-3), (684, 5, 4, 'rbf', 5, 1e-3), (684, 5, 1, 'matern52', 5, 1e-3), (1100, 4, 7, 'rbf', 8, 1e-2),
                                  (684, 5, 3, 'rbf', 1, 1e-3), (2048, 6, 5, 'rbf', 1, 1e-2), (2048, 6, 3, 'matern52', 3, 1e-2), (300, 10, 8, 'matern52', 3, 1e-2),
                                  (4096, 10, 5, 'rbf', 1, 1e-2), (4096, 10, 2, 'rbf', 2, 1e-2)):
    rng = np.random.default_rng(801778405)
    X = rng.random((n, d)); Y = np.sin(3 * X).sum(1) + 0.1 * rng.standard_normal(n); Y = (Y - Y.mean()) / Y.std()
    Xs = rng.random((N, d))
    var = 0.5 + rng.random(S); ls = 0.3 + rng.random((S, d)); nz = noise * (0.5 + rng.random(S))
    eng = Engine(0)
    eng.fit(kernel, X, Y, var, ls, nz)
    kind = 'EI' if S == 1 else 'EI_MCMC'
    f0, df0 = eng.acq(kind, 0.01, Xs, grad=True)
    v0 = eng.acq(kind, 0.01, Xs)
    bad_g = bad_v = 0; worst = 0.0
    for r in range(reps):
        f, df = eng.acq(kind, 0.01, Xs, grad=True)
        if not (np.array_equal(f, f0) and np.array_equal(df, df0)):
            bad_g += 1; worst = max(worst, float(np.max(np.abs(f - f0) / np.maximum(np.abs(f0), 1e-300))))
        if r % 4 == 0:
            v = eng.acq(kind, 0.01, Xs)
            if not np.array_equal(v, v0):
                bad_v += 1
    print(json.dumps({"n": n, "d": d, "N": N, "kernel": kernel, "S": S, "calls": reps, "grad_calls_differing": bad_g, "value_calls_differing": bad_v,
                      "worst_rel_diff_f": worst}), flush=True)
    eng.close()
