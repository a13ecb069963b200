"""Small end-to-end case for compute-sanitizer (refit + value/grad sweeps + MCMC + LP + top-k)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import Engine
rng = np.random.default_rng(0)
eng = Engine(0)
for (n, d, S, kern) in ((200, 3, 1, 'matern52'), (300, 6, 3, 'rbf')):
    X = rng.random((n, d)); Y = np.sin(4 * X).sum(1)
    eng.fit(kern, X, Y, np.ones(S), np.full((S, d), 0.5), np.full(S, 1e-2))
    Xs = rng.random((300, d))
    kind = 'EI_MCMC' if S > 1 else 'EI'
    eng.acq(kind, 0.01, Xs); eng.acq(kind, 0.01, Xs, grad=True); eng.predict(Xs); eng.predict_grad(Xs[:3])
    eng.lp_set(Xs[:2], [0.1, 0.2], [0.3, 0.3], 'none'); eng.acq(kind, 0.01, Xs, grad=True); eng.lp_set(None, None, None, None)
    print(eng.acq_topk(kind, 0.01, Xs, 4)[1], eng.get_loglik()[0][:1], eng.get_loglik_grad()[0][:2])
eng.close(); print('done')
