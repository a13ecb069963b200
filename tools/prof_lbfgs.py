"""A few lock-step L-BFGS rounds at the C5 shape (n=4096, d=10, 5 anchors, LP with 15 penalisers) for the ncu launch list."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d, M = 10, 5
rng = np.random.default_rng(0)
b = np.array([[-4., 6.]] * d)
X = -4 + 10 * rng.random((n, d)); Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
eng = Engine(0); eng.fit('rbf', X, Y, [1.0], [[3.0] * d], [1e-2])
Xs = -4 + 10 * rng.random((1000, d))
_, idx = eng.acq_topk('EI', 0.01, Xs, M)
eng.lp_set(Xs[:15], np.full(15, 0.5), np.full(15, 0.3), 'none')
t0 = time.perf_counter()
x, f, it, nf = eng.lbfgs('EI', 0.01, Xs[idx], b, maxiter=int(sys.argv[2]) if len(sys.argv) > 2 else 6, pgtol=1e-5, factr=1e7)
dt = time.perf_counter() - t0
print('ok', f, it, nf, 'rounds', nf.max(), 'ms/round %.3f' % (dt * 1e3 / nf.max()))
