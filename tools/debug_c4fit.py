import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import Engine
def attempt(tag, n, d, S, bn, lo=1.0, span=9.0, ls0=2.0, prefit=None):
    rng = np.random.default_rng(0)
    X = lo + span * rng.random((n, d)); Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
    var = np.full(S, 1.0) * np.exp(0.05 * rng.standard_normal(S)) if S > 1 else np.array([1.0])
    eng = Engine(0)
    eng.set_option('bn_fit', bn)
    try:
        if prefit:
            eng.fit('rbf', X[:prefit], Y[:prefit], [1.0], np.full((1, d), ls0), [1e-2])
        eng.fit('rbf', X, Y, var, np.full((S, d), ls0), np.full(S, 1e-2))
        ll = eng.get_loglik()[0]
        K = np.exp(-0.5 * ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1) / ls0 ** 2)
        ref = []
        for z in (0, S - 1):
            Ky = var[z] * K + (1e-2 + 1e-8) * np.eye(n)
            L = np.linalg.cholesky(Ky); al = np.linalg.solve(Ky, Y)
            ref.append(0.5 * (-n * np.log(2 * np.pi) - 2 * np.log(np.diag(L)).sum() - Y @ al))
        print(tag, 'ok', ll[0], ref[0], ll[-1], ref[-1])
    except Exception as e:
        print(tag, 'FAILED', type(e).__name__, eng._lib.gpo_last_error(eng._h).decode())
    eng.close()
for bn in (128, 64):
    attempt('S=64 n=512 bn=%d' % bn, 512, 4, 64, bn)
    attempt('S=1 n=512 bn=%d' % bn, 512, 4, 1, bn)
    attempt('S=2 n=512 bn=%d' % bn, 512, 4, 2, bn)
    attempt('S=64 n=512 unitbox bn=%d' % bn, 512, 4, 64, bn, 0.0, 1.0, 0.5)
    attempt('S=64 n=512 prefit bn=%d' % bn, 512, 4, 64, bn, prefit=100)
    attempt('S=64 n=384 bn=%d' % bn, 384, 4, 64, bn)
    attempt('S=8 n=1024 bn=%d' % bn, 1024, 4, 8, bn)
