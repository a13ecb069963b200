"""cProfile of the Branin BO loop (config 1) on the engine: where the host time goes."""
import os, sys, time, cProfile, pstats
import numpy as np, scipy.optimize
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import GPModel, AcquisitionEI, kern
bounds = [(-5., 10.), (1., 15.)]; lo, hi = np.array(bounds).T


def branin(X, a=1, b=5.1 / (4 * np.pi ** 2), c=5 / np.pi, r=6, s=10, t=1 / (8 * np.pi)):
    x1, x2 = X[:, 0], X[:, 1]
    return (a * (x2 - b * x1 ** 2 + c * x1 - r) ** 2 + s * (1 - t) * np.cos(x1) + s).reshape(-1, 1)


def bo(iters=15):
    np.random.seed(123)
    X = lo + (hi - lo) * np.random.rand(5, 2); Yraw = branin(X)
    model = GPModel(kernel=kern.RBF(2), exact_feval=True, optimize_restarts=5, verbose=False)
    for it in range(iters):
        Y = Yraw - Yraw.mean(); Y = Y / Y.std()
        model.updateModel(X, Y, None, None)
        acq = AcquisitionEI(model, None, None, jitter=0.01)
        cand = lo + (hi - lo) * np.random.rand(1000, 2)
        _, idx = model.model.engine.acq_topk('EI', 0.01, cand, 5)
        best = (np.inf, None)
        for a0 in cand[idx]:
            def f_df(x):
                f, df = acq.acquisition_function_withGradients(np.atleast_2d(x)); return float(f[0, 0]), df[0]
            res = scipy.optimize.fmin_l_bfgs_b(f_df, x0=a0, bounds=bounds, maxiter=1000)
            if res[1] < best[0]: best = (res[1], res[0])
        X = np.vstack([X, best[1]]); Yraw = np.vstack([Yraw, branin(best[1][None, :])])
    return float(Yraw.min())


bo(3)
pr = cProfile.Profile(); pr.enable(); t0 = time.perf_counter(); y = bo(); dt = time.perf_counter() - t0; pr.disable()
print('wall', dt, 'best', y)
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
