import os, sys, time, cProfile, pstats
import numpy as np, scipy.optimize
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import GPModel, AcquisitionEI, AcquisitionLP, kern
from gpyopt_b200 import evaluators as evs
n, d, B = 4096, 10, 16
rng = np.random.default_rng(0)
bounds = [(-4.0, 6.0)] * d; lo, hi = np.array(bounds).T
X = lo + (hi - lo) * rng.random((n, d)); Y = np.sin(X).sum(1, keepdims=True); Y = (Y - Y.mean()) / Y.std()
model = GPModel(kernel=kern.RBF(d, lengthscale=3.0), noise_var=1e-2, max_iters=0, verbose=False)
model.updateModel(X, Y, None, None)
ei = AcquisitionEI(model, None, None, jitter=0.01)
lp = AcquisitionLP(model, None, None, ei, transform='none')
eng = model.model.engine
np.random.seed(3)
def optimise_once():
    cand = lo + (hi - lo) * np.random.rand(1000, d)
    _, idx = eng.acq_topk('EI', 0.01, cand, 5)
    best = (np.inf, None)
    for a0 in cand[idx]:
        def f_df(x):
            f, df = lp.acquisition_function_withGradients(np.atleast_2d(x)); return float(np.ravel(f)[0]), np.ravel(df)
        res = scipy.optimize.fmin_l_bfgs_b(f_df, x0=a0, bounds=bounds, maxiter=1000)
        if res[1] < best[0]: best = (res[1], res[0])
    return best[1]
def run():
    lp.update_batches(None, None, None)
    batch = optimise_once()[None, :]
    L = evs.estimate_L(model.model, bounds); y_min = model.model.Y.min()
    for _ in range(1, B):
        lp.update_batches(batch, L, y_min)
        batch = np.vstack([batch, optimise_once()])
    lp.update_batches(None, None, None)
run()
pr = cProfile.Profile(); pr.enable(); t0 = time.perf_counter(); run(); print('batch s', time.perf_counter() - t0); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(12)
