import os, sys, time, cProfile, pstats
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import GPModel, AcquisitionEI, AcquisitionLP, kern
n, d = 4096, 10
rng = np.random.default_rng(0)
X = -4 + 10 * rng.random((n, d)); Y = np.sin(X).sum(1, keepdims=True); Y = (Y - Y.mean()) / Y.std()
m = GPModel(kernel=kern.RBF(d, lengthscale=3.0), noise_var=1e-2, max_iters=0, verbose=False)
m.updateModel(X, Y, None, None)
ei = AcquisitionEI(m, None, None, jitter=0.01)
lp = AcquisitionLP(m, None, None, ei, transform='none')
lp.update_batches(X[:8] + 0.1, 2.0, float(Y.min()))
x = -4 + 10 * rng.random((1, d))
eng = m.model.engine
for name, fn in (("raw eng.acq", lambda: eng.acq('EI', 0.01, x, grad=True)),
                 ("EI plugin f_df", lambda: ei.acquisition_function_withGradients(x)),
                 ("LP plugin f_df", lambda: lp.acquisition_function_withGradients(x))):
    for _ in range(20): fn()
    t0 = time.perf_counter()
    for _ in range(300): fn()
    print("%-16s %.3f ms" % (name, (time.perf_counter() - t0) / 300 * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(300): lp.acquisition_function_withGradients(x)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
