"""Lipschitz estimate of Local Penalization (500 + n samples, then L-BFGS-B with finite differences): the reference's
estimate_L through predictive_gradients (mean + variance gradient) versus the mean-only device path."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import GPModel, kern
from gpyopt_b200 import evaluators as evs
n, d = (int(sys.argv[1]) if len(sys.argv) > 1 else 4096), 10
rng = np.random.default_rng(0)
X = -4 + 10 * rng.random((n, d)); Y = np.sin(X).sum(1, keepdims=True); Y = (Y - Y.mean()) / Y.std()
m = GPModel(kernel=kern.RBF(d, lengthscale=3.0), noise_var=1e-2, max_iters=0, verbose=False)
m.updateModel(X, Y, None, None)
bounds = [(-4.0, 6.0)] * d
out = {"n": n, "d": d}
flows = (("mean_only", None),
         ("reference_flow_predictive_gradients", lambda x: m.model.predictive_gradients(x)[0][:, :, 0]))
for name, g in flows:
    np.random.seed(1); evs.estimate_L(m.model, bounds, g)
    np.random.seed(1); t0 = time.perf_counter(); L = evs.estimate_L(m.model, bounds, g)
    out[name + "_s"] = time.perf_counter() - t0; out[name + "_L"] = L
print(json.dumps(out))
