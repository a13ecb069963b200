"""Oracle-backed stand-in for gpyopt_b200.Engine -- TEST INFRASTRUCTURE ONLY.

Lets the CPU test-suite drive the host-side plugin classes (gpyopt_b200.models / .acquisitions / .gp) through the
reference's real BO loop in the build container, where there is no GPU, and lets the GPU tests compare an
engine-driven optimisation with an oracle-driven one step by step.  Implements the Engine methods with
oracle/gpy_numpy.py + oracle/gpyopt_numpy.py."""
import numpy as np

from oracle import gpy_numpy as G
from oracle import gpyopt_numpy as og

_K = {'rbf': G.RBF, 'matern52': G.Matern52, 'matern32': G.Matern32}


class FakeEngine(object):
    def __init__(self, device=0):
        self.device, self.n, self.d, self.S = device, 0, 0, 0
        self.gps = []
        self._lp = None

    def close(self):
        pass

    def fit(self, kernel, X, Y, variance, lengthscale, noise):
        X = np.asarray(X, dtype=float)
        Y = np.asarray(Y, dtype=float).reshape(-1, 1)
        variance = np.asarray(variance, dtype=float).reshape(-1)
        S, d = variance.size, X.shape[1]
        ls = np.asarray(lengthscale, dtype=float).reshape(S, -1)
        noise = np.asarray(noise, dtype=float).reshape(-1)
        self.gps = []
        for z in range(S):
            ard = ls.shape[1] > 1 and not np.all(ls[z] == ls[z][0])
            kern = _K[kernel](d, variance=variance[z], lengthscale=ls[z] if ard else ls[z][:1], ARD=ard)
            self.gps.append(G.GPRegression(X, Y, kern, noise_var=noise[z]))
        self._ard = ls.shape[1] > 1
        self.n, self.d, self.S = X.shape[0], d, S

    def get_fmin(self):
        return np.array([og.gp_get_fmin(g) for g in self.gps])

    def get_loglik(self):
        return np.array([g.log_likelihood() for g in self.gps]), np.zeros(self.S)

    def get_loglik_grad(self):
        out = np.zeros((self.S, self.d + 2))
        for z, g in enumerate(self.gps):
            gr = g.gradient
            out[z, 0], out[z, -1] = gr[0], gr[-1]
            if g.kern.ARD:
                out[z, 1:-1] = gr[1:-1]
            else:
                out[z, 1] = gr[1]  # the product sums the per-dimension entries for non-ARD kernels
        return out

    def predict(self, Xs, with_noise=True, out=None, stream=None):
        Xs = np.atleast_2d(np.asarray(Xs, dtype=float))
        res = [og.gp_predict(g, Xs, with_noise) for g in self.gps]
        return np.array([r[0][:, 0] for r in res]), np.array([r[1][:, 0] for r in res])

    def predict_grad(self, Xs, out=None, stream=None):
        Xs = np.atleast_2d(np.asarray(Xs, dtype=float))
        res = [og.gp_predict_withGradients(g, Xs) for g in self.gps]
        return (np.array([r[0][:, 0] for r in res]), np.array([r[1][:, 0] for r in res]),
                np.array([r[2] for r in res]), np.array([r[3] for r in res]))

    def predict_mean_grad(self, Xs, out=None, stream=None):
        m, _, dm, _ = self.predict_grad(Xs)
        return m, dm

    def posterior_cov(self, X1, X2=None, with_noise=False):
        g = self.gps[0]
        X1 = np.atleast_2d(np.asarray(X1, dtype=float))
        if X2 is None:
            return g.predict(X1, full_cov=True, include_likelihood=with_noise)[1]
        return g.posterior_covariance_between_points(X1, np.atleast_2d(np.asarray(X2, dtype=float)))

    def lp_set(self, X_batch, r, s, transform):
        self._lp = None if transform is None else (X_batch, r, s, transform)

    def acq(self, kind, par, Xs, grad=False, out=None, stream=None):
        Xs = np.atleast_2d(np.asarray(Xs, dtype=float))
        m, s, dm, ds = self.predict_grad(Xs)
        means, stds = [a[:, None] for a in m], [a[:, None] for a in s]
        fm = self.get_fmin()
        if kind == 'EI':
            f, df = og.acq_EI_withGradients(means[0], stds[0], dm[0], ds[0], fm[0], par)
        elif kind == 'MPI':
            f, df = og.acq_MPI_withGradients(means[0], stds[0], dm[0], ds[0], fm[0], par)
        elif kind == 'LCB':
            f, df = og.acq_LCB_withGradients(means[0], stds[0], dm[0], ds[0], par)
        elif kind == 'EI_MCMC':
            f, df = og.acq_EI_MCMC_withGradients(means, stds, list(dm), list(ds), list(fm), par)
        elif kind == 'MPI_MCMC':
            f, df = og.acq_MPI_MCMC_withGradients(means, stds, list(dm), list(ds), list(fm), par)
        else:
            f, df = og.acq_LCB_MCMC_withGradients(means, stds, list(dm), list(ds), par)
        if self._lp is not None:
            Xb, r, sc, transform = self._lp
            af, adf = og.acquisition_function_withGradients(f, df)
            val = og.lp_penalized_acquisition(af, Xs, Xb, r, sc, transform)
            if not grad:
                return val
            g = np.vstack([og.lp_d_acquisition(af[i:i + 1], adf[i:i + 1], Xs[i:i + 1], Xb, r, sc, transform)
                           for i in range(Xs.shape[0])])
            return val, g
        return (f[:, 0], df) if grad else f[:, 0]

    def acq_topk(self, kind, par, Xs, k, f_out=None, stream=None):
        f = self.acq(kind, par, Xs)
        score = f if self._lp is not None else -f
        idx = og.anchor_topk(score, k)
        return score[idx], idx

    def design_uniform(self, key, N, bounds, row0=0, out=None, stream=None):
        return og.design_uniform(key, int(N), bounds, row0=int(row0))

    def acq_topk_uniform(self, kind, par, key, N, bounds, k, row0=0, stream=None):
        X = self.design_uniform(key, N, bounds, row0)
        vals, idx = self.acq_topk(kind, par, X, k)
        return vals, idx, X[idx]
