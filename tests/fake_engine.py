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

    def lbfgs(self, kind, par, x0, bounds, maxiter=1000, pgtol=1e-5, factr=1e7):
        """NumPy mirror of the device optimiser's state machine (csrc/kernels.cuh: lbfgs_update_kernel), anchor by anchor:
        lets the CPU suite check the algorithm (convergence, bounds, stopping rules) and the wrapper's plumbing."""
        x0 = np.atleast_2d(np.asarray(x0, dtype=float))
        lo, hi = np.asarray(bounds, dtype=float)[:, 0], np.asarray(bounds, dtype=float)[:, 1]
        M, d = x0.shape
        X, Fv = np.empty((M, d)), np.empty(M)
        iters, nfev = np.zeros(M, dtype=np.int32), np.zeros(M, dtype=np.int32)

        def evaluate(x):
            if self._lp is not None:
                v, g = self.acq(kind, par, x[None, :], grad=True)
                return float(v[0]), np.asarray(g[0], dtype=float)
            f, df = self.acq(kind, par, x[None, :], grad=True)
            return -float(f[0]), -np.asarray(df[0], dtype=float)

        for a in range(M):
            X[a], Fv[a], iters[a], nfev[a] = lbfgs_mirror(evaluate, x0[a], lo, hi, maxiter, pgtol, factr)
        return X, Fv, iters, nfev


# ---- mirror of csrc/kernels.cuh: lb_dcstep / lb_dcsrch / lbfgs_update_kernel ----------------------------------------
def _dcstep(st, stp, fp, dp, stpmin, stpmax):
    stx, fx, dx, sty, fy, dy, brackt = st['stx'], st['fx'], st['gx'], st['sty'], st['fy'], st['gy'], st['brackt']
    sgnd = dp * (dx / abs(dx))
    if fp > fx:
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * np.sqrt(max(0.0, (theta / s) ** 2 - (dx / s) * (dp / s)))
        if stp < stx:
            gamma = -gamma
        p = (gamma - dx) + theta; q = ((gamma - dx) + gamma) + dp; r = p / q
        stpc = stx + r * (stp - stx)
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx)
        stpf = stpc if abs(stpc - stx) < abs(stpq - stx) else stpc + (stpq - stpc) / 2.0
        brackt = True
    elif sgnd < 0.0:
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * np.sqrt(max(0.0, (theta / s) ** 2 - (dx / s) * (dp / s)))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta; q = ((gamma - dp) + gamma) + dx; r = p / q
        stpc = stp + r * (stx - stp)
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
        brackt = True
    elif abs(dp) < abs(dx):
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * np.sqrt(max(0.0, (theta / s) ** 2 - (dx / s) * (dp / s)))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta; q = (gamma + (dx - dp)) + gamma; r = p / q
        if r < 0.0 and gamma != 0.0:
            stpc = stp + r * (stx - stp)
        elif stp > stx:
            stpc = stpmax
        else:
            stpc = stpmin
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        if brackt:
            stpf = stpc if abs(stpc - stp) < abs(stpq - stp) else stpq
            stpf = min(stp + 0.66 * (sty - stp), stpf) if stp > stx else max(stp + 0.66 * (sty - stp), stpf)
        else:
            stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
            stpf = max(stpmin, min(stpmax, stpf))
    else:
        if brackt:
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp
            s = max(abs(theta), abs(dy), abs(dp))
            gamma = s * np.sqrt(max(0.0, (theta / s) ** 2 - (dy / s) * (dp / s)))
            if stp > sty:
                gamma = -gamma
            p = (gamma - dp) + theta; q = ((gamma - dp) + gamma) + dy; r = p / q
            stpf = stp + r * (sty - stp)
        elif stp > stx:
            stpf = stpmax
        else:
            stpf = stpmin
    if fp > fx:
        sty, fy, dy = stp, fp, dp
    else:
        if sgnd < 0.0:
            sty, fy, dy = stx, fx, dx
        stx, fx, dx = stp, fp, dp
    st.update(stx=stx, fx=fx, gx=dx, sty=sty, fy=fy, gy=dy, brackt=brackt)
    return stpf


def _dcsrch_start(st, stp, f, g, stpmax):
    st.update(brackt=False, stage=1, finit=f, ginit=g, gtest=1e-3 * g, width=stpmax, width1=2.0 * stpmax, stx=0.0, fx=f, gx=g,
              sty=0.0, fy=f, gy=g, stmin=0.0, stmax=stp + 4.0 * stp)


def _dcsrch_next(st, stp, f, g, stpmax):
    """-> (done, new stp).  MINPACK-2 dcsrch with L-BFGS-B's constants (ftol 1e-3, gtol 0.9, xtol 0.1, stpmin 0)."""
    ftest = st['finit'] + stp * st['gtest']
    if st['stage'] == 1 and f <= ftest and g >= 0.0:
        st['stage'] = 2
    warn = False
    if st['brackt'] and (stp <= st['stmin'] or stp >= st['stmax']):
        warn = True
    if st['brackt'] and st['stmax'] - st['stmin'] <= 0.1 * st['stmax']:
        warn = True
    if stp == stpmax and f <= ftest and g <= st['gtest']:
        warn = True
    if stp == 0.0 and (f > ftest or g >= st['gtest']):
        warn = True
    if f <= ftest and abs(g) <= 0.9 * (-st['ginit']):
        return True, stp
    if warn:
        return True, stp
    if st['stage'] == 1 and f <= st['fx'] and f > ftest:
        gt = st['gtest']
        sm = dict(stx=st['stx'], fx=st['fx'] - st['stx'] * gt, gx=st['gx'] - gt, sty=st['sty'], fy=st['fy'] - st['sty'] * gt,
                  gy=st['gy'] - gt, brackt=st['brackt'])
        stp = _dcstep(sm, stp, f - stp * gt, g - gt, st['stmin'], st['stmax'])
        st.update(stx=sm['stx'], sty=sm['sty'], brackt=sm['brackt'], fx=sm['fx'] + sm['stx'] * gt, fy=sm['fy'] + sm['sty'] * gt,
                  gx=sm['gx'] + gt, gy=sm['gy'] + gt)
    else:
        stp = _dcstep(st, stp, f, g, st['stmin'], st['stmax'])
    if st['brackt']:
        if abs(st['sty'] - st['stx']) >= 0.66 * st['width1']:
            stp = st['stx'] + 0.5 * (st['sty'] - st['stx'])
        st['width1'] = st['width']
        st['width'] = abs(st['sty'] - st['stx'])
        st['stmin'], st['stmax'] = min(st['stx'], st['sty']), max(st['stx'], st['sty'])
    else:
        st['stmin'] = stp + 1.1 * (stp - st['stx'])
        st['stmax'] = stp + 4.0 * (stp - st['stx'])
    stp = min(max(stp, 0.0), stpmax)
    if (st['brackt'] and (stp <= st['stmin'] or stp >= st['stmax'])) or (st['brackt'] and st['stmax'] - st['stmin'] <= 0.1 * st['stmax']):
        stp = st['stx']
    return False, stp


def lbfgs_mirror(evaluate, x0, lo, hi, maxiter, pgtol, factr, H=10):
    eps = 2.220446049250313e-16
    ftol = factr * eps
    d = x0.size
    x = np.clip(np.asarray(x0, dtype=float), lo, hi)
    S, Y, R = [], [], []
    f, g = evaluate(x)
    nfev, it = 1, 0

    def pgn(xx, gg):
        return float(np.max(np.abs(np.clip(xx - gg, lo, hi) - xx)))

    if not np.isfinite(f) or pgn(x, g) <= pgtol:
        return x, f, it, nfev
    steepest = True
    while True:
        # ---- direction: two-loop recursion on the free variables ----
        free = ~(((x <= lo) & (g > 0)) | ((x >= hi) & (g < 0)))
        hist = [] if steepest else list(range(len(S) - 1, -1, -1))
        while True:
            v = np.where(free, g, 0.0)
            al = []
            for h in hist:
                al.append(R[h] * S[h].dot(v)); v = v - al[-1] * Y[h]
            if hist:
                v = v / (R[-1] * Y[-1].dot(Y[-1]))
            for i in range(len(hist) - 1, -1, -1):
                h = hist[i]
                v = v + (al[i] - R[h] * Y[h].dot(v)) * S[h]
            dirn = np.where(free, -v, 0.0)
            # a free variable sitting on a bound and pushed outward by the quasi-Newton coupling: freeze it and redo
            block = free & (((x <= lo) & (dirn < 0)) | ((x >= hi) & (dirn > 0)))
            if not block.any():
                break
            free = free & ~block
        gd = g.dot(dirn)
        if not gd < 0 and hist:
            S, Y, R = [], [], []
            steepest = True
            continue
        dn = dirn.dot(dirn)
        if not dn > 0 or not gd < 0:
            break
        with np.errstate(divide='ignore', invalid='ignore'):
            room = np.where(dirn > 0, (hi - x) / dirn, np.where(dirn < 0, (lo - x) / dirn, np.inf))
        stpmx = float(min(np.min(room), 1e10))
        stp = min(1.0, stpmx)
        st = {}
        _dcsrch_start(st, stp, f, gd, stpmx)
        nls = 0
        accepted = False
        while True:
            xt = np.clip(x + stp * dirn, lo, hi)
            F, G = evaluate(xt); nfev += 1; nls += 1
            if not (np.isfinite(F) and np.all(np.isfinite(G))):
                F, gdt = st['fx'] + 1e10 * (1.0 + abs(st['fx'])), 1e10
            else:
                gdt = G.dot(dirn)
            done, stp_new = _dcsrch_next(st, stp, F, gdt, stpmx)
            if done:
                accepted = np.isfinite(F) and F < 1e299
                break
            if nls >= 20:
                break
            stp = stp_new
        if not accepted:
            if S:
                S, Y, R = [], [], []
                steepest = True
                continue
            break
        s, y = xt - x, G - g
        sy = s.dot(y)
        if sy > eps * (-gd * stp):
            S.append(s); Y.append(y); R.append(1.0 / sy)
            if len(S) > H:
                S.pop(0); Y.pop(0); R.pop(0)
        f0 = f
        x, f, g = xt, F, G
        it += 1
        steepest = False
        if pgn(x, g) <= pgtol or (f0 - f) <= ftol * max(abs(f0), abs(f), 1.0) or it >= maxiter:
            break
    return x, f, it, nfev
