"""NumPy restatement of the GPyOpt-side formulas on the hot path.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py).  Every function cites the reference file:line it follows; the GP
object passed in is a GPy-duck-typed model (oracle.gpy_numpy.GPRegression).
"""
import numpy as np
from scipy.special import erfc
from scipy.stats import norm


# -- GPyOpt/models/gpmodel.py -------------------------------------------------------------
def gp_predict(gp, X, with_noise=True):
    """GPModel.predict -> (m[N,1], s[N,1])   (gpmodel.py:95-112)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[None, :]
    m, v = gp.predict(X, full_cov=False, include_likelihood=with_noise)
    v = np.clip(v, 1e-10, np.inf)
    return m, np.sqrt(v)


def gp_predict_withGradients(gp, X):
    """GPModel.predict_withGradients -> (m, s, dmdx, dsdx)   (gpmodel.py:131-142)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[None, :]
    m, v = gp.predict(X)
    v = np.clip(v, 1e-10, np.inf)
    dmdx, dvdx = gp.predictive_gradients(X)
    dmdx = dmdx[:, :, 0]
    dsdx = dvdx / (2 * np.sqrt(v))
    return m, np.sqrt(v), dmdx, dsdx


def gp_get_fmin(gp):
    """GPModel.get_fmin (gpmodel.py:125-129): min posterior mean (with-noise predict) at training X."""
    return gp.predict(gp.X)[0].min()


def _mcmc_loop(gp, hmc_samples, fn):
    """GPModel_MCMC sample loop (gpmodel.py:257-324): write sample, refit, evaluate, restore."""
    ps = gp.param_array.copy()
    out = []
    for s in hmc_samples:
        if gp._fixes_ is None:
            gp[:] = s
        else:
            gp[gp._fixes_] = s
        gp._trigger_params_changed()
        out.append(fn(gp))
    gp.param_array[:] = ps
    gp._trigger_params_changed()
    return out


def gp_mcmc_predict(gp, hmc_samples, X):
    """GPModel_MCMC.predict -> (means, stds) lists (gpmodel.py:257-277)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[None, :]
    res = _mcmc_loop(gp, hmc_samples, lambda g: g.predict(X))
    return [r[0] for r in res], [np.sqrt(np.clip(r[1], 1e-10, np.inf)) for r in res]


def gp_mcmc_get_fmin(gp, hmc_samples):
    """GPModel_MCMC.get_fmin (gpmodel.py:279-295)."""
    return _mcmc_loop(gp, hmc_samples, lambda g: g.predict(g.X)[0].min())


def gp_mcmc_predict_withGradients(gp, hmc_samples, X):
    """GPModel_MCMC.predict_withGradients (gpmodel.py:297-324)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[None, :]

    def one(g):
        m, v = g.predict(X)
        std = np.sqrt(np.clip(v, 1e-10, np.inf))
        dmdx, dvdx = g.predictive_gradients(X)
        return m, std, dmdx[:, :, 0], dvdx / (2 * std)

    res = _mcmc_loop(gp, hmc_samples, one)
    return [r[0] for r in res], [r[1] for r in res], [r[2] for r in res], [r[3] for r in res]


# -- GPyOpt/util/general.py ---------------------------------------------------------------
def get_quantiles(acquisition_par, fmin, m, s):
    """get_quantiles (util/general.py:113-128); clamps s in place like the reference."""
    if isinstance(s, np.ndarray):
        s[s < 1e-10] = 1e-10
    elif s < 1e-10:
        s = 1e-10
    u = (fmin - m - acquisition_par) / s
    phi = np.exp(-0.5 * u ** 2) / np.sqrt(2 * np.pi)
    Phi = 0.5 * erfc(-u / np.sqrt(2))
    return phi, Phi, u


# -- GPyOpt/acquisitions/{EI,MPI,LCB}.py ---------------------------------------------------
def acq_EI(m, s, fmin, jitter):
    """AcquisitionEI._compute_acq (EI.py:32-40)."""
    phi, Phi, u = get_quantiles(jitter, fmin, m, s)
    return s * (u * Phi + phi)


def acq_EI_withGradients(m, s, dmdx, dsdx, fmin, jitter):
    """AcquisitionEI._compute_acq_withGradients (EI.py:42-51)."""
    phi, Phi, u = get_quantiles(jitter, fmin, m, s)
    return s * (u * Phi + phi), dsdx * phi - Phi * dmdx


def acq_MPI(m, s, fmin, jitter):
    """AcquisitionMPI._compute_acq (MPI.py:32-40)."""
    _, Phi, _ = get_quantiles(jitter, fmin, m, s)
    return Phi


def acq_MPI_withGradients(m, s, dmdx, dsdx, fmin, jitter):
    """AcquisitionMPI._compute_acq_withGradients (MPI.py:42-51)."""
    phi, Phi, u = get_quantiles(jitter, fmin, m, s)
    return Phi, -(phi / s) * (dmdx + dsdx * u)


def acq_LCB(m, s, exploration_weight):
    """AcquisitionLCB._compute_acq (LCB.py:35-41)."""
    return -m + exploration_weight * s


def acq_LCB_withGradients(m, s, dmdx, dsdx, exploration_weight):
    """AcquisitionLCB._compute_acq_withGradients (LCB.py:43-50)."""
    return -m + exploration_weight * s, -dmdx + exploration_weight * dsdx


# -- GPyOpt/acquisitions/{EI,MPI,LCB}_mcmc.py ----------------------------------------------
def acq_EI_MCMC(means, stds, fmins, jitter):
    """AcquisitionEI_MCMC._compute_acq (EI_mcmc.py:29-39): note the +jitter quirk."""
    f = 0
    for m, s, fmin in zip(means, stds, fmins):
        phi, Phi, _ = get_quantiles(jitter, fmin, m, s)
        f = f + (fmin - m + jitter) * Phi + s * phi
    return f / len(means)


def acq_EI_MCMC_withGradients(means, stds, dmdxs, dsdxs, fmins, jitter):
    """AcquisitionEI_MCMC._compute_acq_withGradients (EI_mcmc.py:41-59)."""
    f_acqu, df_acqu = None, None
    for m, s, fmin, dmdx, dsdx in zip(means, stds, fmins, dmdxs, dsdxs):
        phi, Phi, _ = get_quantiles(jitter, fmin, m, s)
        f = (fmin - m + jitter) * Phi + s * phi
        df = dsdx * phi - Phi * dmdx
        if f_acqu is None:
            f_acqu, df_acqu = f, df
        else:
            f_acqu = f_acqu + f
            df_acqu = df_acqu + df
    return f_acqu / len(means), df_acqu / len(means)


def acq_MPI_MCMC(means, stds, fmins, jitter):
    """AcquisitionMPI_MCMC._compute_acq (MPI_mcmc.py:29-39)."""
    f = 0
    for m, s, fmin in zip(means, stds, fmins):
        _, Phi, _ = get_quantiles(jitter, fmin, m, s)
        f = f + Phi
    return f / len(means)


def acq_MPI_MCMC_withGradients(means, stds, dmdxs, dsdxs, fmins, jitter):
    """AcquisitionMPI_MCMC._compute_acq_withGradients (MPI_mcmc.py:41-59)."""
    f_acqu, df_acqu = None, None
    for m, s, fmin, dmdx, dsdx in zip(means, stds, fmins, dmdxs, dsdxs):
        phi, Phi, u = get_quantiles(jitter, fmin, m, s)
        f = Phi
        df = -(phi / s) * (dmdx + dsdx * u)
        if f_acqu is None:
            f_acqu, df_acqu = f, df
        else:
            f_acqu = f_acqu + f
            df_acqu = df_acqu + df
    return f_acqu / len(means), df_acqu / len(means)


def acq_LCB_MCMC(means, stds, exploration_weight):
    """AcquisitionLCB_MCMC._compute_acq (LCB_mcmc.py:26-35)."""
    f = 0
    for m, s in zip(means, stds):
        f = f + (-m + exploration_weight * s)
    return f / len(means)


def acq_LCB_MCMC_withGradients(means, stds, dmdxs, dsdxs, exploration_weight):
    """AcquisitionLCB_MCMC._compute_acq_withGradients (LCB_mcmc.py:37-52)."""
    f_acqu, df_acqu = None, None
    for m, s, dmdx, dsdx in zip(means, stds, dmdxs, dsdxs):
        f = -m + exploration_weight * s
        df = -dmdx + exploration_weight * dsdx
        if f_acqu is None:
            f_acqu, df_acqu = f, df
        else:
            f_acqu = f_acqu + f
            df_acqu = df_acqu + df
    return f_acqu / len(means), df_acqu / len(means)


# -- GPyOpt/acquisitions/base.py -----------------------------------------------------------
def acquisition_function(f_acqu):
    """AcquisitionBase.acquisition_function with constant cost, no constraints (base.py:33-40)."""
    return -(f_acqu * 1.0) / np.ones((f_acqu.shape[0], 1))


def acquisition_function_withGradients(f_acqu, df_acqu):
    """AcquisitionBase.acquisition_function_withGradients, constant cost (base.py:43-52)."""
    cost_x = np.ones((f_acqu.shape[0], 1))
    cost_grad_x = np.zeros(df_acqu.shape)
    f_acq_cost = f_acqu / cost_x
    df_acq_cost = (df_acqu * cost_x - f_acqu * cost_grad_x) / (cost_x ** 2)
    return -f_acq_cost * 1.0, -df_acq_cost * 1.0


# -- GPyOpt/acquisitions/LP.py -------------------------------------------------------------
def lp_hammer_precompute(m_x0, std_x0, L, Min):
    """AcquisitionLP._hammer_function_precompute (LP.py:48-62); std_x0 is what GPModel.predict
    returns (a std) and the reference takes another sqrt of it (quirk)."""
    pred = std_x0.copy()
    pred[pred < 1e-16] = 1e-16
    s = np.sqrt(pred)
    r_x0 = (m_x0 - Min) / L
    s_x0 = s / L
    return r_x0.flatten(), s_x0.flatten()


def lp_hammer_function(x, x0, r_x0, s_x0):
    """AcquisitionLP._hammer_function (LP.py:64-68)."""
    return norm.logcdf((np.sqrt((np.square(np.atleast_2d(x)[:, None, :] - np.atleast_2d(x0)[None, :, :])).sum(-1))
                        - r_x0) / s_x0)


def lp_penalized_acquisition(neg_acq_col, x, X_batch, r_x0, s_x0, transform):
    """AcquisitionLP._penalized_acquisition (LP.py:70-89).  ``neg_acq_col`` is
    ``self.acq.acquisition_function(x)`` ([N,1], already negated)."""
    fval = -neg_acq_col[:, 0]
    if transform == 'softplus':
        fval_org = fval.copy()
        fval[fval_org >= 40.] = np.log(fval_org[fval_org >= 40.])
        fval[fval_org < 40.] = np.log(np.log1p(np.exp(fval_org[fval_org < 40.])))
    elif transform == 'none':
        fval = np.log(fval + 1e-50)
    fval = -fval
    if X_batch is not None:
        h_vals = lp_hammer_function(x, X_batch, r_x0, s_x0)
        fval += -h_vals.sum(axis=-1)
    return fval


def lp_d_hammer_function(x, X_batch, r_x0, s_x0):
    """AcquisitionLP._d_hammer_function (LP.py:91-103) -> [N,1] (direction dropped, quirk B.6)."""
    dx = np.atleast_2d(x)[:, None, :] - np.atleast_2d(X_batch)[None, :, :]
    nm = np.sqrt((np.square(dx)).sum(-1))
    z = (nm - r_x0) / s_x0
    h_func = norm.cdf(z)
    d = 1. / (s_x0 * np.sqrt(2 * np.pi) * h_func) * np.exp(-np.square(z) / 2) / nm
    d[h_func < 1e-50] = 0.
    d = d[:, :, None]
    return d.sum(axis=1)


def lp_d_acquisition(neg_acq_col, neg_grad, x, X_batch, r_x0, s_x0, transform):
    """AcquisitionLP.d_acquisition_function (LP.py:112-132)."""
    x = np.atleast_2d(x)
    if transform == 'softplus':
        fval = -neg_acq_col[:, 0]
        scale = 1. / (np.log1p(np.exp(fval)) * (1. + np.exp(-fval)))
    elif transform == 'none':
        fval = -neg_acq_col[:, 0]
        scale = 1. / fval
    else:
        scale = 1.
    if X_batch is None:
        return scale * neg_grad
    return scale * neg_grad - lp_d_hammer_function(x, X_batch, r_x0, s_x0)


# -- GPyOpt/optimization/anchor_points_generator.py ----------------------------------------
def anchor_topk(scores, k):
    """AnchorPointsGenerator.get (anchor_points_generator.py:67): indices of the k smallest
    scores.  The reference uses NumPy's default (unstable) argsort; ties are broken here by
    lowest index (kind='stable'), which is the engine's documented contract."""
    scores = np.asarray(scores).flatten()
    return np.argsort(scores, kind='stable')[:min(len(scores), k)]


# -- GPyOpt/objective_examples (synthetic data generators for the BASELINE configs) --------
def alpine2(X):
    """experimentsNd.py:59-67  alpine2.f: -prod(sqrt(x)) * prod(sin(x)), returned [n,1]."""
    X = np.atleast_2d(X)
    return (-(np.cumprod(np.sqrt(X), axis=1)[:, -1] * np.cumprod(np.sin(X), axis=1)[:, -1])).reshape(-1, 1)


def gSobol(X, a):
    """experimentsNd.py:89-99  gSobol.f."""
    X = np.atleast_2d(X)
    aux = (abs(4 * X - 2) + np.ones(X.shape[1]).reshape(1, -1) * a) / (1 + np.ones(X.shape[1]).reshape(1, -1) * a)
    return np.cumprod(aux, axis=1)[:, -1].reshape(-1, 1)


def branin(X, a=1, b=5.1 / (4 * np.pi ** 2), c=5 / np.pi, r=6, s=10, t=1 / (8 * np.pi)):
    """experiments2d.py:202-217  branin.f."""
    X = np.atleast_2d(X)
    x1, x2 = X[:, 0], X[:, 1]
    return (a * (x2 - b * x1 ** 2 + c * x1 - r) ** 2 + s * (1 - t) * np.cos(x1) + s).reshape(-1, 1)


def sixhumpcamel(X):
    """experiments2d.py:278-293  sixhumpcamel.f."""
    X = np.atleast_2d(X)
    x1, x2 = X[:, 0], X[:, 1]
    return ((4 - 2.1 * x1 ** 2 + x1 ** 4 / 3) * x1 ** 2 + x1 * x2 + (-4 + 4 * x2 ** 2) * x2 ** 2).reshape(-1, 1)


def normalize_stats(Y):
    """util/general.py:203-223 normalize(Y, 'stats')."""
    Y_norm = Y - Y.mean()
    std = Y.std()
    if std > 0:
        Y_norm = Y_norm / std
    return Y_norm


# --------------------------------------------------------------------------------------------------
# Device-side random design.  The engine generates uniform candidates with a counter-based
# generator instead of copying a host design; the CONTRACT of that generator is NumPy's own Philox bit generator
# (third-party: numpy >= 1.17, numpy/random/src/philox/philox.h, Philox4x64-10 of Salmon et al., "Parallel random
# numbers: as easy as 1, 2, 3", SC'11).  `design_uniform` below IS numpy; `philox4x64_10` restates the published
# algorithm in plain integers and is pinned against numpy's raw stream by tests/test_oracle_pins.py.
# The reference's samples_multidimensional_uniform (experiment_design/random_design.py:66-76) draws column by
# column from the legacy global MT19937 stream; a counter-based design has the same distribution, not the same
# numbers -- parity for this row is "same candidates in -> same scores / top-k out", with the candidates
# regenerated on the host from the same key.
# --------------------------------------------------------------------------------------------------
_PHILOX_M0, _PHILOX_M1 = 0xD2E7470EE14C6C93, 0xCA5A826395121157
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B97F4A7C15, 0xBB67AE8584CAA73B
_U64 = (1 << 64) - 1


def philox4x64_10(counter, key):
    """One Philox4x64-10 block: counter (4 words), key (2 words) -> 4 output words (plain Python integers)."""
    c, k = [int(v) for v in counter], [int(v) for v in key]
    for r in range(10):
        if r:
            k = [(k[0] + _PHILOX_W0) & _U64, (k[1] + _PHILOX_W1) & _U64]
        p0, p1 = _PHILOX_M0 * c[0], _PHILOX_M1 * c[2]
        c = [(p1 >> 64) ^ c[1] ^ k[0], p1 & _U64, (p0 >> 64) ^ c[3] ^ k[1], p0 & _U64]
    return c


def design_uniform_restated(key, N, bounds, row0=0):
    """Pure-Python restatement of the device generator (small N only): element e of the row-major design uses lane
    e & 3 of block (e >> 2) + 1 (numpy increments the counter before generating), u = (v >> 11) * 2**-53,
    x = lo + (hi - lo) * u."""
    bounds = np.atleast_2d(np.asarray(bounds, dtype=np.float64))
    d = bounds.shape[0]
    lo, rng = bounds[:, 0], bounds[:, 1] - bounds[:, 0]
    key = _philox_key(key)
    out = np.empty(N * d)
    cache = {}
    for j in range(N * d):
        e = row0 * d + j
        b = e >> 2
        if b not in cache:
            cache[b] = philox4x64_10([b + 1, 0, 0, 0], key)
        u = (cache[b][e & 3] >> 11) * (1.0 / 9007199254740992.0)
        out[j] = lo[e % d] + rng[e % d] * u
    return out.reshape(N, d)


def _philox_key(key):
    key = np.atleast_1d(np.asarray(key, dtype=np.uint64)).reshape(-1)
    return np.array([key[0], 0], dtype=np.uint64) if key.size == 1 else key


def design_uniform(key, N, bounds, row0=0):
    """Rows [row0, row0+N) of numpy.random.Generator(numpy.random.Philox(key=key)).uniform(lo, hi, (rows, d))."""
    bounds = np.atleast_2d(np.asarray(bounds, dtype=np.float64))
    d = bounds.shape[0]
    e0 = int(row0) * d
    bg = np.random.Philox(key=_philox_key(key), counter=np.array([e0 >> 2, 0, 0, 0], dtype=np.uint64))
    g = np.random.Generator(bg)
    skip = e0 & 3
    u = g.random(skip + N * d)[skip:]
    lo, rng = bounds[:, 0], bounds[:, 1] - bounds[:, 0]
    q = (e0 + np.arange(N * d)) % d
    return (lo[q] + rng[q] * u).reshape(N, d)
