"""NumPy/SciPy float64 restatement of GPy's exact-GP arithmetic.  TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py) -- never imported by the product package.

The hot path's arithmetic lives in the third-party dependency GPy (``GPy>=1.8`` un-pinned,
/root/reference/setup.py:26, requirements.txt:7; paramz>=0.7.0 requirements.txt:6).  GPy is
not vendored under /root/reference and cannot be installed here, so this module restates the
published GPy 1.9.x algorithm, anchored on the reference's call sites:

  GPyOpt/models/gpmodel.py:58,67      kern.Matern52 / models.GPRegression construction
  GPyOpt/models/gpmodel.py:73,76      Gaussian_noise.constrain_fixed / constrain_bounded
  GPyOpt/models/gpmodel.py:85,91,93   set_XY / optimize / optimize_restarts
  GPyOpt/models/gpmodel.py:98,129,136 predict
  GPyOpt/models/gpmodel.py:138        predictive_gradients
  GPyOpt/models/gpmodel.py:165,171    model[:] / parameter_names_flat
  GPyOpt/models/gpmodel.py:177        posterior_covariance_between_points
  GPyOpt/models/gpmodel.py:221-232    kern.RBF, priors.Gamma.from_EV, constrain_positive
  GPyOpt/models/gpmodel.py:251-255    optimize(max_iters=200), inference.mcmc.HMC(...).sample
  GPyOpt/models/gpmodel.py:267-276    model[:]=s / model[_fixes_]=s / _trigger_params_changed
  GPyOpt/core/evaluators/batch_local_penalization.py:37-38,58,65  .predictive_gradients/.X/.Y

GPy symbols restated (names are GPy's): kern.src.stationary.Stationary{K,_scaled_dist,
_unscaled_dist,K_of_r,dK_dr,gradients_X,_inv_dist,update_gradients_full}, kern.src.rbf.RBF,
kern.src.stationary.{Matern32,Matern52}, inference.latent_function_inference.
exact_gaussian_inference.ExactGaussianInference.inference, ...posterior.PosteriorExact.
_raw_predict, core.gp.GP.{predict,predictive_gradients,posterior_covariance_between_points,
set_XY}, likelihoods.gaussian.Gaussian.predictive_values, util.linalg.{jitchol,pdinv,dpotrs,
dpotri,dtrtrs}, paramz transformations Logexp/Logistic, paramz Model.{optimize,
optimize_restarts,randomize}, core.parameterization.priors.Gamma, inference.mcmc.hmc.HMC.

Pin: GPyOpt/testing/core_tests/test_cost.py:35-42 (see tests/test_oracle_pins.py).
"""
import numpy as np
import scipy.linalg as sla
from scipy import optimize as _sopt

_LOG_2_PI = np.log(2.0 * np.pi)
_lim_val = 36.0
_log_lim_val = np.log(np.finfo(np.float64).max)


# ----------------------------------------------------------------------------------------
# GPy.util.linalg
# ----------------------------------------------------------------------------------------
def jitchol(A, maxtries=5):
    """GPy.util.linalg.jitchol: lower Cholesky with escalating diagonal jitter."""
    A = np.ascontiguousarray(A)
    try:
        return sla.cholesky(A, lower=True, check_finite=False)
    except sla.LinAlgError:
        pass
    diagA = np.diag(A)
    if np.any(diagA <= 0.0):
        raise np.linalg.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    num_tries = 1
    while num_tries <= maxtries and np.isfinite(jitter):
        try:
            return sla.cholesky(A + np.eye(A.shape[0]) * jitter, lower=True, check_finite=False)
        except sla.LinAlgError:
            jitter *= 10
        num_tries += 1
    raise np.linalg.LinAlgError("not positive definite, even with jitter.")


def dtrtrs(A, B, lower=1, trans=0):
    return sla.solve_triangular(A, B, lower=bool(lower), trans=trans, check_finite=False)


def dpotrs(L, B):
    return sla.cho_solve((L, True), B, check_finite=False)


def dpotri(L):
    """LAPACK dpotri on the lower factor, symmetrised (GPy.util.linalg.dpotri + symmetrify)."""
    Ai, info = sla.lapack.dpotri(L, lower=1)
    if info != 0:
        raise np.linalg.LinAlgError("dpotri failed")
    Ai = np.tril(Ai)
    return Ai + np.tril(Ai, -1).T


def pdinv(A):
    """GPy.util.linalg.pdinv -> (Ai, L, Li, logdet)."""
    L = jitchol(A)
    logdet = 2.0 * np.sum(np.log(np.diag(L)))
    Li = sla.lapack.dtrtri(L, lower=1)[0]
    Ai = dpotri(L)
    return Ai, L, Li, logdet


# ----------------------------------------------------------------------------------------
# paramz transformations
# ----------------------------------------------------------------------------------------
class Logexp(object):
    def f(self, x):
        return np.where(x > _lim_val, x, np.log1p(np.exp(np.clip(x, -_log_lim_val, _lim_val))))

    def finv(self, f):
        return np.where(f > _lim_val, f, np.log(np.expm1(f)))

    def gradfactor(self, f, df):
        return df * np.where(f > _lim_val, 1.0, -np.expm1(-f))

    def initialize(self, f):
        return np.abs(f)

    def log_jacobian(self, p):
        return np.where(p > _lim_val, p, np.log(np.expm1(p))) - p

    def log_jacobian_grad(self, p):
        return 1.0 / np.expm1(p)


class Logistic(object):
    def __init__(self, lower, upper):
        self.lower, self.upper = float(lower), float(upper)
        self.difference = self.upper - self.lower

    def f(self, x):
        x = np.where(x < -300.0, -300.0, x)
        return self.lower + self.difference / (1.0 + np.exp(-x))

    def finv(self, f):
        return np.log(np.clip(f - self.lower, 1e-10, np.inf) / np.clip(self.upper - f, 1e-10, np.inf))

    def gradfactor(self, f, df):
        return df * (f - self.lower) * (self.upper - f) / self.difference

    def initialize(self, f):
        f = np.asarray(f, dtype=float)
        bad = np.logical_or(f < self.lower, f > self.upper)
        return np.where(bad, self.f(f * 0.0), f)

    def log_jacobian(self, p):
        return np.log((p - self.lower) * (self.upper - p) / self.difference) * 0.0  # unused

    def log_jacobian_grad(self, p):
        return 0.0 * p


class Gamma(object):
    """GPy.core.parameterization.priors.Gamma (shape a, rate b)."""

    def __init__(self, a, b):
        self.a, self.b = float(a), float(b)
        from scipy.special import gammaln
        self.constant = -gammaln(self.a) + self.a * np.log(self.b)

    @staticmethod
    def from_EV(E, V):
        return Gamma(np.square(E) / V, E / V)

    def lnpdf(self, x):
        return self.constant + (self.a - 1.0) * np.log(x) - self.b * x

    def lnpdf_grad(self, x):
        return (self.a - 1.0) / x - self.b


# ----------------------------------------------------------------------------------------
# GPy.kern.src.stationary
# ----------------------------------------------------------------------------------------
class Stationary(object):
    """Stationary kernel, parameters [variance, lengthscale(1 or input_dim)]."""
    name = 'stationary'

    def __init__(self, input_dim, variance=1.0, lengthscale=None, ARD=False, name=None):
        self.input_dim = int(input_dim)
        self.ARD = bool(ARD)
        if lengthscale is None:
            lengthscale = np.ones(self.input_dim if self.ARD else 1)
        lengthscale = np.atleast_1d(np.asarray(lengthscale, dtype=np.float64))
        if self.ARD and lengthscale.size == 1:
            lengthscale = np.ones(self.input_dim) * lengthscale
        if not self.ARD:
            assert lengthscale.size == 1, "Only 1 lengthscale needed for non-ARD kernel"
        self._p = np.concatenate([[float(variance)], lengthscale]).astype(np.float64)
        self._prior = None
        self._model = None
        if name is not None:
            self.name = name

    # parameter views (re-bound into the model's flat array by GPRegression)
    @property
    def variance(self):
        return self._p[0:1]

    @property
    def lengthscale(self):
        return self._p[1:]

    @property
    def size(self):
        return self._p.size

    def copy(self):
        k = self.__class__(self.input_dim, variance=float(self._p[0]), lengthscale=self._p[1:].copy(),
                           ARD=self.ARD)
        k._prior = self._prior
        return k

    def set_prior(self, prior, warning=True):
        self._prior = prior
        if self._model is not None:
            self._model._priors['variance'] = prior
            self._model._priors['lengthscale'] = prior

    def parameter_names(self):
        return ['variance', 'lengthscale']

    # -- distances --------------------------------------------------------
    def _unscaled_dist(self, X, X2=None):
        if X2 is None:
            Xsq = np.sum(np.square(X), 1)
            r2 = -2.0 * np.dot(X, X.T) + (Xsq[:, None] + Xsq[None, :])
            r2[np.diag_indices_from(r2)] = 0.0
            r2 = np.clip(r2, 0, np.inf)
            return np.sqrt(r2)
        X1sq = np.sum(np.square(X), 1)
        X2sq = np.sum(np.square(X2), 1)
        r2 = -2.0 * np.dot(X, X2.T) + (X1sq[:, None] + X2sq[None, :])
        r2 = np.clip(r2, 0, np.inf)
        return np.sqrt(r2)

    def _scaled_dist(self, X, X2=None):
        if self.ARD:
            if X2 is not None:
                X2 = X2 / self.lengthscale
            return self._unscaled_dist(X / self.lengthscale, X2)
        return self._unscaled_dist(X, X2) / self.lengthscale

    def _inv_dist(self, X, X2=None):
        dist = self._scaled_dist(X, X2).copy()
        return 1.0 / np.where(dist != 0.0, dist, np.inf)

    def K(self, X, X2=None):
        return self.K_of_r(self._scaled_dist(X, X2))

    def Kdiag(self, X):
        ret = np.empty(X.shape[0])
        ret[:] = self.variance
        return ret

    def dK_dr_via_X(self, X, X2):
        return self.dK_dr(self._scaled_dist(X, X2))

    def gradients_X(self, dL_dK, X, X2=None):
        invdist = self._inv_dist(X, X2)
        dL_dr = self.dK_dr_via_X(X, X2) * dL_dK
        tmp = invdist * dL_dr
        if X2 is None:
            tmp = tmp + tmp.T
            X2 = X
        grad = np.empty(X.shape, dtype=np.float64)
        for q in range(self.input_dim):
            np.sum(tmp * (X[:, q][:, None] - X2[:, q][None, :]), axis=1, out=grad[:, q])
        return grad / self.lengthscale ** 2

    def gradients_X_diag(self, dL_dKdiag, X):
        return np.zeros(X.shape)

    def param_gradients(self, dL_dK, X, X2=None):
        """Stationary.update_gradients_full -> d/d[variance, lengthscale...]."""
        r = self._scaled_dist(X, X2)
        K = self.K_of_r(r)
        g_var = np.sum(K * dL_dK) / self.variance[0]
        dL_dr = self.dK_dr(r) * dL_dK
        if self.ARD:
            tmp = dL_dr * self._inv_dist(X, X2)
            if X2 is None:
                X2 = X
            g_l = np.array([-np.sum(tmp * np.square(X[:, q:q + 1] - X2[:, q:q + 1].T)) / self.lengthscale[q] ** 3
                            for q in range(self.input_dim)])
        else:
            g_l = np.array([-np.sum(dL_dr * r) / self.lengthscale[0]])
        return np.concatenate([[g_var], g_l])


class RBF(Stationary):
    name = 'rbf'

    def K_of_r(self, r):
        return self.variance * np.exp(-0.5 * r ** 2)

    def dK_dr(self, r):
        return -r * self.K_of_r(r)


class Matern52(Stationary):
    name = 'Mat52'

    def K_of_r(self, r):
        return self.variance * (1 + np.sqrt(5.) * r + 5. / 3 * r ** 2) * np.exp(-np.sqrt(5.) * r)

    def dK_dr(self, r):
        return self.variance * (10. / 3 * r - 5. * r - 5. * np.sqrt(5.) / 3 * r ** 2) * np.exp(-np.sqrt(5.) * r)


class Matern32(Stationary):
    name = 'Mat32'

    def K_of_r(self, r):
        return self.variance * (1. + np.sqrt(3.) * r) * np.exp(-np.sqrt(3.) * r)

    def dK_dr(self, r):
        return -3. * self.variance * r * np.exp(-np.sqrt(3.) * r)


# ----------------------------------------------------------------------------------------
# likelihood + parameter handles
# ----------------------------------------------------------------------------------------
class _ParamHandle(object):
    """A view on a slice of the model's flat parameter array with paramz-like constrain_* API."""

    def __init__(self, owner, name):
        self._owner, self._name = owner, name

    @property
    def values(self):
        return self._owner._get(self._name)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __float__(self):
        return float(self.values[0])

    def __getitem__(self, i):
        return self.values[i]

    def set_prior(self, prior, warning=True):
        self._owner._set_prior(self._name, prior)

    def constrain_fixed(self, value=None, warning=True):
        self._owner._constrain(self._name, 'fixed', value)

    def constrain_bounded(self, lower, upper, warning=True):
        self._owner._constrain(self._name, Logistic(lower, upper), None)

    def constrain_positive(self, warning=True):
        self._owner._constrain(self._name, Logexp(), None)


class Gaussian(object):
    """GPy.likelihoods.Gaussian; ``variance`` is the noise variance."""
    name = 'Gaussian_noise'

    def __init__(self, variance=1.0):
        self._p = np.array([float(variance)])
        self._model = None

    @property
    def variance(self):
        if self._model is not None:
            return _ParamHandle(self._model, 'noise')
        return self._p

    def _noise(self):
        return float(self._p[0])

    # the reference calls these on model.Gaussian_noise directly (gpmodel.py:73,76,232,234)
    def constrain_fixed(self, value=None, warning=True):
        self._model._constrain('noise', 'fixed', value)

    def constrain_bounded(self, lower, upper, warning=True):
        self._model._constrain('noise', Logistic(lower, upper), None)

    def constrain_positive(self, warning=True):
        self._model._constrain('noise', Logexp(), None)

    def predictive_values(self, mu, var, full_cov=False):
        if full_cov:
            return mu, var + np.eye(var.shape[0]) * self._noise()
        return mu, var + self._noise()


class _Posterior(object):
    """PosteriorExact: woodbury_chol = L, woodbury_vector = alpha, lazy woodbury_inv (A.3)."""

    def __init__(self, woodbury_chol, woodbury_vector, K):
        self.woodbury_chol = woodbury_chol
        self.woodbury_vector = woodbury_vector
        self.K = K
        self._woodbury_inv = None

    @property
    def woodbury_inv(self):
        if self._woodbury_inv is None:
            self._woodbury_inv = dpotri(self.woodbury_chol)
        return self._woodbury_inv


class _OptRun(object):
    def __init__(self, x_opt, f_opt):
        self.x_opt, self.f_opt = x_opt, f_opt


# ----------------------------------------------------------------------------------------
# GPy.models.GPRegression (core.gp.GP + ExactGaussianInference)
# ----------------------------------------------------------------------------------------
class GPRegression(object):
    name = 'GP_regression'

    def __init__(self, X, Y, kernel=None, Y_metadata=None, normalizer=None, noise_var=1.0, mean_function=None):
        assert mean_function is None, "oracle: mean functions are out of scope"
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        assert X.ndim == 2 and Y.ndim == 2
        if kernel is None:
            kernel = RBF(X.shape[1])
        self.X, self.Y = X.copy(), Y.copy()
        self.kern = kernel
        self.likelihood = Gaussian(variance=noise_var)
        self.Gaussian_noise = self.likelihood
        self.input_dim = X.shape[1]
        self.output_dim = Y.shape[1]
        # flat parameter array; rebind kernel + likelihood storage as views into it
        nk = kernel._p.size
        self.param_array = np.concatenate([kernel._p, self.likelihood._p]).astype(np.float64)
        kernel._p = self.param_array[:nk]
        kernel._model = self
        self.likelihood._p = self.param_array[nk:nk + 1]
        self.likelihood._model = self
        self._slices = {'variance': slice(0, 1), 'lengthscale': slice(1, nk), 'noise': slice(nk, nk + 1)}
        self._transforms = {}  # index -> transformation object
        for i in range(self.param_array.size):
            self._transforms[i] = Logexp()
        self._fixes_ = None
        self._priors = {}  # name -> prior
        if kernel._prior is not None:
            self._priors['variance'] = kernel._prior
            self._priors['lengthscale'] = kernel._prior
        self.optimization_runs = []
        self._trigger_params_changed()

    # -- parameter plumbing --------------------------------------------------------------
    def _get(self, name):
        return self.param_array[self._slices[name]]

    def _set_prior(self, name, prior):
        self._priors[name] = prior

    def _constrain(self, name, what, value):
        sl = self._slices[name]
        idx = range(*sl.indices(self.param_array.size))
        if isinstance(what, str) and what == 'fixed':
            if value is not None:
                self.param_array[sl] = value
            if self._fixes_ is None:
                self._fixes_ = np.ones(self.param_array.size, dtype=bool)
            self._fixes_[sl] = False
        else:
            self.param_array[sl] = what.initialize(self.param_array[sl])
            for i in idx:
                self._transforms[i] = what
        self._trigger_params_changed()

    def _free_mask(self):
        if self._fixes_ is None:
            return np.ones(self.param_array.size, dtype=bool)
        return self._fixes_

    def __getitem__(self, key):
        return self.param_array[key].copy()

    def __setitem__(self, key, value):
        self.param_array[key] = value
        self._trigger_params_changed()

    @property
    def unfixed_param_array(self):
        return self.param_array[self._free_mask()].copy()

    @property
    def optimizer_array(self):
        free = np.flatnonzero(self._free_mask())
        return np.array([float(self._transforms[i].finv(self.param_array[i])) for i in free])

    @optimizer_array.setter
    def optimizer_array(self, x):
        free = np.flatnonzero(self._free_mask())
        for xi, i in zip(np.asarray(x, dtype=float), free):
            self.param_array[i] = float(self._transforms[i].f(np.asarray(xi)))
        self._trigger_params_changed()

    def parameter_names(self):
        return ['%s.variance' % self.kern.name, '%s.lengthscale' % self.kern.name, 'Gaussian_noise.variance']

    def parameter_names_flat(self):
        names = ['%s.%s.variance' % (self.name, self.kern.name)]
        nl = self.kern.lengthscale.size
        if nl == 1:
            names.append('%s.%s.lengthscale' % (self.name, self.kern.name))
        else:
            names += ['%s.%s.lengthscale[[%d]]' % (self.name, self.kern.name, q) for q in range(nl)]
        names.append('%s.Gaussian_noise.variance' % self.name)
        return np.array(names)[self._free_mask()]

    # -- inference ----------------------------------------------------------
    def _trigger_params_changed(self):
        self.parameters_changed()

    def parameters_changed(self):
        K = self.kern.K(self.X)
        Ky = K.copy()
        Ky[np.diag_indices_from(Ky)] += self.likelihood._noise() + 1e-8
        Wi, LW, LWi, W_logdet = pdinv(Ky)
        alpha = dpotrs(LW, self.Y)
        self._log_marginal_likelihood = 0.5 * (-self.Y.size * _LOG_2_PI - self.Y.shape[1] * W_logdet
                                               - np.sum(alpha * self.Y))
        self.posterior = _Posterior(LW, alpha, K)
        self.posterior._woodbury_inv = Wi
        dL_dK = 0.5 * (np.dot(alpha, alpha.T) - self.Y.shape[1] * Wi)
        self._dL_dK = dL_dK
        g_kern = self.kern.param_gradients(dL_dK, self.X)
        g_noise = np.trace(dL_dK)
        self.gradient = np.concatenate([g_kern, [g_noise]])

    def set_XY(self, X=None, Y=None):
        if Y is not None:
            self.Y = np.asarray(Y, dtype=np.float64).copy()
        if X is not None:
            self.X = np.asarray(X, dtype=np.float64).copy()
        self._trigger_params_changed()

    def log_likelihood(self):
        return self._log_marginal_likelihood

    def log_prior(self):
        if not self._priors:
            return 0.0
        lp = 0.0
        for name, prior in self._priors.items():
            sl = self._slices[name]
            x = self.param_array[sl]
            lp += np.sum(prior.lnpdf(x))
            for i in range(*sl.indices(self.param_array.size)):
                lp += float(self._transforms[i].log_jacobian(self.param_array[i]))
        return lp

    def _log_prior_gradients(self):
        g = np.zeros(self.param_array.size)
        for name, prior in self._priors.items():
            sl = self._slices[name]
            x = self.param_array[sl]
            g[sl] += prior.lnpdf_grad(x)
            for i in range(*sl.indices(self.param_array.size)):
                g[i] += float(self._transforms[i].log_jacobian_grad(self.param_array[i]))
        return g

    def objective_function(self):
        return -float(self.log_likelihood()) - self.log_prior()

    def objective_function_gradients(self):
        return -(self.gradient + self._log_prior_gradients())

    def _transform_gradients(self, g):
        free = np.flatnonzero(self._free_mask())
        return np.array([float(self._transforms[i].gradfactor(self.param_array[i], g[i])) for i in free])

    def _objective_grads(self, x):
        try:
            self.optimizer_array = x
            obj = self.objective_function()
            grads = self._transform_gradients(self.objective_function_gradients())
            grads = np.clip(grads, -1e10, 1e10)
            if not np.isfinite(obj):
                return 1e300, np.zeros_like(grads)
            return obj, grads
        except (np.linalg.LinAlgError, ZeroDivisionError, ValueError, FloatingPointError):
            return 1e300, np.zeros(int(self._free_mask().sum()))

    def optimize(self, optimizer=None, start=None, messages=False, max_iters=1000, ipython_notebook=False,
                 clear_after_finish=False, **kwargs):
        """paramz Model.optimize with the 'lbfgsb' optimiser (both 'bfgs' and 'lbfgs' resolve to it)."""
        x0 = self.optimizer_array.copy() if start is None else np.asarray(start, dtype=float)
        if x0.size == 0:
            return
        res = _sopt.fmin_l_bfgs_b(self._objective_grads, x0, maxfun=max_iters, maxiter=max_iters)
        x_opt = res[0]
        f_opt = self._objective_grads(x_opt)[0]
        self.optimization_runs.append(_OptRun(x_opt, f_opt))
        self.optimizer_array = x_opt

    def randomize(self):
        x = np.random.normal(loc=0, scale=1, size=int(self._free_mask().sum()))
        self.optimizer_array = x
        # (paramz re-draws priored parameters from their prior here; GPyOpt only calls
        #  optimize_restarts on the prior-free GPModel, gpmodel.py:93, so that branch is omitted.)

    def optimize_restarts(self, num_restarts=10, robust=False, verbose=True, parallel=False, num_processes=None,
                          **kwargs):
        initial_length = len(self.optimization_runs)
        initial_parameters = self.optimizer_array.copy()
        for i in range(num_restarts):
            try:
                if i > 0:
                    self.randomize()
                self.optimize(**kwargs)
            except Exception:
                if robust:
                    continue
                raise
        if len(self.optimization_runs) > initial_length:
            i = int(np.argmin([o.f_opt for o in self.optimization_runs[initial_length:]]))
            self.optimizer_array = self.optimization_runs[initial_length + i].x_opt
        else:
            self.optimizer_array = initial_parameters

    # -- prediction -----------------------------------------------------
    def _raw_predict(self, Xnew, full_cov=False):
        Kx = self.kern.K(self.X, Xnew)
        mu = np.dot(Kx.T, self.posterior.woodbury_vector)
        tmp = dtrtrs(self.posterior.woodbury_chol, Kx)
        if full_cov:
            var = self.kern.K(Xnew) - np.dot(tmp.T, tmp)
        else:
            var = (self.kern.Kdiag(Xnew) - np.square(tmp).sum(0))[:, None]
        return mu, var

    def predict(self, Xnew, full_cov=False, Y_metadata=None, kern=None, likelihood=None, include_likelihood=True):
        Xnew = np.asarray(Xnew, dtype=np.float64)
        mu, var = self._raw_predict(Xnew, full_cov=full_cov)
        if include_likelihood:
            mu, var = self.likelihood.predictive_values(mu, var, full_cov)
        return mu, var

    def predictive_gradients(self, Xnew, kern=None):
        Xnew = np.asarray(Xnew, dtype=np.float64)
        mean_jac = np.empty((Xnew.shape[0], Xnew.shape[1], self.output_dim))
        for i in range(self.output_dim):
            mean_jac[:, :, i] = self.kern.gradients_X(self.posterior.woodbury_vector[:, i:i + 1].T, Xnew, self.X)
        dv_dX = self.kern.gradients_X_diag(np.ones(Xnew.shape[0]), Xnew)
        alpha = -2.0 * np.dot(self.kern.K(Xnew, self.X), self.posterior.woodbury_inv)
        dv_dX = dv_dX + self.kern.gradients_X(alpha, Xnew, self.X)
        return mean_jac, dv_dX

    def posterior_covariance_between_points(self, X1, X2):
        L = self.posterior.woodbury_chol
        t1 = dtrtrs(L, self.kern.K(self.X, X1))
        t2 = dtrtrs(L, self.kern.K(self.X, X2))
        return self.kern.K(X1, X2) - np.dot(t1.T, t2)


# ----------------------------------------------------------------------------------------
# GPy.inference.mcmc.HMC
# ----------------------------------------------------------------------------------------
class HMC(object):
    def __init__(self, model, M=None, stepsize=1e-1):
        self.model = model
        self.stepsize = stepsize
        self.p = np.empty_like(model.optimizer_array.copy())
        self.M = np.eye(self.p.size) if M is None else M
        self.Minv = np.linalg.inv(self.M)

    def sample(self, num_samples=1000, hmc_iters=20):
        params = np.empty((num_samples, self.p.size))
        for i in range(num_samples):
            self.p[:] = np.random.multivariate_normal(np.zeros(self.p.size), self.M)
            H_old = self._computeH()
            theta_old = self.model.optimizer_array.copy()
            params[i] = self.model.unfixed_param_array
            self._update(hmc_iters)
            H_new = self._computeH()
            k = 1.0 if H_old > H_new else np.exp(H_old - H_new)
            if np.random.rand() < k:
                params[i] = self.model.unfixed_param_array
            else:
                self.model.optimizer_array = theta_old
        return params

    def _update(self, hmc_iters):
        for _ in range(hmc_iters):
            self.p[:] += -self.stepsize / 2. * self.model._transform_gradients(self.model.objective_function_gradients())
            self.model.optimizer_array = self.model.optimizer_array + self.stepsize * np.dot(self.Minv, self.p)
            self.p[:] += -self.stepsize / 2. * self.model._transform_gradients(self.model.objective_function_gradients())

    def _computeH(self):
        return (self.model.objective_function() + self.p.size * np.log(2 * np.pi) / 2.
                + np.log(np.linalg.det(self.M)) / 2. + float(np.dot(self.p, np.dot(self.Minv, self.p[:, None]))[0]) / 2.)
