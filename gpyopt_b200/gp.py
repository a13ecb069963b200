"""Device-resident exact GP regression object: what ``GPModel.model`` is in the reference
(a ``GPy.models.GPRegression``, GPyOpt/models/gpmodel.py:67), re-hosted on the B200 engine.

It is duck-type compatible with the parts of GPy's model API that GPyOpt reaches through
(SURVEY.md section 8b): ``.X .Y .predict .predictive_gradients [:] param_array _fixes_
_trigger_params_changed set_XY optimize optimize_restarts parameter_names_flat kern``.
All arithmetic -- Gram matrix, Cholesky, inverse, alpha, log-likelihood and its gradient, prediction -- runs in
the CUDA library behind ``gpyopt_b200._lib.Engine``; this file holds only the host-side control flow GPy/paramz
provide around it (parameter transforms, L-BFGS-B ML-II with restarts, HMC).
"""
import numpy as np
from scipy import optimize as _sopt
from scipy.special import gammaln as _gammaln

from ._lib import Engine
from .kern import as_spec

_LIM = 36.0
_LOG_LIM = np.log(np.finfo(np.float64).max)
_LOG_2PI = np.log(2.0 * np.pi)


# -- paramz transformations (Logexp for positive, Logistic for bounded parameters) -------------------
class _Positive(object):
    def fwd(self, x):
        return x if x > _LIM else np.log1p(np.exp(np.clip(x, -_LOG_LIM, _LIM)))

    def inv(self, f):
        return f if f > _LIM else np.log(np.expm1(f))

    def gradfactor(self, f):
        return 1.0 if f > _LIM else -np.expm1(-f)

    def initialize(self, f):
        return abs(f)

    def log_jacobian(self, f):
        return (f if f > _LIM else np.log(np.expm1(f))) - f

    def log_jacobian_grad(self, f):
        return 1.0 / np.expm1(f)


class _Bounded(object):
    def __init__(self, lower, upper):
        self.lower, self.upper = float(lower), float(upper)
        self.span = self.upper - self.lower

    def fwd(self, x):
        return self.lower + self.span / (1.0 + np.exp(-max(x, -300.0)))

    def inv(self, f):
        return np.log(max(f - self.lower, 1e-10) / max(self.upper - f, 1e-10))

    def gradfactor(self, f):
        return (f - self.lower) * (self.upper - f) / self.span

    def initialize(self, f):
        return self.fwd(0.0) if (f < self.lower or f > self.upper) else f

    def log_jacobian(self, f):
        return 0.0

    def log_jacobian_grad(self, f):
        return 0.0


class GammaPrior(object):
    """Gamma(a, b) prior on a positive hyper-parameter (GPy.priors.Gamma; from_EV as gpmodel.py:231-232)."""

    def __init__(self, a, b):
        self.a, self.b = float(a), float(b)
        self.const = -_gammaln(self.a) + self.a * np.log(self.b)

    @staticmethod
    def from_EV(E, V):
        return GammaPrior(E * E / V, E / V)

    def lnpdf(self, x):
        return self.const + (self.a - 1.0) * np.log(x) - self.b * x

    def lnpdf_grad(self, x):
        return (self.a - 1.0) / x - self.b


class _NoiseHandle(object):
    """``model.Gaussian_noise`` / ``model.likelihood``: constrain_* calls used by gpmodel.py:73,76,232,234."""

    def __init__(self, gp):
        self._gp = gp

    @property
    def variance(self):
        return self

    def __float__(self):
        return float(self._gp.param_array[-1])

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._gp.param_array[-1:], dtype=dtype)

    def constrain_fixed(self, value=None, warning=True):
        self._gp._constrain_noise('fixed', value)

    def constrain_bounded(self, lower, upper, warning=True):
        self._gp._constrain_noise(_Bounded(lower, upper), None)

    def constrain_positive(self, warning=True):
        self._gp._constrain_noise(_Positive(), None)

    def set_prior(self, prior, warning=True):
        self._gp._priors[self._gp.param_array.size - 1] = prior


class _KernView(object):
    """``model.kern``: live view of the kernel hyper-parameters held in the model's flat parameter array."""

    def __init__(self, gp):
        self._gp = gp

    @property
    def name(self):
        return self._gp._spec.name

    @property
    def input_dim(self):
        return self._gp.input_dim

    @property
    def ARD(self):
        return self._gp._spec.ARD

    @property
    def variance(self):
        return self._gp.param_array[0:1]

    @property
    def lengthscale(self):
        return self._gp.param_array[1:-1]

    def copy(self):
        s = self._gp._spec.copy()
        s.variance = self.variance.copy()
        s.lengthscale = self.lengthscale.copy()
        return s

    def set_prior(self, prior, warning=True):
        for i in range(self._gp.param_array.size - 1):
            self._gp._priors[i] = prior


class _OptRun(object):
    def __init__(self, x_opt, f_opt):
        self.x_opt, self.f_opt = x_opt, f_opt


class GPRegressionB200(object):
    """Exact GP regression with a stationary kernel and Gaussian noise; posterior state lives on the GPU.

    Parameter order (as GPy): ``[kern.variance, kern.lengthscale (1 or d), Gaussian_noise.variance]``.
    """
    name = 'GP_regression'

    def __init__(self, X, Y, kernel=None, noise_var=1.0, device=0, engine=None, adopt_state=False):
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        Y = np.ascontiguousarray(np.asarray(Y, dtype=np.float64))
        if X.ndim != 2 or Y.ndim != 2 or Y.shape[1] != 1 or Y.shape[0] != X.shape[0]:
            raise ValueError("X must be [n, d] and Y [n, 1]")
        self.X, self.Y = X.copy(), Y.copy()
        self.input_dim, self.output_dim = X.shape[1], 1
        self._spec = as_spec(kernel if kernel is not None else 'rbf', self.input_dim)
        if self._spec.input_dim != self.input_dim:
            raise ValueError("kernel input_dim %d != data dimension %d" % (self._spec.input_dim, self.input_dim))
        self.engine = engine if engine is not None else Engine(device)
        self.param_array = np.concatenate([self._spec.variance, self._spec.lengthscale, [float(noise_var)]])
        self._transforms = [_Positive() for _ in range(self.param_array.size)]
        self._fixes_ = None
        self._priors = {}
        self.kern = _KernView(self)
        self.likelihood = _NoiseHandle(self)
        self.Gaussian_noise = self.likelihood
        self.optimization_runs = []
        self._samples = None  # [S, p] full parameter rows when several hyper-parameter sets are resident
        # adopt_state: `engine` already holds the posterior for exactly these (X, Y, hyper-parameters) -- e.g. it was
        # replicated from another rank over NCCL (gpyopt_b200.parallel) -- so the constructor must not refit
        self._frozen = bool(adopt_state)
        self._trigger_params_changed()

    # -- parameter plumbing ---------------------------------------------------------------------------
    def _constrain_noise(self, what, value):
        i = self.param_array.size - 1
        if isinstance(what, str):  # fixed
            if value is not None:
                self.param_array[i] = value
            if self._fixes_ is None:
                self._fixes_ = np.ones(self.param_array.size, dtype=bool)
            self._fixes_[i] = False
        else:
            self.param_array[i] = what.initialize(self.param_array[i])
            self._transforms[i] = what
        self._trigger_params_changed()

    def _free(self):
        return np.ones(self.param_array.size, dtype=bool) if self._fixes_ is None else self._fixes_

    def __getitem__(self, key):
        return self.param_array[key].copy()

    def __setitem__(self, key, value):
        self.param_array[key] = value
        self._trigger_params_changed()

    @property
    def unfixed_param_array(self):
        return self.param_array[self._free()].copy()

    @property
    def optimizer_array(self):
        return np.array([self._transforms[i].inv(self.param_array[i]) for i in np.flatnonzero(self._free())])

    @optimizer_array.setter
    def optimizer_array(self, x):
        for xi, i in zip(np.asarray(x, dtype=float), np.flatnonzero(self._free())):
            self.param_array[i] = self._transforms[i].fwd(float(xi))
        self._trigger_params_changed()

    def parameter_names(self):
        k = self.kern.name
        return ['%s.variance' % k, '%s.lengthscale' % k, 'Gaussian_noise.variance']

    def parameter_names_flat(self):
        k, nl = self.kern.name, self.param_array.size - 2
        names = ['%s.%s.variance' % (self.name, k)]
        names += (['%s.%s.lengthscale' % (self.name, k)] if nl == 1 else
                  ['%s.%s.lengthscale[[%d]]' % (self.name, k, q) for q in range(nl)])
        names.append('%s.Gaussian_noise.variance' % self.name)
        return np.array(names)[self._free()]

    # -- refit on the device ----------------------------------------------------------------------------
    def _rows_to_theta(self, rows):
        rows = np.atleast_2d(rows)
        return rows[:, 0], rows[:, 1:-1], rows[:, -1]

    def _trigger_params_changed(self):
        """GPy ``parameters_changed``: Ky = K + (noise + 1e-8) I -> L, alpha, W^-1 (K4 on the GPU)."""
        self._samples = None
        if self._frozen:
            return
        var, ls, noise = self._rows_to_theta(self.param_array[None, :])
        self.engine.fit(self._spec.kind, self.X, self.Y[:, 0], var, ls, noise)

    parameters_changed = _trigger_params_changed

    def load_samples(self, rows):
        """Make S hyper-parameter sets resident at once (the GPModel_MCMC prediction state)."""
        rows = np.atleast_2d(np.asarray(rows, dtype=np.float64))
        var, ls, noise = self._rows_to_theta(rows)
        self.engine.fit(self._spec.kind, self.X, self.Y[:, 0], var, ls, noise)
        self._samples = rows.copy()

    def unfreeze(self):
        """Leave adopt_state mode: later parameter / data changes refit on this engine again."""
        self._frozen = False

    def set_XY(self, X=None, Y=None):
        self._frozen = False
        if X is not None:
            self.X = np.ascontiguousarray(np.asarray(X, dtype=np.float64)).copy()
        if Y is not None:
            self.Y = np.ascontiguousarray(np.asarray(Y, dtype=np.float64)).copy()
        self._trigger_params_changed()

    def _ensure_own_state(self):
        """The engine may be holding a set of hyper-parameter samples (GPModel_MCMC's prediction state).  GPy's model
        always answers for its CURRENT parameters (the reference restores them after every sample loop,
        gpmodel.py:275-276), so inner-model calls first put the single current set back."""
        if self._samples is not None:
            frozen, self._frozen = self._frozen, False
            self._trigger_params_changed()
            self._frozen = frozen

    # -- marginal likelihood ------------------------------------------------------------------------------
    def log_likelihood(self):
        self._ensure_own_state()
        return float(self.engine.get_loglik()[0][0])

    def _loglik_gradient(self):
        self._ensure_own_state()
        g = self.engine.get_loglik_grad()[0]
        d = self.input_dim
        gl = g[1:1 + d] if self._spec.ARD else np.array([g[1:1 + d].sum()])
        return np.concatenate([[g[0]], gl, [g[d + 1]]])

    def log_prior(self):
        lp = 0.0
        for i, prior in self._priors.items():
            lp += float(prior.lnpdf(self.param_array[i])) + self._transforms[i].log_jacobian(self.param_array[i])
        return lp

    def _log_prior_gradient(self):
        g = np.zeros(self.param_array.size)
        for i, prior in self._priors.items():
            g[i] += prior.lnpdf_grad(self.param_array[i]) + self._transforms[i].log_jacobian_grad(self.param_array[i])
        return g

    def objective_function(self):
        return -self.log_likelihood() - self.log_prior()

    def objective_function_gradients(self):
        return -(self._loglik_gradient() + self._log_prior_gradient())

    def _transform_gradients(self, g):
        return np.array([g[i] * self._transforms[i].gradfactor(self.param_array[i]) for i in np.flatnonzero(self._free())])

    def _objective_grads(self, x):
        try:
            self.optimizer_array = x
            obj = self.objective_function()
            grads = np.clip(self._transform_gradients(self.objective_function_gradients()), -1e10, 1e10)
            if not np.isfinite(obj):
                return 1e300, np.zeros_like(grads)
            return obj, grads
        except (np.linalg.LinAlgError, ZeroDivisionError, ValueError, FloatingPointError):
            return 1e300, np.zeros(int(self._free().sum()))

    def optimize(self, optimizer=None, start=None, messages=False, max_iters=1000, ipython_notebook=False, **kwargs):
        """ML-II by L-BFGS-B on the transformed parameters (paramz resolves both 'bfgs' and 'lbfgs' to it)."""
        x0 = self.optimizer_array.copy() if start is None else np.asarray(start, dtype=float)
        if x0.size == 0:
            return
        res = _sopt.fmin_l_bfgs_b(self._objective_grads, x0, maxfun=max_iters, maxiter=max_iters)
        f_opt = self._objective_grads(res[0])[0]
        self.optimization_runs.append(_OptRun(res[0], f_opt))
        self.optimizer_array = res[0]

    def randomize(self):
        self.optimizer_array = np.random.normal(loc=0, scale=1, size=int(self._free().sum()))

    def optimize_restarts(self, num_restarts=10, robust=False, verbose=True, parallel=False, num_processes=None, **kwargs):
        start_len = len(self.optimization_runs)
        x_init = self.optimizer_array.copy()
        for i in range(num_restarts):
            try:
                if i > 0:
                    self.randomize()
                self.optimize(**kwargs)
                if verbose:
                    print("Optimization restart %d/%d, f = %s" % (i + 1, num_restarts, self.optimization_runs[-1].f_opt))
            except Exception:
                if not robust:
                    raise
        if len(self.optimization_runs) > start_len:
            best = int(np.argmin([o.f_opt for o in self.optimization_runs[start_len:]]))
            self.optimizer_array = self.optimization_runs[start_len + best].x_opt
        else:
            self.optimizer_array = x_init

    # -- prediction (GPy signatures; variances, not standard deviations) -------------------------------------
    def predict(self, Xnew, full_cov=False, Y_metadata=None, kern=None, likelihood=None, include_likelihood=True):
        Xnew = np.atleast_2d(np.asarray(Xnew, dtype=np.float64))
        self._ensure_own_state()
        if full_cov:
            m, _ = self.engine.predict(Xnew, with_noise=include_likelihood)
            return m[0][:, None], self.engine.posterior_cov(Xnew, None, with_noise=include_likelihood)
        m, s = self.engine.predict(Xnew, with_noise=include_likelihood)
        return m[0][:, None], np.square(s[0])[:, None]

    def posterior_covariance_between_points(self, X1, X2):
        self._ensure_own_state()
        return self.engine.posterior_cov(np.atleast_2d(X1), np.atleast_2d(X2))

    def predictive_gradients(self, Xnew, kern=None):
        Xnew = np.atleast_2d(np.asarray(Xnew, dtype=np.float64))
        self._ensure_own_state()
        m, s, dm, ds = self.engine.predict_grad(Xnew)
        return dm[0][:, :, None], ds[0] * (2.0 * s[0][:, None])


    def mean_gradients(self, Xnew):
        """d mean / dx [N, d] alone (the first output of predictive_gradients) without the variance contraction."""
        Xnew = np.atleast_2d(np.asarray(Xnew, dtype=np.float64))
        self._ensure_own_state()
        return self.engine.predict_mean_grad(Xnew)[1][0]


class HMC(object):
    """Hybrid Monte Carlo over the (transformed) hyper-parameters: GPy.inference.mcmc.HMC as used by
    GPModel_MCMC.updateModel (gpmodel.py:253-255).  Every leapfrog gradient is one device refit."""

    def __init__(self, model, M=None, stepsize=1e-1):
        self.model, self.stepsize = model, stepsize
        self.p = np.empty_like(model.optimizer_array.copy())
        self.M = np.eye(self.p.size) if M is None else M
        self.Minv = np.linalg.inv(self.M)

    def sample(self, num_samples=1000, hmc_iters=20):
        params = np.empty((num_samples, self.p.size))
        for i in range(num_samples):
            self.p[:] = np.random.multivariate_normal(np.zeros(self.p.size), self.M)
            H_old = self._H()
            theta_old = self.model.optimizer_array.copy()
            params[i] = self.model.unfixed_param_array
            self._leapfrog(hmc_iters)
            H_new = self._H()
            k = 1.0 if H_old > H_new else np.exp(H_old - H_new)
            if np.random.rand() < k:
                params[i] = self.model.unfixed_param_array
            else:
                self.model.optimizer_array = theta_old
        return params

    def _grad(self):
        return self.model._transform_gradients(self.model.objective_function_gradients())

    def _leapfrog(self, hmc_iters):
        for _ in range(hmc_iters):
            self.p[:] += -self.stepsize / 2. * self._grad()
            self.model.optimizer_array = self.model.optimizer_array + self.stepsize * np.dot(self.Minv, self.p)
            self.p[:] += -self.stepsize / 2. * self._grad()

    def _H(self):
        return (self.model.objective_function() + self.p.size * _LOG_2PI / 2. + np.log(np.linalg.det(self.M)) / 2.
                + float(np.dot(self.p, np.dot(self.Minv, self.p))) / 2.)
