"""Model plugins: drop-in counterparts of ``GPyOpt.models.GPModel`` / ``GPModel_MCMC``
(GPyOpt/models/gpmodel.py:9-178, 180-355) whose posterior state and arithmetic live on a B200.

Same constructor arguments, method names, shapes and error behaviour as the reference classes; the extra keyword
``device`` selects the GPU.  ``self.model`` is a :class:`gpyopt_b200.gp.GPRegressionB200` (the GPy-duck-typed
inner object that LocalPenalization / plotting reach through).
"""
import numpy as np

from ._compat import BOModel
from . import kern as _kern
from .gp import GPRegressionB200, GammaPrior, HMC


class GPModel(BOModel):
    """Exact GP with ML-II hyper-parameters (reference: gpmodel.py:9-178)."""

    analytical_gradient_prediction = True

    def __init__(self, kernel=None, noise_var=None, exact_feval=False, optimizer='bfgs', max_iters=1000,
                 optimize_restarts=5, sparse=False, num_inducing=10, verbose=True, ARD=False, mean_function=None,
                 device=0, engine=None, adopt_state=False):
        if sparse:
            raise NotImplementedError("sparse GP (VarDTC) is a different posterior and out of the engine's scope")
        if mean_function is not None:
            raise NotImplementedError("mean functions are out of the engine's scope")
        self.kernel = kernel
        self.noise_var = noise_var
        self.exact_feval = exact_feval
        self.optimize_restarts = optimize_restarts
        self.optimizer = optimizer
        self.max_iters = max_iters
        self.verbose = verbose
        self.sparse = sparse
        self.num_inducing = num_inducing
        self.model = None
        self.ARD = ARD
        self.mean_function = mean_function
        self.device = device
        # multi-GPU replicas: `engine` is an Engine whose state was broadcast from the fitting rank
        # (gpyopt_b200.parallel.ShardedSweep.fit); with adopt_state the first updateModel does not refit it
        self._engine = engine
        self._adopt_state = bool(adopt_state)

    @staticmethod
    def fromConfig(config):
        return GPModel(**config)

    def _create_model(self, X, Y):
        """gpmodel.py:50-76: default Matern52(variance=1), noise 0.01 Var(Y), noise fixed 1e-6 or bounded."""
        self.input_dim = X.shape[1]
        if self.kernel is None:
            kern = _kern.Matern52(self.input_dim, variance=1., ARD=self.ARD)
        else:
            kern = self.kernel
            self.kernel = None
        noise_var = Y.var() * 0.01 if self.noise_var is None else self.noise_var
        engine = self.model.engine if self.model is not None else self._engine
        self.model = GPRegressionB200(X, Y, kernel=kern, noise_var=noise_var, device=self.device, engine=engine,
                                      adopt_state=self._adopt_state)
        if self.exact_feval:
            self.model.Gaussian_noise.constrain_fixed(1e-6, warning=False)
        else:
            self.model.Gaussian_noise.constrain_bounded(1e-9, 1e6, warning=False)
        if self._adopt_state:
            self._adopt_state = False
            self.model.unfreeze() if self.max_iters > 0 else None

    def updateModel(self, X_all, Y_all, X_new, Y_new):
        """gpmodel.py:78-93."""
        if self.model is None:
            self._create_model(X_all, Y_all)
        else:
            self.model.set_XY(X_all, Y_all)
        if self.max_iters > 0:
            if self.optimize_restarts == 1:
                self.model.optimize(optimizer=self.optimizer, max_iters=self.max_iters, messages=False, ipython_notebook=False)
            else:
                self.model.optimize_restarts(num_restarts=self.optimize_restarts, optimizer=self.optimizer,
                                             max_iters=self.max_iters, verbose=self.verbose)

    def predict(self, X, with_noise=True):
        """-> (m [N,1], s [N,1]); s = sqrt(clip(var, 1e-10)) (gpmodel.py:95-112)."""
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[None, :]
        m, s = self.model.engine.predict(X, with_noise=with_noise)
        return m[0][:, None], s[0][:, None]

    def _predict(self, X, full_cov, include_likelihood):
        """gpmodel.py:95-100."""
        if X.ndim == 1:
            X = X[None, :]
        m, v = self.model.predict(X, full_cov=full_cov, include_likelihood=include_likelihood)
        v = np.clip(v, 1e-10, np.inf)
        return m, v

    def predict_covariance(self, X, with_noise=True):
        """Full predictive covariance matrix at X (gpmodel.py:114-123)."""
        _, v = self._predict(np.asarray(X, dtype=np.float64), True, with_noise)
        return v

    def get_fmin(self):
        """Minimum posterior mean at the training inputs (gpmodel.py:125-129); cached by the engine per refit."""
        return float(self.model.engine.get_fmin()[0])

    def predict_withGradients(self, X):
        """-> (m [N,1], s [N,1], dmdx [N,d], dsdx [N,d]) (gpmodel.py:131-142)."""
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[None, :]
        m, s, dm, ds = self.model.engine.predict_grad(X)
        return m[0][:, None], s[0][:, None], dm[0], ds[0]

    def copy(self):
        copied = GPModel(kernel=self.model.kern.copy(), noise_var=self.noise_var, exact_feval=self.exact_feval,
                         optimizer=self.optimizer, max_iters=self.max_iters, optimize_restarts=self.optimize_restarts,
                         verbose=self.verbose, ARD=self.ARD, device=self.device)
        copied._create_model(self.model.X, self.model.Y)
        copied.updateModel(self.model.X, self.model.Y, None, None)
        return copied

    def get_model_parameters(self):
        return np.atleast_2d(self.model[:])

    def get_model_parameters_names(self):
        return self.model.parameter_names_flat().tolist()

    def get_covariance_between_points(self, x1, x2):
        """Posterior covariance between two sets of points (gpmodel.py:173-177)."""
        return self.model.posterior_covariance_between_points(x1, x2)


class GPModel_MCMC(BOModel):
    """GP with HMC samples of the hyper-parameters (reference: gpmodel.py:180-355).

    Unlike the reference -- which re-factorises every sample inside every predict / get_fmin call
    (gpmodel.py:265-276) -- the S factorisations are done once per ``updateModel`` and stay resident in HBM.
    """

    MCMC_sampler = True
    analytical_gradient_prediction = True

    def __init__(self, kernel=None, noise_var=None, exact_feval=False, n_samples=10, n_burnin=100, subsample_interval=10,
                 step_size=1e-1, leapfrog_steps=20, verbose=False, device=0):
        self.kernel = kernel
        self.noise_var = noise_var
        self.exact_feval = exact_feval
        self.verbose = verbose
        self.n_samples = n_samples
        self.subsample_interval = subsample_interval
        self.n_burnin = n_burnin
        self.step_size = step_size
        self.leapfrog_steps = leapfrog_steps
        self.model = None
        self.device = device
        self.hmc_samples = None
        self._resident = False

    def _create_model(self, X, Y):
        """gpmodel.py:213-238: default RBF, Gamma(E=2, V=4) priors, noise fixed 1e-6 or positive."""
        self.input_dim = X.shape[1]
        if self.kernel is None:
            kern = _kern.RBF(self.input_dim, variance=1.)
        else:
            kern = self.kernel
            self.kernel = None
        noise_var = Y.var() * 0.01 if self.noise_var is None else self.noise_var
        engine = self.model.engine if self.model is not None else None
        self.model = GPRegressionB200(X, Y, kernel=kern, noise_var=noise_var, device=self.device, engine=engine)
        self.model.kern.set_prior(GammaPrior.from_EV(2., 4.))
        self.model.likelihood.variance.set_prior(GammaPrior.from_EV(2., 4.))
        if self.exact_feval:
            self.model.Gaussian_noise.constrain_fixed(1e-6, warning=False)
        else:
            self.model.Gaussian_noise.constrain_positive(warning=False)
        self._resident = False

    def updateModel(self, X_all, Y_all, X_new, Y_new):
        """gpmodel.py:240-255: optimize(max_iters=200), jitter the parameters, HMC, thin."""
        if self.model is None:
            self._create_model(X_all, Y_all)
        else:
            self.model.set_XY(X_all, Y_all)
        self.model.optimize(max_iters=200)
        self.model.param_array[:] = self.model.param_array * (1. + np.random.randn(self.model.param_array.size) * 0.01)
        self.model._trigger_params_changed()
        self.hmc = HMC(self.model, stepsize=self.step_size)
        ss = self.hmc.sample(num_samples=self.n_burnin + self.n_samples * self.subsample_interval,
                             hmc_iters=self.leapfrog_steps)
        self.set_hmc_samples(ss[self.n_burnin::self.subsample_interval])

    def set_hmc_samples(self, samples):
        """Install hyper-parameter samples (rows = unfixed parameters, as ``model.unfixed_param_array``)."""
        self.hmc_samples = np.atleast_2d(np.asarray(samples, dtype=np.float64)).copy()
        self._resident = False

    def adopt_samples(self, X, Y, rows, engine):
        """Multi-GPU replicas: `engine` already holds the S factorisations for the full parameter rows `rows`
        ([S, p] = variance, lengthscale(s), noise) of this (X, Y) -- replicated from the fitting rank by
        ``gpyopt_b200.parallel.ShardedSweep.fit`` -- so nothing is refitted here."""
        rows = np.atleast_2d(np.asarray(rows, dtype=np.float64))
        self.input_dim = X.shape[1]
        kern = self.kernel if self.kernel is not None else _kern.RBF(self.input_dim, variance=1.)
        self.kernel = None
        self.model = GPRegressionB200(X, Y, kernel=kern, noise_var=float(rows[0, -1]), device=self.device, engine=engine,
                                      adopt_state=True)
        self.model._samples = rows.copy()
        free = np.ones(rows.shape[1], dtype=bool) if self.model._fixes_ is None else self.model._fixes_
        self.hmc_samples = rows[:, free].copy()
        self._resident = True

    def _make_resident(self):
        if self._resident and self.model._samples is not None:
            return
        rows = np.tile(self.model.param_array, (self.hmc_samples.shape[0], 1))
        free = np.ones(self.model.param_array.size, dtype=bool) if self.model._fixes_ is None else self.model._fixes_
        rows[:, free] = self.hmc_samples
        self.model.load_samples(rows)
        self._resident = True

    def predict(self, X):
        """-> (means, stds): lists of S arrays [N,1] (gpmodel.py:257-277)."""
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[None, :]
        self._make_resident()
        m, s = self.model.engine.predict(X, with_noise=True)
        return [mi[:, None] for mi in m], [si[:, None] for si in s]

    def get_fmin(self):
        """-> list of S floats (gpmodel.py:279-295)."""
        self._make_resident()
        return [float(v) for v in self.model.engine.get_fmin()]

    def predict_withGradients(self, X):
        """-> (means, stds, dmdxs, dsdxs): lists of S arrays (gpmodel.py:297-324)."""
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[None, :]
        self._make_resident()
        m, s, dm, ds = self.model.engine.predict_grad(X)
        return ([mi[:, None] for mi in m], [si[:, None] for si in s], [g for g in dm], [g for g in ds])

    def copy(self):
        copied = GPModel_MCMC(kernel=self.model.kern.copy(), noise_var=self.noise_var, exact_feval=self.exact_feval,
                              n_samples=self.n_samples, n_burnin=self.n_burnin, subsample_interval=self.subsample_interval,
                              step_size=self.step_size, leapfrog_steps=self.leapfrog_steps, verbose=self.verbose,
                              device=self.device)
        copied._create_model(self.model.X, self.model.Y)
        copied.updateModel(self.model.X, self.model.Y, None, None)
        return copied

    def get_model_parameters(self):
        return np.atleast_2d(self.model[:])

    def get_model_parameters_names(self):
        return self.model.parameter_names()
