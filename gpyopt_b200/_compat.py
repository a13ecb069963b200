"""Base classes of the two plugin boundaries.

If the reference package is importable (``import GPyOpt``), the engine's model / acquisition classes derive
from ITS ``BOModel`` / ``AcquisitionBase`` so that ``BayesianOptimization(model=..., acquisition=...)``'s
``isinstance`` checks pass (GPyOpt/methods/bayesian_optimization.py:126-130,146-150).  Otherwise local
stand-ins with the same interface are used (GPyOpt/models/base.py:7-33, GPyOpt/acquisitions/base.py:5-70).
"""
import abc

import numpy as np

try:  # pragma: no cover - depends on the environment
    from GPyOpt.models.base import BOModel
    from GPyOpt.acquisitions.base import AcquisitionBase
    from GPyOpt.core.task.cost import constant_cost_withGradients
    from GPyOpt.acquisitions.LP import AcquisitionLP as LPBase  # LocalPenalization asserts isinstance of this class
    HAVE_GPYOPT = True
except Exception:  # GPyOpt (or its GPy dependency) absent: same contracts, restated
    HAVE_GPYOPT = False

    class BOModel(abc.ABC):
        """The model plugin boundary (mirrors GPyOpt/models/base.py:7-33)."""
        MCMC_sampler = False
        analytical_gradient_prediction = False

        @abc.abstractmethod
        def updateModel(self, X_all, Y_all, X_new, Y_new):
            return

        @abc.abstractmethod
        def predict(self, X):
            return

        def predict_withGradients(self, X):
            return

        @abc.abstractmethod
        def get_fmin(self):
            return

    def constant_cost_withGradients(x):
        """cost = 1, d_cost = 0 (GPyOpt/core/task/cost.py:76-80)."""
        return np.ones(x.shape[0])[:, None], np.zeros(x.shape)

    class AcquisitionBase(object):
        """The acquisition plugin boundary (mirrors GPyOpt/acquisitions/base.py:5-70)."""
        analytical_gradient_prediction = False

        def __init__(self, model, space, optimizer, cost_withGradients=None):
            self.model = model
            self.space = space
            self.optimizer = optimizer
            self.analytical_gradient_acq = self.analytical_gradient_prediction and self.model.analytical_gradient_prediction
            self.cost_withGradients = constant_cost_withGradients if cost_withGradients is None else cost_withGradients

        def _indicator(self, x):
            if self.space is None:
                return 1.0
            x_z = x if self.space.model_dimensionality == self.space.objective_dimensionality else self.space.zip_inputs(x)
            return self.space.indicator_constraints(x_z)

        def acquisition_function(self, x):
            f_acqu = self._compute_acq(x)
            cost_x, _ = self.cost_withGradients(x)
            return -(f_acqu * self._indicator(x)) / cost_x

        def acquisition_function_withGradients(self, x):
            f_acqu, df_acqu = self._compute_acq_withGradients(x)
            cost_x, cost_grad_x = self.cost_withGradients(x)
            f_acq_cost = f_acqu / cost_x
            df_acq_cost = (df_acqu * cost_x - f_acqu * cost_grad_x) / (cost_x ** 2)
            ind = self._indicator(x)
            return -f_acq_cost * ind, -df_acq_cost * ind

        def optimize(self, duplicate_manager=None):
            if not self.analytical_gradient_acq:
                return self.optimizer.optimize(f=self.acquisition_function, duplicate_manager=duplicate_manager)
            return self.optimizer.optimize(f=self.acquisition_function, f_df=self.acquisition_function_withGradients,
                                           duplicate_manager=duplicate_manager)

        def _compute_acq(self, x):
            raise NotImplementedError('')

        def _compute_acq_withGradients(self, x):
            raise NotImplementedError('')

    class LPBase(AcquisitionBase):
        """Host side of Local Penalization for an environment without GPyOpt: the contract of
        GPyOpt/acquisitions/LP.py:40-140 (penaliser parameters, log-space penalised value, gradient with the direction
        vector dropped).  With GPyOpt importable the B200 class inherits the reference's own methods instead."""

        def update_batches(self, X_batch, L, Min):
            self.X_batch = X_batch
            if X_batch is not None:
                self.r_x0, self.s_x0 = self._hammer_function_precompute(X_batch, L, Min, self.model)

        def _hammer_function_precompute(self, x0, L, Min, model):
            if x0 is None:
                return None, None
            x0 = np.atleast_2d(x0)
            mean, spread = model.predict(x0)
            spread = np.sqrt(np.maximum(spread, 1e-16))     # sqrt of the returned std: the reference's quirk (LP.py:55-59)
            return ((mean - Min) / L).ravel(), (spread / L).ravel()

        def _z_and_norm(self, x):
            nm = np.linalg.norm(np.atleast_2d(x)[:, None, :] - np.atleast_2d(self.X_batch)[None, :, :], axis=-1)
            return (nm - self.r_x0) / self.s_x0, nm

        def _inner(self, x):
            return -self.acq.acquisition_function(x)[:, 0]

        def acquisition_function(self, x):
            from scipy.stats import norm
            g = self._inner(x)
            if self.transform == 'softplus':
                big = g >= 40.
                g = np.where(big, np.log(np.where(big, g, 1.)), np.log(np.log1p(np.exp(np.where(big, 0., g)))))
            elif self.transform == 'none':
                g = np.log(g + 1e-50)
            out = -g
            if self.X_batch is not None:
                out = out - norm.logcdf(self._z_and_norm(x)[0]).sum(axis=-1)
            return out

        def d_acquisition_function(self, x):
            from scipy.stats import norm
            x = np.atleast_2d(x)
            scale = 1.
            if self.transform == 'softplus':
                g = self._inner(x)
                scale = 1. / (np.log1p(np.exp(g)) * (1. + np.exp(-g)))
            elif self.transform == 'none':
                scale = 1. / self._inner(x)
            grad = scale * self.acq.acquisition_function_withGradients(x)[1]
            if self.X_batch is None:
                return grad
            z, nm = self._z_and_norm(x)
            cdf = norm.cdf(z)
            with np.errstate(divide='ignore', invalid='ignore'):
                pen = np.exp(-0.5 * z * z) / (self.s_x0 * np.sqrt(2. * np.pi) * cdf * nm)
            pen[cdf < 1e-50] = 0.
            return grad - pen.sum(axis=1)[:, None]

        def acquisition_function_withGradients(self, x):
            return self.acquisition_function(x), self.d_acquisition_function(x)
