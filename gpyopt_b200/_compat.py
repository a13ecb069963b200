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

    LPBase = AcquisitionBase
