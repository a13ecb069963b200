"""Acquisition plugins: drop-in counterparts of GPyOpt's AcquisitionEI / MPI / LCB (+ _MCMC) and AcquisitionLP
(GPyOpt/acquisitions/{EI,MPI,LCB,EI_mcmc,MPI_mcmc,LCB_mcmc,LP}.py).  ``_compute_acq`` and
``_compute_acq_withGradients`` are ONE fused device sweep (K1 -> DMMA GEMM -> K3) instead of
predict + get_fmin + NumPy; cost weighting and constraints stay in ``AcquisitionBase`` (host, unchanged).
"""
import numpy as np

from ._compat import AcquisitionBase, LPBase
from .models import GPModel, GPModel_MCMC


def _engine_of(model):
    if not isinstance(model, (GPModel, GPModel_MCMC)):
        raise TypeError("the B200 acquisitions need a gpyopt_b200 GPModel / GPModel_MCMC (got %r)" % type(model))
    if model.model is None:
        raise RuntimeError("model has no data yet: call updateModel first")
    if isinstance(model, GPModel_MCMC):
        model._make_resident()
    return model.model.engine


class _FusedAcquisition(AcquisitionBase):
    analytical_gradient_prediction = True
    _kind = None

    def _par(self):
        raise NotImplementedError

    def _compute_acq(self, x):
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        f = _engine_of(self.model).acq(self._kind, self._par(), x, grad=False)
        return f[:, None]

    def _compute_acq_withGradients(self, x):
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        f, df = _engine_of(self.model).acq(self._kind, self._par(), x, grad=True)
        return f[:, None], df


class AcquisitionEI(_FusedAcquisition):
    """Expected improvement (EI.py:6-51)."""
    _kind = 'EI'

    def __init__(self, model, space, optimizer=None, cost_withGradients=None, jitter=0.01):
        self.optimizer = optimizer
        super(AcquisitionEI, self).__init__(model, space, optimizer, cost_withGradients=cost_withGradients)
        self.jitter = jitter

    @staticmethod
    def fromConfig(model, space, optimizer, cost_withGradients, config):
        return AcquisitionEI(model, space, optimizer, cost_withGradients, jitter=config['jitter'])

    def _par(self):
        return self.jitter


class AcquisitionMPI(_FusedAcquisition):
    """Maximum probability of improvement (MPI.py:6-51)."""
    _kind = 'MPI'

    def __init__(self, model, space, optimizer=None, cost_withGradients=None, jitter=0.01):
        self.optimizer = optimizer
        super(AcquisitionMPI, self).__init__(model, space, optimizer, cost_withGradients=cost_withGradients)
        self.jitter = jitter

    @staticmethod
    def fromConfig(model, space, optimizer, cost_withGradients, config):
        return AcquisitionMPI(model, space, optimizer, cost_withGradients, jitter=config['jitter'])

    def _par(self):
        return self.jitter


class AcquisitionLCB(_FusedAcquisition):
    """GP lower confidence bound (LCB.py:6-50); ignores cost like the reference."""
    _kind = 'LCB'

    def __init__(self, model, space, optimizer=None, cost_withGradients=None, exploration_weight=2):
        self.optimizer = optimizer
        super(AcquisitionLCB, self).__init__(model, space, optimizer)
        self.exploration_weight = exploration_weight
        if cost_withGradients is not None:
            print('The set cost function is ignored! LCB acquisition does not make sense with cost.')

    def _par(self):
        return self.exploration_weight


class AcquisitionEI_MCMC(AcquisitionEI):
    """Integrated EI over the HMC samples (EI_mcmc.py:8-59, incl. its +jitter value formula)."""
    _kind = 'EI_MCMC'

    def __init__(self, model, space, optimizer=None, cost_withGradients=None, jitter=0.01):
        super(AcquisitionEI_MCMC, self).__init__(model, space, optimizer, cost_withGradients, jitter)
        assert self.model.MCMC_sampler, 'Samples from the hyper-parameters are needed to compute the integrated EI'


class AcquisitionMPI_MCMC(AcquisitionMPI):
    """Integrated MPI (MPI_mcmc.py:8-59)."""
    _kind = 'MPI_MCMC'

    def __init__(self, model, space, optimizer=None, cost_withGradients=None, jitter=0.01):
        super(AcquisitionMPI_MCMC, self).__init__(model, space, optimizer, cost_withGradients, jitter)
        assert self.model.MCMC_sampler, 'Samples from the hyper-parameters are needed to compute the integrated MPI'


class AcquisitionLCB_MCMC(AcquisitionLCB):
    """Integrated LCB (LCB_mcmc.py:6-52)."""
    _kind = 'LCB_MCMC'

    def __init__(self, model, space, optimizer=None, cost_withGradients=None, exploration_weight=2):
        super(AcquisitionLCB_MCMC, self).__init__(model, space, optimizer, cost_withGradients, exploration_weight)
        assert self.model.MCMC_sampler, 'Samples from the hyper-parameters are needed to compute the integrated GP-LCB'


class AcquisitionLP(LPBase):
    """Local Penalization wrapper (LP.py:10-140) around one of the fused acquisitions above.

    value  = -log g(acq(x)) - sum_k log Phi((|x - x_k| - r_k) / s_k)         (LP.py:70-89)
    grad   = scale * d(-acq)/dx - sum_k d_k  (direction dropped, LP.py:91-132)
    Both are evaluated inside the K3 epilogue on the device.  Like the reference, ``acquisition_function`` returns
    a 1-D array [N], and the gradient is only meaningful row by row (how L-BFGS calls it).
    When the inner acquisition uses a non-constant cost, the space has constraints or the transform is neither 'none' nor
    'softplus', the reference's host formulas (inherited) are applied to the fused inner acquisition instead.
    """

    analytical_gradient_prediction = True

    def __init__(self, model, space, optimizer, acquisition, transform='none'):
        AcquisitionBase.__init__(self, model, space, optimizer)
        if not isinstance(acquisition, _FusedAcquisition):
            raise TypeError("AcquisitionLP (B200) wraps a gpyopt_b200 acquisition")
        self.acq = acquisition
        self.transform = transform.lower()
        if isinstance(acquisition, AcquisitionLCB) and self.transform == 'none':
            self.transform = 'softplus'
        self.X_batch = None
        self.r_x0 = None
        self.s_x0 = None

    def _device_path_ok(self):
        from ._compat import constant_cost_withGradients
        plain_cost = self.acq.cost_withGradients is constant_cost_withGradients
        no_constraints = self.acq.space is None or not getattr(self.acq.space, 'constraints', None)
        same_dim = self.acq.space is None or (self.acq.space.model_dimensionality == self.acq.space.objective_dimensionality)
        # any transform string other than 'none' / 'softplus' is the identity in the reference (LP.py:79-83, 113-121): host path
        known_transform = self.transform in ('none', 'softplus')
        return plain_cost and no_constraints and same_dim and known_transform

    def _run(self, x, grad):
        eng = _engine_of(self.model)
        eng.lp_set(self.X_batch, self.r_x0, self.s_x0, self.transform)
        try:
            return eng.acq(self.acq._kind, self.acq._par(), x, grad=grad)
        finally:
            eng.lp_set(None, None, None, None)

    # update_batches / _hammer_function_precompute and the host formulas are inherited: the reference's own methods when
    # GPyOpt is importable, gpyopt_b200._compat.LPBase otherwise
    def acquisition_function(self, x):
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        if self._device_path_ok():
            return self._run(x, grad=False)
        return super(AcquisitionLP, self).acquisition_function(x)

    def d_acquisition_function(self, x):
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        if self._device_path_ok():
            return self._run(x, grad=True)[1]
        return super(AcquisitionLP, self).d_acquisition_function(x)

    def acquisition_function_withGradients(self, x):
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        if self._device_path_ok():
            return self._run(x, grad=True)
        return super(AcquisitionLP, self).acquisition_function_withGradients(x)
