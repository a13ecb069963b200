"""Batch evaluator for Local Penalization whose Lipschitz estimate runs on the mean-only device path
("LP's Lipschitz estimate is itself a K1 sweep").

The reference's ``estimate_L`` (GPyOpt/core/evaluators/batch_local_penalization.py:52-72) asks the inner GPy model for
``predictive_gradients`` -- mean Jacobian AND variance gradient -- at 500 + n sample points and then at every
finite-difference point of a 200-iteration L-BFGS-B run, and throws the variance gradient away.  On the engine the
variance gradient is the dense O(n^2)-per-point contraction; ``gpo_predict_mean_grad`` skips it.  The numbers that reach
SciPy are bit-identical to the ones the reference flow would see (same kernels produce the mean gradient in both
calls), so the estimated constant, and therefore the batch, is the same.  Needs GPyOpt importable.
"""
import numpy as np
import scipy.optimize

from ._compat import HAVE_GPYOPT

if HAVE_GPYOPT:  # pragma: no branch
    from GPyOpt.core.evaluators.batch_local_penalization import LocalPenalization as _ReferenceLP
else:
    _ReferenceLP = object


def _uniform_box(bounds, count):
    """count points uniform in the box, drawn dimension by dimension from NumPy's global stream -- the consumption order
    of samples_multidimensional_uniform (GPyOpt/util/general.py), so a seeded run sees the reference's samples."""
    return np.column_stack([np.random.uniform(lo, hi, count) for lo, hi in bounds])


def estimate_L(gp, bounds, mean_gradient=None):
    """Lipschitz constant of the posterior mean: max over the box of ||d mean / dx||, located from 500 uniform samples plus
    the training inputs and refined by L-BFGS-B, exactly as the reference does it (same sample stream, same optimiser
    call).  ``gp`` is the inner model (``GPModel.model``); ``mean_gradient(X) -> [N, d]`` defaults to its mean-only
    device path ``gp.mean_gradients``."""
    grad = gp.mean_gradients if mean_gradient is None else mean_gradient

    def neg_norm(x):
        return -np.sqrt(np.square(grad(np.atleast_2d(x))).sum(axis=1))

    candidates = np.vstack([_uniform_box(bounds, 500), gp.X])
    start = candidates[np.argmin(neg_norm(candidates))]
    res = scipy.optimize.minimize(neg_norm, start, method='L-BFGS-B', bounds=bounds, options={'maxiter': 200})
    L = -float(np.ravel(res.fun)[0])
    return 10.0 if L < 1e-7 else L      # flat model: the reference's fallback value


class LocalPenalization(_ReferenceLP):
    """Drop-in for ``GPyOpt.core.evaluators.LocalPenalization`` (batch_local_penalization.py:9-49): same sequence of
    penalised optimisations, Lipschitz constant from :func:`estimate_L` above."""

    def __init__(self, acquisition, batch_size):
        if not HAVE_GPYOPT:
            raise ImportError("LocalPenalization builds on GPyOpt's evaluator classes; GPyOpt is not importable")
        super(LocalPenalization, self).__init__(acquisition, batch_size)

    def compute_batch(self, duplicate_manager=None, context_manager=None):
        acq = self.acquisition
        if not hasattr(acq.model.model, 'mean_gradients'):   # not an engine-backed model: reference behaviour
            return super(LocalPenalization, self).compute_batch(duplicate_manager, context_manager)
        acq.update_batches(None, None, None)
        batch = acq.optimize()[0]
        if self.batch_size > 1:
            L = estimate_L(acq.model.model, acq.space.get_bounds())
            y_min = acq.model.model.Y.min()
        for _ in range(1, self.batch_size):
            acq.update_batches(batch, L, y_min)
            batch = np.vstack((batch, acq.optimize()[0]))
        acq.update_batches(None, None, None)
        return batch
