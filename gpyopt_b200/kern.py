"""Kernel specifications accepted by the engine's models: the counterpart of the ``GPy.kern`` objects that
the reference passes as ``GPModel(kernel=...)`` (GPyOpt/models/gpmodel.py:58,221).  Any object with
``name / input_dim / variance / lengthscale / ARD`` attributes (a real GPy stationary kernel included) is
accepted by :func:`as_spec`.
"""
import numpy as np

_NAMES = {'rbf': 'rbf', 'mat52': 'matern52', 'matern52': 'matern52', 'mat32': 'matern32', 'matern32': 'matern32'}
_GPY_NAME = {'rbf': 'rbf', 'matern52': 'Mat52', 'matern32': 'Mat32'}


class Stationary(object):
    kind = None

    def __init__(self, input_dim, variance=1., lengthscale=None, ARD=False):
        self.input_dim = int(input_dim)
        self.ARD = bool(ARD)
        if lengthscale is None:
            lengthscale = np.ones(self.input_dim if self.ARD else 1)
        lengthscale = np.atleast_1d(np.asarray(lengthscale, dtype=np.float64)).copy()
        if self.ARD and lengthscale.size == 1:
            lengthscale = np.ones(self.input_dim) * lengthscale
        if not self.ARD and lengthscale.size != 1:
            raise ValueError("Only 1 lengthscale needed for non-ARD kernel")
        self.variance = np.array([float(np.asarray(variance).reshape(-1)[0])])
        self.lengthscale = lengthscale

    @property
    def name(self):
        return _GPY_NAME[self.kind]

    def copy(self):
        return self.__class__(self.input_dim, variance=self.variance[0], lengthscale=self.lengthscale.copy(), ARD=self.ARD)

    def __repr__(self):
        return "%s(input_dim=%d, variance=%g, lengthscale=%s, ARD=%s)" % (
            self.__class__.__name__, self.input_dim, self.variance[0], self.lengthscale.tolist(), self.ARD)


class RBF(Stationary):
    kind = 'rbf'


class Matern52(Stationary):
    kind = 'matern52'


class Matern32(Stationary):
    kind = 'matern32'


_CLASSES = {'rbf': RBF, 'matern52': Matern52, 'matern32': Matern32}


def as_spec(kernel, input_dim=None):
    """Normalise a kernel argument (spec, GPy-like kernel object, or name string) into a Stationary spec."""
    if isinstance(kernel, Stationary):
        return kernel.copy()
    if isinstance(kernel, str):
        return _CLASSES[_NAMES[kernel.lower()]](input_dim)
    name = str(getattr(kernel, 'name', '')).lower()
    if name not in _NAMES:
        raise ValueError("unsupported kernel %r: the B200 engine implements RBF, Matern52 and Matern32" % (kernel,))
    return _CLASSES[_NAMES[name]](int(kernel.input_dim), variance=float(np.asarray(kernel.variance).reshape(-1)[0]),
                                  lengthscale=np.asarray(kernel.lengthscale, dtype=float).reshape(-1),
                                  ARD=bool(getattr(kernel, 'ARD', False)))
