"""Stub ``GPy`` package: lets the UNMODIFIED reference /root/reference/GPyOpt import and run in
the build container, with GPy's exact-GP arithmetic supplied by oracle/gpy_numpy.py.
TEST INFRASTRUCTURE ONLY (golden-vector generation and oracle pinning); see oracle/__init__.py.

Only the attributes GPyOpt dereferences on the hot path exist (gpmodel.py:58,67,221-232,253).
"""
import os as _os
import sys as _sys
import types as _types

_here = _os.path.dirname(_os.path.abspath(__file__))
_root = _os.path.abspath(_os.path.join(_here, '..', '..', '..'))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from oracle import gpy_numpy as _g  # noqa: E402

__version__ = '1.9.9-oracle-stub'

kern = _types.ModuleType('GPy.kern')
kern.RBF = _g.RBF
kern.Matern52 = _g.Matern52
kern.Matern32 = _g.Matern32
kern.Stationary = _g.Stationary

models = _types.ModuleType('GPy.models')
models.GPRegression = _g.GPRegression


def _unsupported(*a, **k):
    raise NotImplementedError("oracle GPy stub: out of the hot-path scope (SURVEY.md section 2)")


models.SparseGPRegression = _unsupported
models.WarpedGP = _unsupported
models.InputWarpedGP = _unsupported
models.GradientChecker = _unsupported

priors = _types.ModuleType('GPy.priors')
priors.Gamma = _g.Gamma

inference = _types.ModuleType('GPy.inference')
inference.mcmc = _types.ModuleType('GPy.inference.mcmc')
inference.mcmc.HMC = _g.HMC

util = _types.ModuleType('GPy.util')
util.linalg = _types.ModuleType('GPy.util.linalg')
for _n in ('jitchol', 'pdinv', 'dpotrs', 'dpotri', 'dtrtrs'):
    setattr(util.linalg, _n, getattr(_g, _n))

for _m in (kern, models, priors, inference, inference.mcmc, util, util.linalg):
    _sys.modules[_m.__name__] = _m
