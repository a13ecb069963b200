"""CPU tests of the host-side logic that surrounds the CUDA path: kernel specs, parameter transforms,
plugin base-class contracts, candidate sharding across ranks (gloo, world_size 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from gpyopt_b200 import kern
from gpyopt_b200 import gp as gpmod
from gpyopt_b200 import _compat
from gpyopt_b200 import parallel


def test_kernel_specs():
    k = kern.Matern52(3, variance=2.0, ARD=True)
    assert k.name == 'Mat52' and k.lengthscale.shape == (3,) and k.variance[0] == 2.0
    k2 = kern.as_spec(k)
    k2.lengthscale[0] = 9.0
    assert k.lengthscale[0] == 1.0  # copies are independent
    assert kern.as_spec('rbf', 4).kind == 'rbf'

    class FakeGPy(object):  # a GPy-like kernel object
        name, input_dim, variance, lengthscale, ARD = 'rbf', 2, np.array([1.5]), np.array([0.3]), False
    s = kern.as_spec(FakeGPy())
    assert isinstance(s, kern.RBF) and s.variance[0] == 1.5 and s.lengthscale[0] == 0.3
    with pytest.raises(ValueError):
        kern.RBF(2, lengthscale=[1., 2.], ARD=False)

    class Linear(object):
        name, input_dim = 'linear', 2
    with pytest.raises(ValueError):
        kern.as_spec(Linear())


def test_transforms_roundtrip_and_gradfactors():
    pos, bnd = gpmod._Positive(), gpmod._Bounded(1e-9, 1e6)
    for f in (1e-6, 0.3, 5.0, 50.0):
        assert np.isclose(pos.fwd(pos.inv(f)), f, rtol=1e-12)
        h = 1e-6
        num = (pos.fwd(pos.inv(f) + h) - pos.fwd(pos.inv(f) - h)) / (2 * h)
        assert np.isclose(num, pos.gradfactor(f), rtol=1e-5)
    for f in (1e-3, 1.0, 1e3):
        assert np.isclose(bnd.fwd(bnd.inv(f)), f, rtol=1e-9)
        x, h = bnd.inv(f), 1e-6
        assert np.isclose((bnd.fwd(x + h) - bnd.fwd(x - h)) / (2 * h), bnd.gradfactor(f), rtol=1e-4)
    assert bnd.initialize(1e7) == bnd.fwd(0.0) and bnd.initialize(0.5) == 0.5


def test_gamma_prior_from_EV():
    p = gpmod.GammaPrior.from_EV(2., 4.)
    assert p.a == 1.0 and p.b == 0.5
    from scipy.stats import gamma
    assert np.isclose(p.lnpdf(1.3), gamma.logpdf(1.3, a=1.0, scale=2.0))


def test_acquisition_base_contract():
    """acquisition_function = -(f * indicator) / cost; gradient by the quotient rule (base.py:33-52)."""
    class M(object):
        analytical_gradient_prediction = True

    class Space(object):
        model_dimensionality = objective_dimensionality = 2
        def indicator_constraints(self, x):
            return (x[:, :1] > 0).astype(float)

    class A(_compat.AcquisitionBase):
        analytical_gradient_prediction = True
        def _compute_acq(self, x):
            return x.sum(1, keepdims=True)
        def _compute_acq_withGradients(self, x):
            return x.sum(1, keepdims=True), np.ones_like(x)

    cost = lambda x: (2 * np.ones((x.shape[0], 1)), 0.5 * np.ones(x.shape))
    a = A(M(), Space(), None, cost_withGradients=cost)
    x = np.array([[1., 2.], [-1., 2.]])
    assert np.allclose(a.acquisition_function(x), [[-1.5], [0.]])
    f, df = a.acquisition_function_withGradients(x)
    assert np.allclose(f, [[-1.5], [0.]])
    assert np.allclose(df[0], -(1 * 2 - 3 * 0.5) / 4)


def test_shard_rows_partition():
    for N in (0, 1, 7, 10, 1000003):
        for W in (1, 2, 3, 8):
            parts = [parallel.shard_rows(N, r, W) for r in range(W)]
            assert parts[0][0] == 0 and parts[-1][1] == N
            assert all(parts[i][1] == parts[i + 1][0] for i in range(W - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_merge_topk_is_rank_order_independent():
    rng = np.random.default_rng(0)
    scores = rng.standard_normal(1000)
    scores[17] = scores[801]  # a tie across shards
    k = 5
    recs = []
    for r in range(4):
        lo, hi = parallel.shard_rows(1000, r, 4)
        idx = np.argsort(scores[lo:hi], kind='stable')[:k]
        recs.append((scores[lo:hi][idx], idx + lo))
    v, i = parallel.merge_topk(recs, k)
    ref = np.argsort(scores, kind='stable')[:k]
    assert np.array_equal(i, ref) and np.array_equal(v, scores[ref])
    v2, i2 = parallel.merge_topk(recs[::-1], k)
    assert np.array_equal(i2, ref)


@pytest.mark.timeout(180)
def test_two_rank_gloo_sweep_protocol():
    """world_size-2 gloo run of the sharded sweep protocol (broadcast state, shard rows, all-gather top-k)
    with a NumPy stand-in scorer: checks the collective plumbing without a GPU."""
    script = os.path.join(ROOT, 'tests', 'dist_worker.py')
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29631', PYTHONPATH=ROOT)
    res = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                          '--master-addr', '127.0.0.1', '--master-port', '29631', script],
                         capture_output=True, text=True, env=env, timeout=170)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert 'DIST_OK' in res.stdout


def _compat_without_gpyopt():
    """gpyopt_b200._compat as it loads on a machine without GPyOpt (the GPU box): the restated base classes."""
    import importlib.util
    blocked = {k: sys.modules.get(k) for k in list(sys.modules) if k == 'GPyOpt' or k.startswith('GPyOpt.')}
    for k in blocked:
        sys.modules.pop(k)
    sys.modules['GPyOpt'] = None          # import GPyOpt -> ImportError
    try:
        spec = importlib.util.spec_from_file_location('_compat_nogpyopt', os.path.join(ROOT, 'gpyopt_b200', '_compat.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.modules.pop('GPyOpt', None)
        sys.modules.update({k: v for k, v in blocked.items() if v is not None})
    assert not mod.HAVE_GPYOPT
    return mod


@pytest.mark.parametrize('transform', ['none', 'softplus', 'identity'])
def test_fallback_lp_host_formulas_match_the_oracle(transform):
    """The Local-Penalization host formulas used when GPyOpt is absent (gpyopt_b200._compat.LPBase) against the oracle's
    restatement of LP.py, incl. the precompute quirk (sqrt of the std) and a transform string the reference treats as
    the identity."""
    from oracle import gpyopt_numpy as og
    C = _compat_without_gpyopt()
    rng = np.random.default_rng(4)
    d, N, K = 3, 40, 4

    class Model(object):
        analytical_gradient_prediction = True
        def predict(self, x):
            return np.sin(x).sum(1, keepdims=True), 0.3 + np.abs(np.cos(x)).sum(1, keepdims=True) * 1e-3

    class Inner(C.AcquisitionBase):
        analytical_gradient_prediction = True
        def _compute_acq(self, x):
            return 0.5 + np.square(x).sum(1, keepdims=True)
        def _compute_acq_withGradients(self, x):
            return self._compute_acq(x), 2 * x

    model = Model()
    lp = C.LPBase(model, None, None)
    lp.acq, lp.transform, lp.X_batch, lp.r_x0, lp.s_x0 = Inner(model, None, None), transform, None, None, None
    x, Xb = rng.random((N, d)), rng.random((K, d))
    lp.update_batches(Xb, 0.7, -1.2)
    m, s = model.predict(Xb)
    r_o, s_o = og.lp_hammer_precompute(m, s, 0.7, -1.2)
    assert np.array_equal(lp.r_x0, r_o) and np.array_equal(lp.s_x0, s_o)
    af, adf = lp.acq.acquisition_function_withGradients(x)
    assert np.allclose(lp.acquisition_function(x), og.lp_penalized_acquisition(af, x, Xb, r_o, s_o, transform), rtol=1e-13)
    g_o = np.vstack([og.lp_d_acquisition(af[i:i + 1], adf[i:i + 1], x[i:i + 1], Xb, r_o, s_o, transform) for i in range(N)])
    g = np.vstack([lp.d_acquisition_function(x[i:i + 1]) for i in range(N)])
    assert np.allclose(g, g_o, rtol=1e-12, atol=1e-14)
    lp.update_batches(None, None, None)
    assert np.allclose(lp.acquisition_function(x), og.lp_penalized_acquisition(af, x, None, None, None, transform), rtol=1e-13)
