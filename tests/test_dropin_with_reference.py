"""CPU test of the drop-in boundary against the REAL reference BO loop (build container only: needs
/root/reference).  The engine's plugin classes (gpyopt_b200.GPModel / AcquisitionEI / AcquisitionLP ...) are
driven by the oracle-backed FakeEngine, plugged into the unmodified ``GPyOpt.methods.BayesianOptimization``,
and must suggest the same points as the reference's own GPModel + acquisition running on the same GP arithmetic.
This pins the HOST logic of the plugins (shapes, ML-II transforms and restarts, LP wiring, isinstance contracts);
the CUDA arithmetic is pinned by tests/test_gpu_parity.py."""
import os
import sys
import warnings

import numpy as np
import pytest

from conftest import ROOT

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'GPyOpt')), reason="reference tree not present")


@pytest.fixture(scope='module')
def gpyopt():
    warnings.filterwarnings('ignore')
    for p in (os.path.join(ROOT, 'oracle', 'gpy_stub'), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import GPyOpt
    import importlib
    import gpyopt_b200._compat as compat
    importlib.reload(compat)           # pick up the real BOModel / AcquisitionBase now that GPyOpt imports
    import gpyopt_b200.models as models
    import gpyopt_b200.acquisitions as acqs
    importlib.reload(models)
    importlib.reload(acqs)
    import gpyopt_b200.gp as gpmod
    from fake_engine import FakeEngine
    gpmod.Engine = FakeEngine          # no GPU in this container: oracle-backed engine (tests only)
    assert compat.HAVE_GPYOPT
    return GPyOpt, models, acqs


def branin(x):
    from oracle import gpyopt_numpy as og
    return og.branin(x)


DOMAIN = [{'name': 'x1', 'type': 'continuous', 'domain': (-5, 10)}, {'name': 'x2', 'type': 'continuous', 'domain': (1, 15)}]


def _run(GPyOpt, model, acquisition_factory, evaluator_type='sequential', batch_size=1, iters=4, seed=7):
    np.random.seed(seed)
    space = GPyOpt.Design_space(DOMAIN)
    objective = GPyOpt.core.task.SingleObjective(branin)
    aopt = GPyOpt.optimization.AcquisitionOptimizer(space, 'lbfgs')
    acq = acquisition_factory(model, space, aopt)
    X0 = GPyOpt.experiment_design.initial_design('random', space, 5)
    if evaluator_type == 'sequential':
        ev = GPyOpt.core.evaluators.Sequential(acq)
    elif evaluator_type == 'lp_fast':
        import importlib
        import gpyopt_b200.evaluators as evs
        importlib.reload(evs)
        ev = evs.LocalPenalization(acq, batch_size)
    else:
        ev = GPyOpt.core.evaluators.LocalPenalization(acq, batch_size)
    bo = GPyOpt.methods.ModularBayesianOptimization(model, space, objective, acq, ev, X0)
    bo.run_optimization(max_iter=iters, verbosity=False)
    return bo.X.copy(), bo.Y.copy()


def test_plugins_are_accepted_as_reference_types(gpyopt):
    GPyOpt, models, acqs = gpyopt
    from GPyOpt.models.base import BOModel
    from GPyOpt.acquisitions.base import AcquisitionBase
    assert issubclass(models.GPModel, BOModel) and issubclass(models.GPModel_MCMC, BOModel)
    assert issubclass(acqs.AcquisitionEI, AcquisitionBase) and issubclass(acqs.AcquisitionLP, AcquisitionBase)
    m = models.GPModel(max_iters=0)
    bo = GPyOpt.methods.BayesianOptimization(f=branin, domain=DOMAIN, model=m, initial_design_numdata=3)
    assert bo.model is m


@pytest.mark.parametrize('acq_name', ['EI', 'MPI', 'LCB'])
def test_sequential_bo_suggests_the_same_points(gpyopt, acq_name):
    """x_next parity (1e-6) over a short Branin run: reference GPModel+acquisition vs the engine's plugins."""
    GPyOpt, models, acqs = gpyopt
    ref_acq = {'EI': GPyOpt.acquisitions.AcquisitionEI, 'MPI': GPyOpt.acquisitions.AcquisitionMPI,
               'LCB': GPyOpt.acquisitions.AcquisitionLCB}[acq_name]
    new_acq = {'EI': acqs.AcquisitionEI, 'MPI': acqs.AcquisitionMPI, 'LCB': acqs.AcquisitionLCB}[acq_name]
    import GPy
    Xr, Yr = _run(GPyOpt, GPyOpt.models.GPModel(kernel=GPy.kern.RBF(2), exact_feval=True, optimize_restarts=2, verbose=False),
                  lambda m, s, o: ref_acq(m, s, o))
    from gpyopt_b200 import kern
    Xn, Yn = _run(GPyOpt, models.GPModel(kernel=kern.RBF(2), exact_feval=True, optimize_restarts=2, verbose=False),
                  lambda m, s, o: new_acq(m, s, o))
    assert Xr.shape == Xn.shape
    assert np.max(np.abs(Xr - Xn)) < 1e-6, np.max(np.abs(Xr - Xn))


def test_local_penalization_batch_matches(gpyopt):
    """LocalPenalization evaluator (estimate_L reaches through model.model.predictive_gradients / .X / .Y)."""
    GPyOpt, models, acqs = gpyopt
    import GPy
    from gpyopt_b200 import kern

    def ref_factory(m, s, o):
        return GPyOpt.acquisitions.AcquisitionLP(m, s, o, GPyOpt.acquisitions.AcquisitionEI(m, s, o))

    def new_factory(m, s, o):
        return acqs.AcquisitionLP(m, s, o, acqs.AcquisitionEI(m, s, o))

    Xr, _ = _run(GPyOpt, GPyOpt.models.GPModel(kernel=GPy.kern.Matern52(2), optimize_restarts=1, verbose=False), ref_factory,
                 evaluator_type='lp', batch_size=3, iters=2, seed=11)
    Xn, _ = _run(GPyOpt, models.GPModel(kernel=kern.Matern52(2), optimize_restarts=1, verbose=False), new_factory,
                 evaluator_type='lp', batch_size=3, iters=2, seed=11)
    assert Xr.shape == Xn.shape
    assert np.max(np.abs(Xr - Xn)) < 1e-5, np.max(np.abs(Xr - Xn))


def test_fast_lp_evaluator_gives_the_reference_batches(gpyopt):
    """gpyopt_b200.evaluators.LocalPenalization (Lipschitz estimate on the mean-only path) == the reference evaluator
    run on the same model: same estimate_L, same batches, same BO trajectory."""
    GPyOpt, models, acqs = gpyopt
    from gpyopt_b200 import kern
    import importlib
    import gpyopt_b200.evaluators as evs
    importlib.reload(evs)
    from GPyOpt.core.evaluators.batch_local_penalization import estimate_L as ref_estimate_L
    from oracle import gpyopt_numpy as og

    def new_factory(m, s, o):
        return acqs.AcquisitionLP(m, s, o, acqs.AcquisitionEI(m, s, o))

    mk = lambda: models.GPModel(kernel=kern.Matern52(2), optimize_restarts=1, verbose=False)
    Xa, _ = _run(GPyOpt, mk(), new_factory, evaluator_type='lp', batch_size=3, iters=2, seed=11)
    Xb, _ = _run(GPyOpt, mk(), new_factory, evaluator_type='lp_fast', batch_size=3, iters=2, seed=11)
    assert np.array_equal(Xa, Xb)
    # the constant itself, on a fixed model
    rng = np.random.default_rng(3)
    X = np.column_stack([rng.uniform(-5, 10, 14), rng.uniform(1, 15, 14)])
    m = models.GPModel(kernel=kern.RBF(2, lengthscale=3.0), exact_feval=True, max_iters=0, verbose=False)
    m.updateModel(X, og.normalize_stats(branin(X)), None, None)
    bounds = GPyOpt.Design_space(DOMAIN).get_bounds()
    np.random.seed(5)
    L_ref = ref_estimate_L(m.model, bounds)
    np.random.seed(5)
    L_new = evs.estimate_L(m.model, bounds)
    assert L_new == L_ref and L_new > 0
    g = m.model.mean_gradients(X[:3])
    assert g.shape == (3, 2) and np.array_equal(g, m.model.predictive_gradients(X[:3])[0][:, :, 0])


def test_mcmc_model_plugin_shapes(gpyopt):
    GPyOpt, models, acqs = gpyopt
    np.random.seed(0)
    X = np.random.rand(10, 2)
    Y = np.sin(3 * X[:, :1]) + X[:, 1:]
    m = models.GPModel_MCMC(n_samples=3, n_burnin=4, subsample_interval=2, leapfrog_steps=3)
    m.updateModel(X, Y, None, None)
    assert m.hmc_samples.shape == (3, 3)
    means, stds = m.predict(X[:4])
    assert len(means) == 3 and means[0].shape == (4, 1) and stds[0].shape == (4, 1)
    mm, ss, dm, ds = m.predict_withGradients(X[0])
    assert dm[0].shape == (1, 2) and len(m.get_fmin()) == 3
    space = GPyOpt.Design_space(DOMAIN)
    a = acqs.AcquisitionEI_MCMC(m, space, None)
    f, df = a.acquisition_function_withGradients(X[:4])
    assert f.shape == (4, 1) and df.shape == (4, 2)


def test_batched_acquisition_optimizer_is_exact(gpyopt):
    """BatchedAcquisitionOptimizer (5 anchors' L-BFGS runs sharing their device calls through a thread rendezvous, fused
    anchor sweep + top-k) returns exactly the reference AcquisitionOptimizer's optimum, with several times fewer calls."""
    GPyOpt, models, acqs = gpyopt
    import importlib
    import gpyopt_b200.optimization as opt
    importlib.reload(opt)
    from gpyopt_b200 import kern
    from oracle import gpyopt_numpy as og
    rng = np.random.default_rng(5)
    space = GPyOpt.Design_space(DOMAIN)
    X = np.column_stack([rng.uniform(-5, 10, 12), rng.uniform(1, 15, 12)])
    Y = og.normalize_stats(branin(X))
    model = models.GPModel(kernel=kern.RBF(2, lengthscale=3.0), exact_feval=True, max_iters=0, verbose=False)
    model.updateModel(X, Y, None, None)
    calls = {'n': 0}

    for acq_cls in (acqs.AcquisitionEI, acqs.AcquisitionLCB):
        ref_opt = GPyOpt.optimization.AcquisitionOptimizer(space, 'lbfgs')
        acq_ref = acq_cls(model, space, ref_opt)
        orig = model.model.engine.acq

        def counting(*a, **k):
            calls['n'] += 1
            return orig(*a, **k)
        model.model.engine.acq = counting
        np.random.seed(42)
        calls['n'] = 0
        x_ref, f_ref = acq_ref.optimize()
        n_ref = calls['n']
        new_opt = opt.BatchedAcquisitionOptimizer(space, 'lbfgs')
        acq_new = acq_cls(model, space, new_opt)
        new_opt.acquisition = acq_new
        np.random.seed(42)
        calls['n'] = 0
        x_new, f_new = acq_new.optimize()
        n_new = calls['n']
        model.model.engine.acq = orig
        # (bit-identical on the device engine, whose per-row results do not depend on the batch; the NumPy stand-in's
        #  BLAS calls round differently for different batch shapes, hence a tolerance here)
        assert np.allclose(x_ref, x_new, rtol=0, atol=1e-7) and np.allclose(f_ref, f_new, rtol=0, atol=1e-9)
        assert n_new * 2 < n_ref, (n_new, n_ref)
        assert new_opt.last_stats['rows'] >= new_opt.last_stats['device_calls']


def test_device_design_anchor_points(gpyopt):
    """DeviceDesignAnchorPointsGenerator: anchors = the k best rows of the counter-based design, i.e. what the
    reference's ObjectiveAnchorPointsGenerator returns when it is handed that same design; mixed / constrained
    spaces fall back to the host design."""
    GPyOpt, models, acqs = gpyopt
    import importlib
    import gpyopt_b200.optimization as opt
    importlib.reload(opt)
    from gpyopt_b200 import kern
    from oracle import gpyopt_numpy as og
    rng = np.random.default_rng(9)
    space = GPyOpt.Design_space(DOMAIN)
    X = np.column_stack([rng.uniform(-5, 10, 15), rng.uniform(1, 15, 15)])
    Y = og.normalize_stats(branin(X))
    model = models.GPModel(kernel=kern.RBF(2, lengthscale=3.0), exact_feval=True, max_iters=0, verbose=False)
    model.updateModel(X, Y, None, None)
    acq = acqs.AcquisitionEI(model, space, None, jitter=0.01)
    gen = opt.DeviceDesignAnchorPointsGenerator(space, acq, num_samples=2000, seed=31)
    a0 = gen.get(num_anchor=5)
    a1 = gen.get(num_anchor=5)
    assert a0.shape == (5, 2) and not np.array_equal(a0, a1)            # a fresh design per call
    for call, got in ((0, a0), (1, a1)):
        design = og.design_uniform([31, call], 2000, np.array(space.get_bounds(), dtype=float))
        scores = acq.acquisition_function(design).flatten()               # the reference's scoring + argsort[:k]
        assert np.array_equal(got, design[np.argsort(scores, kind='stable')[:5]])
    # optimiser wiring: device_design_seed switches the anchor source, the result is a valid optimum
    o = opt.BatchedAcquisitionOptimizer(space, 'lbfgs', device_design_seed=31)
    acq2 = acqs.AcquisitionEI(model, space, o, jitter=0.01)
    o.acquisition = acq2
    x, fx = acq2.optimize()
    d1000 = og.design_uniform([31, 0], 1000, np.array(space.get_bounds(), dtype=float))   # the optimiser's first design
    assert x.shape == (1, 2) and np.ravel(fx)[0] <= acq2.acquisition_function(d1000).min() + 1e-12
    assert o._device_gen.calls == 1
    # a space with a discrete variable is not a continuous box: host design
    mixed = GPyOpt.Design_space([{'name': 'a', 'type': 'continuous', 'domain': (-5, 10)},
                                 {'name': 'b', 'type': 'discrete', 'domain': (1, 5, 9, 15)}])
    g2 = opt.DeviceDesignAnchorPointsGenerator(mixed, acqs.AcquisitionEI(model, mixed, None), num_samples=200, seed=1)
    assert not g2._continuous_box()
    np.random.seed(0)
    assert g2.get(num_anchor=3).shape == (3, 2) and g2.calls == 0


def test_thompson_anchor_scores_match_the_reference_loop(gpyopt):
    """FusedThompsonAnchorPointsGenerator draws the reference's scores (same global NumPy stream) in one call."""
    GPyOpt, models, acqs = gpyopt
    import importlib
    import gpyopt_b200.optimization as opt
    importlib.reload(opt)
    from gpyopt_b200 import kern
    from oracle import gpyopt_numpy as og
    from GPyOpt.optimization.anchor_points_generator import ThompsonSamplingAnchorPointsGenerator
    rng = np.random.default_rng(4)
    space = GPyOpt.Design_space(DOMAIN)
    X = np.column_stack([rng.uniform(-5, 10, 10), rng.uniform(1, 15, 10)])
    model = models.GPModel(kernel=kern.RBF(2, lengthscale=3.0), exact_feval=True, max_iters=0, verbose=False)
    model.updateModel(X, og.normalize_stats(branin(X)), None, None)
    cand = np.column_stack([rng.uniform(-5, 10, 300), rng.uniform(1, 15, 300)])
    ref = ThompsonSamplingAnchorPointsGenerator(space, 'random', model, num_samples=300)
    new = opt.FusedThompsonAnchorPointsGenerator(space, 'random', model, num_samples=300)
    np.random.seed(8); a = ref.get_anchor_point_scores(cand)
    np.random.seed(8); b = new.get_anchor_point_scores(cand)
    assert np.array_equal(a, b)
    np.random.seed(9); pa = ref.get(num_anchor=5)
    np.random.seed(9); pb = new.get(num_anchor=5)
    assert np.array_equal(pa, pb)


def test_rendezvous_propagates_errors_and_handles_uneven_workers():
    import threading
    from gpyopt_b200.optimization import Rendezvous
    seen = []

    def f(X):
        seen.append(X.shape[0])
        if np.any(X < 0):
            raise ValueError("negative")
        return X.sum(1, keepdims=True), 2 * X

    rdv = Rendezvous({'f': f}, 3)
    out, errs = {}, []

    def work(i, steps):
        try:
            for s in range(steps):
                v, g = rdv.call('f', np.array([i + 1.0, s]))
                out[(i, s)] = (float(v[0, 0]), g[0].copy())
            if i == 2:
                rdv.call('f', np.array([-1.0, 0.0]))
        except ValueError as e:
            errs.append(str(e))
        finally:
            rdv.worker_done()
    ts = [threading.Thread(target=work, args=(i, 2 + i)) for i in range(3)]
    [t.start() for t in ts]
    rdv.serve()
    [t.join() for t in ts]
    assert out[(1, 2)] == (4.0, out[(1, 2)][1]) and np.array_equal(out[(0, 1)][1], [2.0, 2.0])
    assert errs == ['negative'] and seen[0] == 3 and sum(seen) == 2 + 3 + 4 + 1


def test_device_lbfgs_optimizer_wrapper(gpyopt):
    """BatchedAcquisitionOptimizer(local_optimizer='device'): the anchors' local optimisation runs inside the library
    (gpo_lbfgs; here its NumPy mirror in the FakeEngine).  Same anchors as the reference optimiser (same seed), optimum at
    least as good as the reference's SciPy run and equal to it within the reference's own stopping slack; Local
    Penalization goes through the same call; spaces / acquisitions the device optimiser cannot serve fall back."""
    GPyOpt, models, acqs = gpyopt
    import importlib
    import gpyopt_b200.optimization as opt
    importlib.reload(opt)
    from gpyopt_b200 import kern
    from oracle import gpyopt_numpy as og
    from scipy.optimize import fmin_l_bfgs_b
    rng = np.random.default_rng(5)
    space = GPyOpt.Design_space(DOMAIN)
    X = np.column_stack([rng.uniform(-5, 10, 14), rng.uniform(1, 15, 14)])
    Y = og.normalize_stats(branin(X))
    model = models.GPModel(kernel=kern.RBF(2, lengthscale=3.0), exact_feval=True, max_iters=0, verbose=False)
    model.updateModel(X, Y, None, None)
    for acq_cls in (acqs.AcquisitionEI, acqs.AcquisitionLCB):
        acq_ref = acq_cls(model, space, GPyOpt.optimization.AcquisitionOptimizer(space, 'lbfgs'))
        np.random.seed(11)
        x_ref, f_ref = acq_ref.optimize()
        new_opt = opt.BatchedAcquisitionOptimizer(space, 'lbfgs', local_optimizer='device')
        acq_new = acq_cls(model, space, new_opt)
        new_opt.acquisition = acq_new
        np.random.seed(11)
        x_new, f_new = acq_new.optimize()
        assert x_new.shape == (1, 2) and 'lbfgs_iterations' in new_opt.last_stats       # the device path was taken
        assert np.ravel(f_new)[0] <= np.ravel(f_ref)[0] + 1e-9
        assert np.max(np.abs(x_new - x_ref)) < 1e-3 and abs(np.ravel(f_new)[0] - np.ravel(f_ref)[0]) < 1e-6
        # a converged SciPy run started from the device optimum stays put
        fun = lambda z: (lambda r: (float(r[0][0, 0]), r[1][0]))(acq_new.acquisition_function_withGradients(z[None, :]))
        r = fmin_l_bfgs_b(fun, x_new[0], bounds=space.get_bounds(), pgtol=1e-10, factr=10.0)
        assert np.max(np.abs(r[0] - x_new[0])) < 1e-6
    # Local Penalization around EI: penalisers set, device optimiser minimises the penalised value
    inner = acqs.AcquisitionEI(model, space, None, jitter=0.01)
    lp_opt = opt.BatchedAcquisitionOptimizer(space, 'lbfgs', local_optimizer='device')
    lp = acqs.AcquisitionLP(model, space, lp_opt, inner)
    lp_opt.acquisition = lp
    lp.update_batches(x_new, 3.0, float(Y.min()))
    np.random.seed(12)
    x_lp, f_lp = lp.optimize()
    assert 'lbfgs_iterations' in lp_opt.last_stats and np.linalg.norm(x_lp - x_new) > 1e-3    # pushed away from the penaliser
    ref_lp_opt = GPyOpt.optimization.AcquisitionOptimizer(space, 'lbfgs')
    lp_ref = acqs.AcquisitionLP(model, space, ref_lp_opt, inner)
    lp_ref.update_batches(x_new, 3.0, float(Y.min()))
    np.random.seed(12)
    x_lpr, f_lpr = lp_ref.optimize()
    # (the reference's LP "gradient" drops the penalisers' direction vectors -- SURVEY Appendix B.6 -- so it is not the gradient
    #  of the LP value and two line searches do not stop at the same point: agreement to 1e-5 in value)
    assert abs(np.ravel(f_lp)[0] - np.ravel(f_lpr)[0]) < 1e-5 and np.max(np.abs(x_lp - x_lpr)) < 1e-2
    # a mixed space is not served by the device optimiser: the rendezvous path answers
    mixed = GPyOpt.Design_space([{'name': 'a', 'type': 'continuous', 'domain': (-5, 10)},
                                 {'name': 'b', 'type': 'discrete', 'domain': (1, 5, 9, 15)}])
    mo = opt.BatchedAcquisitionOptimizer(mixed, 'lbfgs', local_optimizer='device')
    am = acqs.AcquisitionEI(model, mixed, mo, jitter=0.01)
    mo.acquisition = am
    np.random.seed(13)
    xm, _ = am.optimize()
    assert 'lbfgs_iterations' not in mo.last_stats and xm[0, 1] in (1, 5, 9, 15)
