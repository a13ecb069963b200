"""GPU tests of the reference-facing plugin classes (gpyopt_b200.GPModel / GPModel_MCMC / Acquisition*):
same shapes and values as the reference classes produced for the golden vectors, ML-II pin, x_next parity."""
import numpy as np
import pytest

from conftest import load_golden, scaled_err, assert_grad_close, parity_tolerances, oracle_gp_from_golden, near_dup_budgets

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _model_from_golden(g, **kw):
    from gpyopt_b200 import GPModel, kern
    cls = {'rbf': kern.RBF, 'matern52': kern.Matern52, 'matern32': kern.Matern32}[str(g['kernel'])]
    d = g['X'].shape[1]
    k = cls(d, variance=float(g['variance']), lengthscale=g['lengthscale'], ARD=bool(g['ARD']))
    m = GPModel(kernel=k, noise_var=float(g['noise_var']), exact_feval=bool(g['exact_feval']), max_iters=0, verbose=False, **kw)
    m.updateModel(g['X'], g['Y'], None, None)
    return m


@pytest.mark.parametrize('case', ['c3_rbf_alpine2_n256', 'c3_mat32_ard_n130', 'c1_branin_rbf_n20', 'c1_1d_rbf_n7'])
def test_gpmodel_plugin_matches_reference_class(case):
    from gpyopt_b200 import AcquisitionEI, AcquisitionMPI, AcquisitionLCB
    g = load_golden(case)
    model = _model_from_golden(g)
    Xs = g['Xs']
    gp = oracle_gp_from_golden(g)
    TOL, TOLG = parity_tolerances(gp, Xs)
    nd_dm, nd_ds = near_dup_budgets(gp, g)       # the terms the reference drops on the near-duplicate row (conftest.py)
    m, s = model.predict(Xs)
    assert m.shape == g['m'].shape and s.shape == g['s'].shape
    assert scaled_err(m, g['m']) < TOL and scaled_err(s, g['s']) < TOL
    assert scaled_err(model.predict(Xs[5])[0], g['m_1d']) < TOL          # 1-D input is promoted to one row
    m, s, dm, ds = model.predict_withGradients(Xs)
    assert dm.shape == g['dmdx'].shape
    assert_grad_close(dm, g['dmdx'], tol=TOLG, near_dup_abs=nd_dm)
    assert_grad_close(ds, g['dsdx'], tol=TOLG, near_dup_abs=nd_ds)
    assert abs(model.get_fmin() - float(g['fmin'])) < TOL * max(1, abs(float(g['fmin'])))
    assert np.allclose(model.get_model_parameters(), g['params'], rtol=1e-14)
    # the inner GPy-like object that LocalPenalization / plotting reach through
    mi, vi = model.model.predict(Xs)
    assert scaled_err(np.sqrt(vi), g['s']) < TOL
    dmi, dvi = model.model.predictive_gradients(Xs)
    assert dmi.shape == (Xs.shape[0], Xs.shape[1], 1)
    assert_grad_close(dvi / (2 * g['s']), g['dsdx'], tol=TOLG, near_dup_abs=nd_ds)
    for name, cls, kw in (('EI', AcquisitionEI, {'jitter': 0.01}), ('MPI', AcquisitionMPI, {'jitter': 0.01}),
                          ('LCB', AcquisitionLCB, {'exploration_weight': 2})):
        a = cls(model, None, None, **kw)
        af = a.acquisition_function(Xs)
        assert af.shape == g[name + '_af'].shape and scaled_err(af, g[name + '_af']) < TOL
        afg, adf = a.acquisition_function_withGradients(Xs)
        assert scaled_err(afg, g[name + '_afg']) < TOL
        m1, s1 = float(g['mg'][1, 0]), float(g['sg'][1, 0])
        u1 = (float(g['fmin']) - m1 - 0.01) / s1
        phi1 = np.exp(-0.5 * u1 * u1) / np.sqrt(2 * np.pi)
        A, B = {'EI': (1.0, phi1), 'MPI': (phi1 / s1, phi1 / s1 * abs(u1)), 'LCB': (1.0, 2.0)}[name]
        assert_grad_close(adf, g[name + '_adf'], tol=TOLG, near_dup_abs=A * nd_dm + B * nd_ds)


def test_full_covariance_methods_match_oracle():
    """GPModel.predict_covariance / get_covariance_between_points (gpmodel.py:114-123, 173-177): the Entropy-Search-only
    corner of the model API, served from the same K1 / GEMM primitives."""
    g = load_golden('c2_camel_mat52_n55')
    model = _model_from_golden(g)
    gp = oracle_gp_from_golden(g)
    X1, X2 = g['Xs'][:150], g['Xs'][150:190]
    for with_noise in (True, False):
        ref = np.clip(gp.predict(X1, full_cov=True, include_likelihood=with_noise)[1], 1e-10, np.inf)
        cov = model.predict_covariance(X1, with_noise=with_noise)
        assert cov.shape == ref.shape and scaled_err(cov, ref) < TOL
    ref = gp.posterior_covariance_between_points(X1, X2)
    cov = model.get_covariance_between_points(X1, X2)
    assert cov.shape == (150, 40) and scaled_err(cov, ref) < TOL
    # a larger model (n = 256 -> 2 row blocks) and more points than one tile
    g = load_golden('c3_rbf_alpine2_n256')
    model, gp = _model_from_golden(g), oracle_gp_from_golden(g)
    ref = gp.posterior_covariance_between_points(g['Xs'][:300], g['Xs'][100:384])
    assert scaled_err(model.get_covariance_between_points(g['Xs'][:300], g['Xs'][100:384]), ref) < TOL


def test_mcmc_plugin_matches_reference_class():
    from gpyopt_b200 import GPModel_MCMC, AcquisitionEI_MCMC, AcquisitionLCB_MCMC
    for case in ('c4_mcmc_n64_S8', 'c4_mcmc_exact_n48_S4'):
        g = load_golden(case)
        model = GPModel_MCMC(exact_feval=bool(g['exact_feval']))
        model._create_model(g['X'], g['Y'])
        model.set_hmc_samples(g['hmc_samples'])
        means, stds, dms, dss = model.predict_withGradients(g['Xs'])
        assert isinstance(means, list) and len(means) == g['hmc_samples'].shape[0] and means[0].shape == (g['Xs'].shape[0], 1)
        tg = 1e-10 if not bool(g['exact_feval']) else 5e-9   # noise fixed at 1e-6: see parity_tolerances
        assert scaled_err(np.array(means), g['means']) < TOL and scaled_err(np.array(dss), g['dsdxs']) < tg
        assert scaled_err(np.array(model.get_fmin()), g['fmins']) < TOL
        f, df = AcquisitionEI_MCMC(model, None, None, jitter=0.01)._compute_acq_withGradients(g['Xs'])
        assert scaled_err(f, g['EI_MCMC_fg']) < TOL and scaled_err(df, g['EI_MCMC_df']) < tg
        f = AcquisitionLCB_MCMC(model, None, None, exploration_weight=2)._compute_acq(g['Xs'])
        assert scaled_err(f, g['LCB_MCMC_f']) < TOL
        # the inner GPy-like object answers for the model's CURRENT parameters (the reference restores them after each
        # sample loop), not for the first resident sample ...
        from oracle import gpy_numpy as G
        p = model.model.param_array
        gp = G.GPRegression(g['X'], g['Y'], G.RBF(g['X'].shape[1], variance=p[0], lengthscale=p[1]), noise_var=p[2])
        mi, vi = model.model.predict(g['Xs'][:7])
        mo, vo = gp.predict(g['Xs'][:7])
        assert scaled_err(mi, mo) < 1e-9 and scaled_err(vi, np.clip(vo, 1e-10, np.inf)) < 1e-9
        # ... and the sample state comes back for the next integrated prediction
        means2, _ = model.predict(g['Xs'])
        assert scaled_err(np.array(means2), g['means']) < TOL


def test_lp_plugin_matches_reference_class():
    from gpyopt_b200 import AcquisitionEI, AcquisitionLCB, AcquisitionLP
    g, gg = load_golden('c5_lp_gsobol_n128'), load_golden('c5_lp_gsobol_n128_gp')
    model = _model_from_golden(gg)
    Xs = g['Xs']
    for tname, inner, transform in (('EI_none', AcquisitionEI(model, None, None, jitter=0.01), 'none'),
                                    ('LCB_auto', AcquisitionLCB(model, None, None, exploration_weight=2), 'none')):
        lp = AcquisitionLP(model, None, None, inner, transform=transform)
        assert lp.transform == str(g[tname + '_transform'])
        lp.update_batches(None, None, None)
        f0 = lp.acquisition_function(Xs)
        assert f0.ndim == 1 and scaled_err(f0, g[tname + '_f0']) < TOL
        lp.update_batches(g['X_batch'], float(g['L']), float(g['Min']))
        assert np.allclose(lp.r_x0, g[tname + '_r'], rtol=1e-9) and np.allclose(lp.s_x0, g[tname + '_s'], rtol=1e-9)
        f = lp.acquisition_function(Xs)
        assert scaled_err(f, g[tname + '_f']) < 1e-9
        grads = np.vstack([lp.acquisition_function_withGradients(x[None, :])[1] for x in Xs[:16]])
        assert_grad_close(grads, g[tname + '_g'][:16], tol=1e-9)


def test_ml2_fit_reproduces_reference_known_answer():
    """GPyOpt/testing/core_tests/test_cost.py:14-19,35-42 through the B200 GPModel (default Matern52, ML-II with 5
    restarts driven by the device log-likelihood and gradient)."""
    from gpyopt_b200 import GPModel
    np.random.seed(0)
    X = np.array([[3., 0.], [4., 1.], [5., 1.]])
    Y = np.log(np.array([[4.], [5.], [6.]]))
    model = GPModel(verbose=False)
    model.updateModel(X, Y, None, None)
    m, _, _, _ = model.predict_withGradients(X)
    assert np.allclose(np.exp(m), [[4.00000008], [5.00000038], [5.99999939]], rtol=1e-5)
    m, _, dmdx, _ = model.predict_withGradients(np.array([2., 2.]))
    assert np.isclose(np.exp(m)[0, 0], 3.52110177, rtol=1e-5)
    assert np.allclose((np.exp(m) * dmdx)[0], [0.65617088, 0.08849139], rtol=1e-4)
    assert model.get_model_parameters_names() == ['GP_regression.Mat52.variance', 'GP_regression.Mat52.lengthscale',
                                                  'GP_regression.Gaussian_noise.variance']


def test_x_next_parity_engine_vs_oracle():
    """Acquisition optimisation as AcquisitionOptimizer does it (anchor sweep -> top-5 -> L-BFGS-B from each anchor,
    GPyOpt/optimization/acquisition_optimizer.py:47-78, optimizer.py:28-61): the suggested x_next from the engine-backed
    acquisition must match the oracle-backed one within 1e-6."""
    import scipy.optimize
    import gpyopt_b200.gp as gpmod
    from gpyopt_b200 import GPModel, AcquisitionEI, kern, Engine
    from fake_engine import FakeEngine
    from oracle import gpyopt_numpy as og
    rng = np.random.default_rng(123)
    bounds = [(-5., 10.), (1., 15.)]
    X = np.column_stack([rng.uniform(lo, hi, 25) for lo, hi in bounds])
    Y = og.normalize_stats(og.branin(X))
    cand = np.column_stack([rng.uniform(lo, hi, 1000) for lo, hi in bounds])

    def suggest(engine_cls):
        gpmod.Engine = engine_cls
        try:
            model = GPModel(kernel=kern.RBF(2, variance=1.0, lengthscale=2.5), exact_feval=True, max_iters=0, verbose=False)
            model.updateModel(X, Y, None, None)
            acq = AcquisitionEI(model, None, None, jitter=0.01)
            scores = acq.acquisition_function(cand).flatten()
            anchors = cand[np.argsort(scores, kind='stable')[:5]]
            best = (np.inf, None)
            for a in anchors:
                def f_df(x):
                    f, df = acq.acquisition_function_withGradients(np.atleast_2d(x))
                    return float(f[0, 0]), df[0]
                res = scipy.optimize.fmin_l_bfgs_b(f_df, x0=a, bounds=bounds, maxiter=1000)
                if res[1] < best[0]:
                    best = (res[1], res[0])
            return best
        finally:
            gpmod.Engine = Engine
    f_gpu, x_gpu = suggest(Engine)
    f_cpu, x_cpu = suggest(FakeEngine)
    assert np.max(np.abs(x_gpu - x_cpu)) < 1e-6 and abs(f_gpu - f_cpu) < 1e-9


def test_bo_loop_on_branin_engine_vs_oracle_trajectory():
    """Config 1 in miniature (Branin, 5 initial + 15 iterations, GP/RBF with ML-II refits, EI, L-BFGS from the 5 best
    of 1000 random anchors): the whole loop driven by the engine must follow the same trajectory as the loop driven by
    the CPU oracle (same seeds), and must improve on the initial design."""
    import scipy.optimize
    import gpyopt_b200.gp as gpmod
    from gpyopt_b200 import GPModel, AcquisitionEI, kern, Engine
    from fake_engine import FakeEngine
    from oracle import gpyopt_numpy as og
    bounds = [(-5., 10.), (1., 15.)]
    lo, hi = np.array(bounds).T

    def run(engine_cls, iters=15):
        gpmod.Engine = engine_cls
        try:
            np.random.seed(123)
            X = lo + (hi - lo) * np.random.rand(5, 2)
            Yraw = og.branin(X)
            model = GPModel(kernel=kern.RBF(2), exact_feval=True, optimize_restarts=3, verbose=False)
            for it in range(iters):
                model.updateModel(X, og.normalize_stats(Yraw), None, None)
                acq = AcquisitionEI(model, None, None, jitter=0.01)
                cand = lo + (hi - lo) * np.random.rand(1000, 2)
                _, idx = model.model.engine.acq_topk('EI', 0.01, cand, 5)
                best = (np.inf, None)
                for a in cand[idx]:
                    def f_df(x):
                        f, df = acq.acquisition_function_withGradients(np.atleast_2d(x))
                        return float(f[0, 0]), df[0]
                    res = scipy.optimize.fmin_l_bfgs_b(f_df, x0=a, bounds=bounds, maxiter=200)
                    if res[1] < best[0]:
                        best = (res[1], res[0])
                X = np.vstack([X, best[1]])
                Yraw = np.vstack([Yraw, og.branin(best[1])])
            return X, Yraw
        finally:
            gpmod.Engine = Engine
    Xg, Yg = run(Engine)
    Xc, Yc = run(FakeEngine)
    assert np.max(np.abs(Xg - Xc)) < 1e-5, np.max(np.abs(Xg - Xc))
    assert Yg.min() < Yg[:5].min()


# ------------------------------------------------------------------------------------------------------
# HMC: GPModel_MCMC.updateModel on the engine against the oracle's HMC
# ------------------------------------------------------------------------------------------------------
def _hmc_problem(n=60, d=2, seed=11):
    from oracle import gpyopt_numpy as og
    rng = np.random.default_rng(seed)
    X = rng.random((n, d))
    Y = og.normalize_stats(np.sin(4 * X[:, :1]) + np.cos(3 * X[:, 1:2]) + 0.1 * rng.standard_normal((n, 1)))
    return X, Y


def _oracle_mcmc_model(X, Y, theta0):
    """The inner model GPModel_MCMC._create_model builds (gpmodel.py:213-238), on the oracle."""
    from oracle import gpy_numpy as G
    gp = G.GPRegression(X, Y, G.RBF(X.shape[1], variance=theta0[0], lengthscale=theta0[1]), noise_var=theta0[2])
    gp.kern.set_prior(G.Gamma.from_EV(2., 4.))
    gp.likelihood.variance.set_prior(G.Gamma.from_EV(2., 4.))
    gp.Gaussian_noise.constrain_positive(warning=False)
    return gp


def test_hmc_chain_follows_the_oracle_chain_with_the_same_seed():
    """Same NumPy seed, same starting point: the leapfrog trajectories are driven by the DEVICE log-likelihood gradient on
    one side and by the oracle's on the other.  40 draws x 20 leapfrog steps (1600 refits): the same accept / reject
    sequence and the same samples to 1e-8."""
    from oracle import gpy_numpy as G
    from gpyopt_b200 import GPModel_MCMC
    from gpyopt_b200.gp import HMC
    X, Y = _hmc_problem()
    theta0 = np.array([1.3, 0.4, 0.05])
    model = GPModel_MCMC(step_size=1e-1, leapfrog_steps=20)
    model._create_model(X, Y)
    model.model[:] = theta0
    gp = _oracle_mcmc_model(X, Y, theta0)
    assert np.allclose(model.model.optimizer_array, gp.optimizer_array, rtol=1e-14)
    assert abs(model.model.objective_function() - gp.objective_function()) < 1e-9 * abs(gp.objective_function())
    g_e = model.model._transform_gradients(model.model.objective_function_gradients())
    g_o = gp._transform_gradients(gp.objective_function_gradients())
    assert scaled_err(g_e, g_o) < 1e-9
    np.random.seed(7)
    ss_e = HMC(model.model, stepsize=1e-1).sample(num_samples=40, hmc_iters=20)
    np.random.seed(7)
    ss_o = G.HMC(gp, stepsize=1e-1).sample(num_samples=40, hmc_iters=20)
    moved_e = np.any(np.diff(ss_e, axis=0) != 0, axis=1)
    moved_o = np.any(np.diff(ss_o, axis=0) != 0, axis=1)
    assert np.array_equal(moved_e, moved_o)                       # same accepted / rejected proposals
    assert 5 < moved_o.sum()                                      # (the chain does move)
    assert np.max(np.abs(ss_e - ss_o) / np.abs(ss_o)) < 1e-8, np.max(np.abs(ss_e - ss_o) / np.abs(ss_o))


def test_gpmodel_mcmc_updatemodel_matches_the_oracle_flow_and_distribution():
    """The whole GPModel_MCMC.updateModel (gpmodel.py:240-255: optimize(max_iters=200), 1 % parameter jitter, HMC, thinning)
    on the engine against the same flow on the oracle: same seed -> same samples (ML-II end points agree to ~1e-7, the chain
    keeps them together to 1e-5 over the first thinned draws), and 200 thinned samples from a DIFFERENT seed agree in
    distribution (mean of log theta within 3 standard errors, spread within a factor 1.5)."""
    from oracle import gpy_numpy as G
    from gpyopt_b200 import GPModel_MCMC, AcquisitionEI_MCMC
    X, Y = _hmc_problem()

    def oracle_flow(seed, n_samples, n_burnin, interval, leapfrog):
        np.random.seed(seed)
        gp = _oracle_mcmc_model(X, Y, np.array([1.0, 1.0, Y.var() * 0.01]))
        gp.optimize(max_iters=200)
        gp.param_array[:] = gp.param_array * (1. + np.random.randn(gp.param_array.size) * 0.01)
        gp._trigger_params_changed()
        ss = G.HMC(gp, stepsize=1e-1).sample(num_samples=n_burnin + n_samples * interval, hmc_iters=leapfrog)
        return ss[n_burnin::interval]

    model = GPModel_MCMC(n_samples=10, n_burnin=20, subsample_interval=2, step_size=1e-1, leapfrog_steps=10)
    np.random.seed(3)
    model.updateModel(X, Y, None, None)
    ref = oracle_flow(3, 10, 20, 2, 10)
    assert model.hmc_samples.shape == ref.shape == (10, 3)
    assert np.max(np.abs(model.hmc_samples - ref) / np.abs(ref)) < 1e-5
    # the samples are usable by the integrated acquisition straight away
    f = AcquisitionEI_MCMC(model, None, None, jitter=0.01)._compute_acq(X[:5])
    assert f.shape == (5, 1) and np.all(np.isfinite(f))
    # distribution: 200 thinned samples, engine seed 21 vs oracle seed 22
    model = GPModel_MCMC(n_samples=200, n_burnin=100, subsample_interval=3, step_size=1e-1, leapfrog_steps=10)
    np.random.seed(21)
    model.updateModel(X, Y, None, None)
    le, lo = np.log(model.hmc_samples), np.log(oracle_flow(22, 200, 100, 3, 10))
    ess = 200 / 4.0                                              # thinned chain: conservative effective sample size
    se = np.sqrt(le.var(0) / ess + lo.var(0) / ess)
    assert np.all(np.abs(le.mean(0) - lo.mean(0)) < 3 * se), (le.mean(0), lo.mean(0), se)
    assert np.all(le.std(0) / lo.std(0) < 1.5) and np.all(lo.std(0) / le.std(0) < 1.5)


# ------------------------------------------------------------------------------------------------------
# batched multi-start L-BFGS on the device -- needs no GPyOpt: driven against SciPy and against the
# NumPy mirror of the state machine on the same engine
# ------------------------------------------------------------------------------------------------------
def _lbfgs_problem(case):
    from oracle import gpyopt_numpy as og
    rng = np.random.default_rng(3)
    if case == 'branin':
        b = np.array([[-5., 10.], [1., 15.]]); n, ls, nz, fn = 20, 3.0, 1e-6, og.branin
    elif case == 'alpine6':
        b = np.array([[1., 10.]] * 6); n, ls, nz, fn = 300, 2.0, 1e-2, og.alpine2
    else:
        b = np.array([[-4., 6.]] * 10); n, ls, nz, fn = 500, 3.0, 1e-2, lambda X: og.gSobol(X, np.ones(10))
    X = b[:, 0] + (b[:, 1] - b[:, 0]) * rng.random((n, b.shape[0]))
    Y = og.normalize_stats(fn(X))[:, 0]
    Xs = b[:, 0] + (b[:, 1] - b[:, 0]) * rng.random((2000, b.shape[0]))
    return X, Y, b, ls, nz, Xs


@pytest.mark.parametrize('case,kind,par,M', [('branin', 'EI', 0.01, 5), ('branin', 'LCB', 2.0, 5), ('alpine6', 'EI', 0.01, 5),
                                             ('alpine6', 'LCB', 2.0, 16), ('gsobol10', 'EI', 0.01, 32), ('gsobol10', 'MPI', 0.01, 40)])
def test_device_lbfgs_against_scipy_and_mirror(case, kind, par, M):
    from scipy.optimize import fmin_l_bfgs_b
    from gpyopt_b200 import Engine
    from fake_engine import lbfgs_mirror
    X, Y, b, ls, nz, Xs = _lbfgs_problem(case)
    d = b.shape[0]
    eng = Engine(0)
    try:
        eng.fit('rbf', X, Y, [1.0], np.full((1, d), ls), [nz])
        _, idx = eng.acq_topk(kind, par, Xs, M)
        A = Xs[idx]

        def fun(z):
            f, df = eng.acq(kind, par, z[None, :], grad=True)
            return -float(f[0]), -df[0]

        opts = dict(maxiter=1000, pgtol=1e-9, factr=10.0)
        x, f, it, nf = eng.lbfgs(kind, par, A, b, **opts)
        assert np.all(x >= b[:, 0]) and np.all(x <= b[:, 1]) and np.all(nf >= 1)
        fa, ga = eng.acq(kind, par, x, grad=True)
        assert np.allclose(-fa, f, rtol=1e-12, atol=1e-300)                      # the reported value is the value at x
        # (1) a stationary point of the box-constrained problem: projected gradient ~ 0 relative to the start's
        pg = np.max(np.abs(np.clip(x + ga, b[:, 0], b[:, 1]) - x), axis=1)
        f0, g0 = eng.acq(kind, par, A, grad=True)
        pg0 = np.max(np.abs(np.clip(A + g0, b[:, 0], b[:, 1]) - A), axis=1)
        assert np.all(f <= -f0 + 1e-15) and np.all(pg <= np.maximum(1e-3 * pg0, 1e-6)), (pg, pg0)   # (most runs stop on factr)
        # (2) the NumPy mirror of the state machine, driven by single-point calls of the same engine, walks the same path
        same = 0
        for a in range(min(M, 8)):
            xm, fm, itm, nfm = lbfgs_mirror(fun, A[a], b[:, 0], b[:, 1], opts['maxiter'], opts['pgtol'], opts['factr'])
            if np.max(np.abs(xm - x[a])) < 1e-7 * np.max(b[:, 1] - b[:, 0]) and abs(fm - f[a]) <= 1e-9 * max(1.0, abs(fm)):
                same += 1
        assert same >= min(M, 8) - 1, same
        # (3) SciPy's L-BFGS-B from the same anchors: same optimum within 1e-6 wherever both end in the same basin, never worse
        close = 0
        for a in range(min(M, 8)):
            r = fmin_l_bfgs_b(fun, A[a], bounds=[tuple(bb) for bb in b], maxiter=1000, pgtol=1e-9, factr=10.0)
            if np.max(np.abs(r[0] - x[a])) < 1e-6 * np.max(b[:, 1] - b[:, 0]):
                close += 1
                assert abs(r[1] - f[a]) <= 1e-9 * max(1.0, abs(r[1]))
            else:
                assert f[a] <= r[1] + 1e-9 * max(1.0, abs(r[1])) or np.max(np.abs(r[0] - x[a])) > 1e-3
        assert close >= (min(M, 8) * 3) // 4, close
        # (4) polishing the best point with a converged SciPy run does not move it
        best = int(np.argmin(f))
        r = fmin_l_bfgs_b(fun, x[best], bounds=[tuple(bb) for bb in b], maxiter=200, pgtol=1e-10, factr=10.0)
        assert np.max(np.abs(r[0] - x[best])) < 1e-6 * np.max(b[:, 1] - b[:, 0])
        # (5) SciPy's default stopping rules: same optimum to the slack of those rules
        x2, f2, _, nf2 = eng.lbfgs(kind, par, A, b, maxiter=1000, pgtol=1e-5, factr=1e7)
        near = np.max(np.abs(x2 - x), axis=1) < 1e-3 * np.max(b[:, 1] - b[:, 0])
        assert np.all(f2 <= -f0 + 1e-15) and near.sum() >= (M * 3) // 5
        assert np.all(np.abs(f2 - f)[near] <= 1e-6 * np.maximum(1.0, np.abs(f[near])))
    finally:
        eng.close()


def test_device_lbfgs_under_local_penalization_and_maxiter():
    """Local Penalization active (the minimised function is the penalised value itself) and the maxiter stop."""
    from scipy.optimize import fmin_l_bfgs_b
    from gpyopt_b200 import Engine
    X, Y, b, ls, nz, Xs = _lbfgs_problem('alpine6')
    eng = Engine(0)
    try:
        eng.fit('rbf', X, Y, [1.0], np.full((1, 6), ls), [nz])
        Xb = Xs[:3]
        eng.lp_set(Xb, np.full(3, 0.8), np.full(3, 0.4), 'softplus')
        _, idx = eng.acq_topk('EI', 0.01, Xs, 6)
        A = Xs[idx]
        x, f, it, nf = eng.lbfgs('EI', 0.01, A, b, maxiter=1000, pgtol=1e-9, factr=10.0)
        v, g = eng.acq('EI', 0.01, x, grad=True)
        assert np.allclose(v, f, rtol=1e-12)

        def fun(z):
            vv, gg = eng.acq('EI', 0.01, z[None, :], grad=True)
            return float(vv[0]), gg[0]
        v0 = eng.acq('EI', 0.01, A)
        assert np.all(f <= v0 + 1e-15)
        best = int(np.argmin(f))
        r = fmin_l_bfgs_b(fun, x[best], bounds=[tuple(bb) for bb in b], maxiter=200, pgtol=1e-10, factr=10.0)
        assert np.max(np.abs(r[0] - x[best])) < 1e-5 * 9.0 and r[1] >= f[best] - 1e-9 * max(1.0, abs(f[best]))
        x3, f3, it3, _ = eng.lbfgs('EI', 0.01, A, b, maxiter=2, pgtol=1e-12, factr=0.0)
        assert np.all(it3 <= 2) and np.all(f3 >= f - 1e-12)
    finally:
        eng.lp_set(None, None, None, None)
        eng.close()
