"""Generate the committed golden vectors under tests/golden/*.npz.

Runs the UNMODIFIED reference GPyOpt (/root/reference) -- its GPModel / GPModel_MCMC wrappers,
AcquisitionEI/MPI/LCB(+_MCMC), AcquisitionLP and anchor-point selection -- on top of the oracle's
GPy restatement (oracle/gpy_stub/GPy -> oracle/gpy_numpy.py; GPy itself is not installable here).
Only runs in the build container (needs /root/reference); the .npz files travel to the GPU box.

    python tests/golden/make_golden.py
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'gpy_stub'))
sys.path.insert(0, '/root/reference')
warnings.filterwarnings('ignore')

import GPy  # noqa: E402  (the stub)
import GPyOpt  # noqa: E402  (the reference)
from GPyOpt.models.gpmodel import GPModel, GPModel_MCMC  # noqa: E402
from GPyOpt.acquisitions import (AcquisitionEI, AcquisitionMPI, AcquisitionLCB, AcquisitionEI_MCMC,  # noqa: E402
                                 AcquisitionMPI_MCMC, AcquisitionLCB_MCMC, AcquisitionLP)
from GPyOpt.optimization.anchor_points_generator import ObjectiveAnchorPointsGenerator  # noqa: E402
from oracle import gpyopt_numpy as og  # noqa: E402

KERNELS = {'rbf': GPy.kern.RBF, 'matern52': GPy.kern.Matern52, 'matern32': GPy.kern.Matern32}


def space_for(bounds):
    return GPyOpt.Design_space([{'name': 'x%d' % i, 'type': 'continuous', 'domain': tuple(b)}
                                for i, b in enumerate(bounds)])


def uniform(rng, bounds, n):
    b = np.asarray(bounds, dtype=float)
    return b[:, 0] + (b[:, 1] - b[:, 0]) * rng.random((n, len(bounds)))


def make_gp_case(name, fn, bounds, n, N, kernel, variance, lengthscale, ARD, noise_var, exact_feval, seed, mlii=False):
    """mlii: run the reference's own ML-II (GPModel.updateModel with max_iters=1000, one restart, seeded) before taking the
    vectors -- GPyOpt's default flow; with exact_feval=True (noise fixed at 1e-6) it lands on ill-conditioned Gram
    matrices (cond 1e6 ... 1e9), where the variance form and the explicit inverses matter."""
    rng = np.random.default_rng(seed)
    d = len(bounds)
    X = uniform(rng, bounds, n)
    Y = og.normalize_stats(fn(X))
    Xs = uniform(np.random.default_rng(seed + 1), bounds, N)
    # a few candidates on / next to training points (r == 0 and tiny-variance edge cases)
    Xs[0] = X[0]
    Xs[1] = X[1] + 1e-9
    kern = KERNELS[kernel](d, variance=variance, lengthscale=lengthscale, ARD=ARD)
    model = GPModel(kernel=kern, noise_var=noise_var, exact_feval=exact_feval, max_iters=1000 if mlii else 0,
                    optimize_restarts=1, verbose=False)
    np.random.seed(seed)
    model.updateModel(X, Y, None, None)
    if mlii:   # the fitted hyper-parameters become the case's fixed ones
        variance = float(model.model.kern.variance[0])
        lengthscale = np.array(model.model.kern.lengthscale, dtype=float).copy()
        model.max_iters = 0
    space = space_for(bounds)
    out = dict(X=X, Y=Y, Xs=Xs, bounds=np.asarray(bounds, float), kernel=kernel, variance=variance,
               lengthscale=np.atleast_1d(np.asarray(lengthscale, float)), ARD=ARD,
               noise_var=float(model.model.likelihood._noise()), exact_feval=exact_feval)
    Ky = model.model.kern.K(X) + (out['noise_var'] + 1e-8) * np.eye(n)
    out['cond'] = float(np.linalg.cond(Ky))
    m, s = model.predict(Xs)
    m0, s0 = model.predict(Xs, with_noise=False)
    mg, sg, dmdx, dsdx = model.predict_withGradients(Xs)
    out.update(m=m, s=s, m_nonoise=m0, s_nonoise=s0, mg=mg, sg=sg, dmdx=dmdx, dsdx=dsdx,
               fmin=float(model.get_fmin()), params=model.get_model_parameters(),
               m_1d=model.predict(Xs[5])[0])
    for aname, acq in (('EI', AcquisitionEI(model, space, jitter=0.01)),
                       ('MPI', AcquisitionMPI(model, space, jitter=0.01)),
                       ('LCB', AcquisitionLCB(model, space, exploration_weight=2))):
        out['%s_f' % aname] = acq._compute_acq(Xs)
        f, df = acq._compute_acq_withGradients(Xs)
        out['%s_fg' % aname], out['%s_df' % aname] = f, df
        out['%s_af' % aname] = acq.acquisition_function(Xs)
        af, adf = acq.acquisition_function_withGradients(Xs)
        out['%s_afg' % aname], out['%s_adf' % aname] = af, adf
        scores = acq.acquisition_function(Xs).flatten()
        out['%s_top5' % aname] = np.argsort(scores)[:5]
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'fmin', out['fmin'], 'EI max', out['EI_f'].max(), 'min s', s.min(), 'cond %.2e' % out['cond'], 'theta', variance, lengthscale)
    return model, space, out


def make_mcmc_case(name, bounds, n, N, S, seed, exact_feval=False):
    rng = np.random.default_rng(seed)
    d = len(bounds)
    X = uniform(rng, bounds, n)
    Y = og.normalize_stats(og.alpine2(1 + 9 * X))
    Xs = uniform(np.random.default_rng(seed + 1), bounds, N)
    model = GPModel_MCMC(exact_feval=exact_feval, verbose=False)
    model._create_model(X, Y)
    theta0 = np.array([1.0, 0.5, 1e-2])[: (2 if exact_feval else 3)]
    model.hmc_samples = theta0 * np.exp(0.1 * rng.standard_normal((S, theta0.size)))
    space = space_for(bounds)
    out = dict(X=X, Y=Y, Xs=Xs, bounds=np.asarray(bounds, float), hmc_samples=model.hmc_samples,
               exact_feval=exact_feval, kernel='rbf')
    means, stds = model.predict(Xs)
    mg, sg, dm, ds = model.predict_withGradients(Xs)
    out.update(means=np.array(means), stds=np.array(stds), dmdxs=np.array(dm), dsdxs=np.array(ds),
               fmins=np.array(model.get_fmin()))
    for aname, acq in (('EI_MCMC', AcquisitionEI_MCMC(model, space, jitter=0.01)),
                       ('MPI_MCMC', AcquisitionMPI_MCMC(model, space, jitter=0.01)),
                       ('LCB_MCMC', AcquisitionLCB_MCMC(model, space, exploration_weight=2))):
        out['%s_f' % aname] = acq._compute_acq(Xs)
        f, df = acq._compute_acq_withGradients(Xs)
        out['%s_fg' % aname], out['%s_df' % aname] = f, df
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'fmins', out['fmins'][:3], 'EI_MCMC max', out['EI_MCMC_f'].max())


def make_lp_case(name, seed):
    bounds = [(-4, 6)] * 10
    a = np.ones(10)
    model, space, base = make_gp_case(name + '_gp', lambda X: og.gSobol(X, a), bounds, 128, 96, 'rbf', 1.0, 3.0,
                                      False, 1e-2, False, seed)
    Xs = base['Xs']
    rng = np.random.default_rng(seed + 7)
    X_batch = uniform(rng, bounds, 3)
    X_batch[0] = Xs[10] + 0.05  # a penaliser close to a candidate
    L, Min = 0.37, float(model.model.Y.min())
    out = dict(Xs=Xs, X_batch=X_batch, L=L, Min=Min)
    for tname, inner, transform in (('EI_none', AcquisitionEI(model, space, jitter=0.01), 'none'),
                                    ('EI_softplus', AcquisitionEI(model, space, jitter=0.01), 'softplus'),
                                    ('LCB_auto', AcquisitionLCB(model, space, exploration_weight=2), 'none')):
        lp = AcquisitionLP(model, space, None, inner, transform=transform)
        out['%s_transform' % tname] = lp.transform
        lp.update_batches(None, None, None)
        out['%s_f0' % tname] = lp.acquisition_function(Xs)
        out['%s_g0' % tname] = np.vstack([lp.acquisition_function_withGradients(x[None, :])[1] for x in Xs])
        lp.update_batches(X_batch, L, Min)
        out['%s_r' % tname], out['%s_s' % tname] = lp.r_x0, lp.s_x0
        out['%s_f' % tname] = lp.acquisition_function(Xs)
        out['%s_g' % tname] = np.vstack([lp.acquisition_function_withGradients(x[None, :])[1] for x in Xs])
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'LP f range', out['EI_none_f'].min(), out['EI_none_f'].max())


def make_epilogue_known_answers():
    """The reference's own mocked known answers (GPyOpt/testing/acquisitions_tests/*), kept as data so
    the GPU box (no /root/reference) can replay them against the epilogue kernel."""
    out = dict(
        ei_m=1.0, ei_s=3.0, ei_fmin=0.1, ei_jitter=0.01, ei_expected=0.79646919,   # test_ei_acquisition.py:21-26
        mpi_expected=0.38081792,                                                       # test_mpi_acquisition.py:19-26
        eig_m=1.0, eig_s=1.0, eig_fmin=0.0, eig_dmdx=0.1, eig_dsdx=0.2,                 # test_ei_acquisition.py:31-37
    )
    np.savez_compressed(os.path.join(HERE, 'epilogue_known_answers.npz'), **out)


if __name__ == '__main__':
    b6 = [(1, 10)] * 6
    make_gp_case('c3_rbf_alpine2_n256', og.alpine2, b6, 256, 384, 'rbf', 1.0, 2.0, False, 1e-2, False, 0)
    make_gp_case('c3_rbf_ard_exact_n200', og.alpine2, b6, 200, 256, 'rbf', 1.3, [1.5, 1.8, 2.1, 2.4, 2.7, 3.0], True,
                 None, True, 10)
    make_gp_case('c3_mat52_alpine2_n192', og.alpine2, b6, 192, 256, 'matern52', 0.8, 2.5, False, None, False, 20)
    make_gp_case('c3_mat32_ard_n130', og.alpine2, b6, 130, 200, 'matern32', 1.1, [2.0, 2.2, 2.4, 2.6, 2.8, 3.0], True,
                 1e-3, False, 30)
    make_gp_case('c1_branin_rbf_n20', og.branin, [(-5, 10), (1, 15)], 20, 300, 'rbf', 1.0, 3.0, False, None, True, 40)
    make_gp_case('c2_camel_mat52_n55', og.sixhumpcamel, [(-2, 2), (-1, 1)], 55, 300, 'matern52', 1.0, 0.7, False,
                 None, False, 50)
    make_gp_case('c1_1d_rbf_n7', lambda X: np.sin(3 * X) + X, [(0, 2)], 7, 64, 'rbf', 1.0, 0.4, False, 1e-4, False, 60)
    # ill-conditioned cases: GPyOpt's default flow (exact_feval=True -> noise 1e-6, ML-II)
    make_gp_case('ill_branin_rbf_exact_n40_mlii', og.branin, [(-5, 10), (1, 15)], 40, 300, 'rbf', 1.0, 1.0, False, None, True, 140,
                 mlii=True)
    make_gp_case('ill_branin_mat52_exact_n70_mlii', og.branin, [(-5, 10), (1, 15)], 70, 300, 'matern52', 1.0, 1.0, False, None,
                 True, 154, mlii=True)
    make_gp_case('ill_alpine2_rbf_ard_exact_n300_mlii', lambda X: og.alpine2(X), [(1, 10)] * 3, 300, 300, 'rbf', 1.0,
                 [1.0, 1.0, 1.0], True, None, True, 160, mlii=True)
    make_gp_case('ill_branin_rbf_exact_n150_fixed', og.branin, [(-5, 10), (1, 15)], 150, 300, 'rbf', 2.0, 4.0, False, None, True, 170)
    make_mcmc_case('c4_mcmc_n64_S8', [(0, 1)] * 4, 64, 128, 8, 2)
    make_mcmc_case('c4_mcmc_exact_n48_S4', [(0, 1)] * 4, 48, 64, 4, 3, exact_feval=True)
    make_lp_case('c5_lp_gsobol_n128', 70)
    make_epilogue_known_answers()
