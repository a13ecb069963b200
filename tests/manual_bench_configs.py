"""Secondary measurements for the BASELINE.json configs other than the headline one (one JSON line each).
  C2  six-hump camel, n=55, LCB, 1M-candidate value-only sweep
  C4  GP_MCMC: S=64 hyper-parameter samples, n=512, d=4, EI_MCMC + gradient
  C5  gSobol d=10, n=4096: refit latency, EI + Local Penalization (15 penalisers) sweep, single-point latency
  C1  Branin: wall time of 15 BO iterations (ML-II refits + anchor sweep + L-BFGS), engine vs CPU oracle
  LAT single-point (N=1) EI+gradient latency through the plugin call at n=64 and n=2048
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))  # (this file lives in tests/: it uses the oracle)
import torch
from gpyopt_b200 import Engine, GPModel, AcquisitionEI, kern
from oracle import gpyopt_numpy as og


def dev_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best * 1e-3


def wall(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def out(d):
    print(json.dumps(d)); sys.stdout.flush()


st = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(0)

# ---- C2 ------------------------------------------------------------------------------------------------
n, d, N = 55, 2, 1_000_000
lo, hi = np.array([-2., -1.]), np.array([2., 1.])
X = lo + (hi - lo) * rng.random((n, d)); Y = og.normalize_stats(og.sixhumpcamel(X))[:, 0]
eng = Engine(0); eng.fit('matern52', X, Y, [1.0], [[0.7, 0.7]], [1e-2])
Xs = torch.from_numpy(lo + (hi - lo) * rng.random((N, d))).cuda(); f = torch.empty(N, dtype=torch.float64, device='cuda')
t = dev_time(lambda: eng.acq('LCB', 2.0, Xs, out=(f, None), stream=st))
Xh = Xs.cpu().numpy()
th = wall(lambda: eng.acq_topk('LCB', 2.0, Xh, 5))
bounds = np.stack([lo, hi], 1)
tg = wall(lambda: eng.acq_topk_uniform('LCB', 2.0, [1, 2], N, bounds, 5))       # design generated on the device
out({"config": "C2 sixhumpcamel n=55 LCB value-only sweep", "candidates": N, "device_cand_per_s": N / t, "ms": t * 1e3,
     "host_sweep_plus_top5_cand_per_s": N / th, "device_design_sweep_plus_top5_cand_per_s": N / tg})
eng.close()

# ---- C3 anchor step: 10M-candidate random design -> EI -> top-5, host design vs device design ------------
n, d, N = 2048, 6, 10_000_000
X = rng.uniform(1, 10, (n, d)); Y = og.normalize_stats(og.alpine2(X))[:, 0]
eng = Engine(0); eng.fit('rbf', X, Y, [1.0], [[2.0]], [1e-2])
bounds = np.array([[1., 10.]] * d)
tg = wall(lambda: eng.acq_topk_uniform('EI', 0.01, [1, 2], N, bounds, 5), reps=2)
Xh = eng.design_uniform([1, 2], N, bounds)
th = wall(lambda: eng.acq_topk('EI', 0.01, Xh, 5), reps=2)
tnp = wall(lambda: og.design_uniform([1, 2], 1_000_000, bounds), reps=1, warm=0)
out({"config": "C3 anchor step n=2048 d=6 EI value-only, 10M random candidates -> top-5", "candidates": N,
     "device_design_cand_per_s": N / tg, "host_design_cand_per_s_excluding_generation": N / th,
     "numpy_philox_generation_only_cand_per_s": 1_000_000 / tnp})
eng.close()

# ---- C4 ------------------------------------------------------------------------------------------------
n, d, S, N = 512, 4, 64, 200_000
X = rng.random((n, d)); Y = og.normalize_stats(og.alpine2(1 + 9 * X))[:, 0]
theta = np.array([1.0, 0.5, 1e-2]) * np.exp(0.1 * np.random.default_rng(2).standard_normal((S, 3)))
eng = Engine(0)
tfit = wall(lambda: eng.fit('rbf', X, Y, theta[:, 0], theta[:, 1:2], theta[:, 2]), reps=3)
Xs = torch.from_numpy(rng.random((N, d))).cuda(); f = torch.empty(N, dtype=torch.float64, device='cuda'); df = torch.empty(N, d, dtype=torch.float64, device='cuda')
t = dev_time(lambda: eng.acq('EI_MCMC', 0.01, Xs, grad=True, out=(f, df), stream=st))
flops = 2.0 * 512 * 512 * S * N
out({"config": "C4 GP_MCMC S=64 n=512 d=4 EI_MCMC+grad", "candidates": N, "device_cand_per_s": N / t, "sample_evals_per_s": N * S / t,
     "dense_tflops": flops / t * 1e-12, "fit_all_64_samples_ms": tfit * 1e3, "info": eng.info()})
eng.close()

# ---- C5 ------------------------------------------------------------------------------------------------
n, d, N = 4096, 10, 500_000
X = -4 + 10 * rng.random((n, d)); Y = og.normalize_stats(og.gSobol(X, np.ones(d)))[:, 0]
eng = Engine(0)
tfit = wall(lambda: eng.fit('rbf', X, Y, [1.0], [[3.0] * d], [1e-2]), reps=2)
Xs = torch.from_numpy(-4 + 10 * rng.random((N, d))).cuda(); f = torch.empty(N, dtype=torch.float64, device='cuda'); df = torch.empty(N, d, dtype=torch.float64, device='cuda')
Xb = -4 + 10 * rng.random((15, d))
eng.lp_set(Xb, np.full(15, 0.3), np.full(15, 0.2), 'none')
t_val = dev_time(lambda: eng.acq('EI', 0.01, Xs, out=(f, None), stream=st))
t_grad = dev_time(lambda: eng.acq('EI', 0.01, Xs, grad=True, out=(f, df), stream=st))
x1 = -4 + 10 * rng.random((1, d))
t_lat = wall(lambda: eng.acq('EI', 0.01, x1, grad=True), reps=20, warm=3)
eng.lp_set(None, None, None, None)
out({"config": "C5 gSobol d=10 n=4096 EI + LP(15 penalisers)", "refit_ms": tfit * 1e3, "candidates": N,
     "value_sweep_cand_per_s": N / t_val, "grad_sweep_cand_per_s": N / t_grad, "grad_dense_tflops": 2.0 * 4096 * 4096 * N / t_grad * 1e-12,
     "single_point_f_df_latency_ms": t_lat * 1e3, "info": eng.info()})
eng.close()

# ---- C5 end to end: one batch iteration (ML-II refit + Local-Penalization batch of 16) ------------------------
import scipy.optimize
from gpyopt_b200 import AcquisitionLP
from gpyopt_b200 import evaluators as evs
n, d, B = 4096, 10, 16
bounds = [(-4.0, 6.0)] * d; lo, hi = np.array(bounds).T
X = lo + (hi - lo) * rng.random((n, d)); Y = og.normalize_stats(og.gSobol(X, np.ones(d)))
model = GPModel(kernel=kern.RBF(d, lengthscale=3.0), noise_var=1e-2, optimize_restarts=1, max_iters=30, verbose=False)
t0 = time.perf_counter(); model.updateModel(X, Y, None, None); t_fit = time.perf_counter() - t0
fits = getattr(model.model, 'n_refits', None)
ei = AcquisitionEI(model, None, None, jitter=0.01)
lp = AcquisitionLP(model, None, None, ei, transform='none')
eng = model.model.engine
np.random.seed(3)
calls = {'n': 0}


def optimise_once():
    cand = lo + (hi - lo) * np.random.rand(1000, d)
    _, idx = eng.acq_topk('EI', 0.01, cand, 5)
    best = (np.inf, None)
    for a0 in cand[idx]:
        def f_df(x):
            calls['n'] += 1
            f, df = lp.acquisition_function_withGradients(np.atleast_2d(x)); return float(np.ravel(f)[0]), np.ravel(df)
        res = scipy.optimize.fmin_l_bfgs_b(f_df, x0=a0, bounds=bounds, maxiter=1000)
        if res[1] < best[0]: best = (res[1], res[0])
    return best[1]


t0 = time.perf_counter()
lp.update_batches(None, None, None)
batch = optimise_once()[None, :]
L = evs.estimate_L(model.model, bounds); y_min = model.model.Y.min()
for _ in range(1, B):
    lp.update_batches(batch, L, y_min)
    batch = np.vstack([batch, optimise_once()])
lp.update_batches(None, None, None)
t_batch = time.perf_counter() - t0
out({"config": "C5 end to end: gSobol d=10 n=4096, ML-II refit (1 restart, <=30 L-BFGS iterations) + LP batch of 16 (1000-candidate sweep, top-5, 5 L-BFGS runs per element)",
     "updateModel_s": t_fit, "batch_of_16_s": t_batch, "acquisition_evaluations": calls['n'], "lipschitz_L": L,
     "distinct_batch_points": int(len(np.unique(np.round(batch, 6), axis=0)))})

# ... the same batch with the local optimisation inside the library (gpo_lbfgs: all 5 anchors in lock-step on the device)
def optimise_once_device(opts):
    cand = lo + (hi - lo) * np.random.rand(1000, d)
    _, idx = eng.acq_topk('EI', 0.01, cand, 5)
    eng.lp_set(lp.X_batch, lp.r_x0, lp.s_x0, lp.transform)
    try:
        x, f, it, nf = eng.lbfgs('EI', 0.01, cand[idx], np.array(bounds), **opts)
    finally:
        eng.lp_set(None, None, None, None)
    dev_stats['evals'] += int(nf.sum()); dev_stats['rounds'] += int(nf.max())
    return x[int(np.argmin(f))]


for label, opts in (("scipy-default stopping rules (pgtol 1e-5, factr 1e7)", dict(maxiter=1000, pgtol=1e-5, factr=1e7)),
                    ("tight stopping rules (pgtol 1e-9, factr 10)", dict(maxiter=1000, pgtol=1e-9, factr=10.0))):
    np.random.seed(3)
    dev_stats = {'evals': 0, 'rounds': 0}
    t0 = time.perf_counter()
    lp.update_batches(None, None, None)
    batch = optimise_once_device(opts)[None, :]
    L = evs.estimate_L(model.model, bounds); y_min = model.model.Y.min()
    for _ in range(1, B):
        lp.update_batches(batch, L, y_min)
        batch = np.vstack([batch, optimise_once_device(opts)])
    lp.update_batches(None, None, None)
    t_dev = time.perf_counter() - t0
    out({"config": "C5 end to end, LP batch of 16 with the device-side lock-step L-BFGS, " + label, "batch_of_16_s": t_dev,
         "acquisition_evaluations": dev_stats['evals'], "lockstep_rounds": dev_stats['rounds'],
         "distinct_batch_points": int(len(np.unique(np.round(batch, 6), axis=0)))})
eng.close()

# ---- LAT -----------------------------------------------------------------------------------------------
for n in (64, 2048):
    d = 6
    X = 1 + 9 * rng.random((n, d)); Y = og.normalize_stats(og.alpine2(X))
    m = GPModel(kernel=kern.RBF(d, lengthscale=2.0), noise_var=1e-2, max_iters=0, verbose=False); m.updateModel(X, Y, None, None)
    a = AcquisitionEI(m, None, None)
    x1 = 1 + 9 * rng.random((1, d))
    t1 = wall(lambda: a.acquisition_function_withGradients(x1), reps=50, warm=5)
    t0 = wall(lambda: a.acquisition_function(x1), reps=50, warm=5)
    tf = wall(lambda: m.model._trigger_params_changed(), reps=5, warm=1)
    out({"config": "LAT single point through the plugin", "n": n, "f_df_ms": t1 * 1e3, "f_ms": t0 * 1e3, "refit_ms": tf * 1e3})
    m.model.engine.close()

# ---- C1 ------------------------------------------------------------------------------------------------
import gpyopt_b200.gp as gpmod
from fake_engine import FakeEngine
bounds = [(-5., 10.), (1., 15.)]; lo, hi = np.array(bounds).T


def bo(engine_cls, iters=15):
    gpmod.Engine = engine_cls
    np.random.seed(123)
    X = lo + (hi - lo) * np.random.rand(5, 2); Yraw = og.branin(X)
    model = GPModel(kernel=kern.RBF(2), exact_feval=True, optimize_restarts=5, verbose=False)
    t0 = time.perf_counter()
    for it in range(iters):
        model.updateModel(X, og.normalize_stats(Yraw), None, None)
        acq = AcquisitionEI(model, None, None, jitter=0.01)
        cand = lo + (hi - lo) * np.random.rand(1000, 2)
        _, idx = model.model.engine.acq_topk('EI', 0.01, cand, 5)
        best = (np.inf, None)
        for a0 in cand[idx]:
            def f_df(x):
                f, df = acq.acquisition_function_withGradients(np.atleast_2d(x)); return float(f[0, 0]), df[0]
            res = scipy.optimize.fmin_l_bfgs_b(f_df, x0=a0, bounds=bounds, maxiter=1000)
            if res[1] < best[0]: best = (res[1], res[0])
        X = np.vstack([X, best[1]]); Yraw = np.vstack([Yraw, og.branin(best[1])])
    gpmod.Engine = Engine
    return time.perf_counter() - t0, float(Yraw.min())


tg, yg = bo(Engine); tc, yc = bo(FakeEngine)
out({"config": "C1 Branin 5+15 BO iterations, GP/RBF ML-II (5 restarts), EI, L-BFGS from 5 anchors", "engine_wall_s": tg, "engine_best_y": yg,
     "cpu_oracle_wall_s": tc, "cpu_oracle_best_y": yc})
