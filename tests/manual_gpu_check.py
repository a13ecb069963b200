"""Bring-up check on a B200: fit state vs LAPACK, sweep outputs vs the oracle, quick timings."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200._lib import Engine
from oracle import gpy_numpy as G, gpyopt_numpy as og

def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))

def case(n, d, kernel, ARD, noise, N, seed=0, S=1):
    rng = np.random.default_rng(seed)
    X = 1 + 9 * rng.random((n, d)); Y = og.normalize_stats(og.alpine2(X))
    Xs = 1 + 9 * np.random.default_rng(seed + 1).random((N, d))
    ls = (1.5 + 1.5 * rng.random(d)) if ARD else np.array([2.0])
    K = {'rbf': G.RBF, 'matern52': G.Matern52, 'matern32': G.Matern32}[kernel]
    gp = G.GPRegression(X, Y, K(d, variance=1.3, lengthscale=ls, ARD=ARD), noise_var=noise)
    eng = Engine(0)
    t0 = time.time(); eng.fit(kernel, X, Y, [1.3], ls[None, :], [noise]); t1 = time.time()
    eng.fit(kernel, X, Y, [1.3], ls[None, :], [noise]); t2 = time.time()
    L, Li, Wi, al = eng.get_state()
    print('--- n=%d d=%d %s ARD=%s noise=%g  fit %.1f ms (first %.1f ms)' % (n, d, kernel, ARD, noise, (t2 - t1) * 1e3, (t1 - t0) * 1e3))
    print('  L    ', rel(np.tril(L[0]), gp.posterior.woodbury_chol))
    print('  Linv ', rel(Li[0], np.linalg.inv(gp.posterior.woodbury_chol)))
    print('  Winv ', rel(Wi[0], gp.posterior.woodbury_inv), 'sym', float(np.abs(Wi[0] - Wi[0].T).max()))
    print('  alpha', rel(al[0], gp.posterior.woodbury_vector[:, 0]))
    print('  fmin ', eng.get_fmin()[0], og.gp_get_fmin(gp))
    ll, ld = eng.get_loglik(); print('  loglik', ll[0], gp.log_likelihood())
    g = eng.get_loglik_grad()[0]
    gref = gp.gradient
    gl = g[1:1 + d] if ARD else np.array([g[1:1 + d].sum()])
    print('  dloglik', rel(np.concatenate([[g[0]], gl, [g[d + 1]]]), gref))
    m, s = eng.predict(Xs); mo, so = og.gp_predict(gp, Xs)
    print('  predict m', rel(m[0], mo[:, 0]), 's', rel(s[0], so[:, 0]))
    m, s, dm, ds = eng.predict_grad(Xs); mo, so, dmo, dso = og.gp_predict_withGradients(gp, Xs)
    print('  predict_grad m', rel(m[0], mo[:, 0]), 's', rel(s[0], so[:, 0]), 'dm', rel(dm[0], dmo), 'ds', rel(ds[0], dso))
    fmin = og.gp_get_fmin(gp)
    f = eng.acq('EI', 0.01, Xs); fo = og.acq_EI(mo.copy(), so.copy(), fmin, 0.01)
    f2, df = eng.acq('EI', 0.01, Xs, grad=True); fo2, dfo = og.acq_EI_withGradients(mo.copy(), so.copy(), dmo, dso, fmin, 0.01)
    print('  EI', rel(f, fo[:, 0]), 'EI(g)', rel(f2, fo2[:, 0]), 'dEI', rel(df, dfo), 'argmax', int(np.argmax(f)), int(np.argmax(fo)))
    v, i = eng.acq_topk('EI', 0.01, Xs, 5); print('  top5', i, og.anchor_topk(-fo[:, 0], 5))
    eng.close()

if __name__ == '__main__':
    case(100, 2, 'rbf', False, 1e-2, 300)
    case(300, 6, 'rbf', True, 1e-2, 1000)
    case(640, 6, 'matern52', False, 1e-3, 500)
    case(130, 3, 'matern32', True, 1e-6, 200)
    case(2048, 6, 'rbf', False, 1e-2, 4096)
    # timing at the benchmark shape
    import torch
    rng = np.random.default_rng(0)
    n, d, N = 2048, 6, 9472 * 8
    X = 1 + 9 * rng.random((n, d)); Y = og.normalize_stats(og.alpine2(X))
    eng = Engine(0); eng.fit('rbf', X, Y, [1.0], [[2.0]], [1e-2])
    Xs = torch.from_numpy(1 + 9 * rng.random((N, d))).cuda()
    f = torch.empty(N, dtype=torch.float64, device='cuda'); df = torch.empty(N, d, dtype=torch.float64, device='cuda')
    for grad in (True, False):
        for it in range(3):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); eng.acq('EI', 0.01, Xs, grad=grad, out=(f, df if grad else None), stream=torch.cuda.current_stream().cuda_stream); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            fl = (2 * n * n if grad else n * n) * N
            print('grad=%s  N=%d  %.2f ms  %.3f Mcand/s  %.2f TFLOP/s (dense part)' % (grad, N, ms, N / ms / 1e3, fl / ms / 1e9))
