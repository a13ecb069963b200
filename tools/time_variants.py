"""Sweep-rate and refit-latency matrix over the GEMM tuning knobs (gpo_set_option) and the variance form, at the headline
shape (n=2048, d=6) and at config 4's (S=64, n=512, d=4).  One JSON line per measurement."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import Engine

st = torch.cuda.current_stream().cuda_stream


def rate(eng, kind, xs, f, df, grad, reps=3):
    out = (f, df) if grad else (f, None)
    for _ in range(2):
        eng.acq(kind, 0.01, xs, grad=grad, out=out, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.acq(kind, 0.01, xs, grad=grad, out=out, stream=st)
    e1.record(); torch.cuda.synchronize()
    return xs.shape[0] / (e0.elapsed_time(e1) / reps) * 1e3


def sweep_matrix(tag, n, d, S, N, kind):
    rng = np.random.default_rng(0)
    X = 1 + 9 * rng.random((n, d)); Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
    var = np.full(S, 1.0) * np.exp(0.05 * rng.standard_normal(S)) if S > 1 else np.array([1.0])
    ls = np.full((S, d), 2.0)
    nz = np.full(S, 1e-2)
    xs = torch.from_numpy(1 + 9 * rng.random((N, d))).cuda()
    f = torch.empty(N, dtype=torch.float64, device='cuda'); df = torch.empty((N, d), dtype=torch.float64, device='cuda')
    eng = Engine(0)
    for strict in (1, 0):
        eng.set_strict_variance(strict)
        eng.fit('rbf', X, Y, var, ls, nz)
        for bn in (128, 64):
            for serp in (0, 1):
                eng.set_option('bn_sweep', bn); eng.set_option('serpentine', serp)
                r = rate(eng, kind, xs, f, df, True)
                rv = rate(eng, kind, xs, f, df, False) if strict == 1 else None
                print(json.dumps({"what": tag, "n": n, "S": S, "N": N, "strict": strict, "bn": bn, "serpentine": serp,
                                  "grad_cand_per_s": r, "value_cand_per_s": rv, "chunk": eng.info()['chunk']})); sys.stdout.flush()
    eng.close()


def refit_matrix():
    d = 6
    rng = np.random.default_rng(0)
    for n in (128, 512, 1024, 2048, 3072, 4096):
        X = 1 + 9 * rng.random((n, d)); Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
        for bn in (128, 64):
            for inv in ((0,) if n < 1024 else (1, 2)):
                eng = Engine(0)
                eng.set_option('bn_fit', bn); eng.set_option('inverse', inv)
                ts = []
                for it in range(10):
                    t0 = time.perf_counter()
                    eng.fit('rbf', X, Y, [1.0 + 0.01 * it], [[2.0]], [1e-2])
                    ts.append(time.perf_counter() - t0)
                print(json.dumps({"what": "refit", "n": n, "bn_fit": bn, "inverse": inv, "fit_ms_min": min(ts[2:]) * 1e3,
                                  "fit_ms_median": float(np.median(ts[2:])) * 1e3})); sys.stdout.flush()
                eng.close()


def prefetch_ab():
    n, d, N = 2048, 6, 9472 * 32
    rng = np.random.default_rng(0)
    X = 1 + 9 * rng.random((n, d)); Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
    xs = torch.from_numpy(1 + 9 * rng.random((N, d))).cuda()
    f = torch.empty(N, dtype=torch.float64, device='cuda'); df = torch.empty((N, d), dtype=torch.float64, device='cuda')
    eng = Engine(0)
    for kernel in ('rbf', 'matern52'):
        eng.fit(kernel, X, Y, [1.0], np.full((1, d), 2.0), [1e-2])
        for rep in range(3):
            for pf in (0, 1):
                eng.set_option('prefetch_kt', pf)
                print(json.dumps({"what": "prefetch_kt", "kernel": kernel, "prefetch": pf, "rep": rep,
                                  "grad_cand_per_s": rate(eng, 'EI', xs, f, df, True, reps=4)})); sys.stdout.flush()
    eng.close()


if __name__ == '__main__':
    what = sys.argv[1:] or ['c3', 'c4', 'refit']
    if 'c3' in what:
        sweep_matrix('c3', 2048, 6, 1, 9472 * 24, 'EI')
    if 'c4' in what:
        sweep_matrix('c4', 512, 4, 64, 1024 * 96, 'EI_MCMC')
    if 'refit' in what:
        refit_matrix()
    if 'pf' in what:
        prefetch_ab()
