"""A/B of GPO_OPT_DESYNC (co-resident GEMM CTAs walk their items in opposite orders) on the headline shape and on C4's shape:
candidate rate with gradients / values only, and bitwise equality of the outputs."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from gpyopt_b200 import Engine
rng = np.random.default_rng(0)
st = torch.cuda.current_stream().cuda_stream
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for (n, d, S, N, kern) in ((2048, 6, 1, 9472 * 16, 'rbf'), (2048, 6, 1, 9472 * 16, 'matern52'), (512, 6, 64, 200000, 'rbf')):
    X = rng.random((n, d)) * 9 + 1
    fv = -(np.prod(np.sqrt(X), 1) * np.prod(np.sin(X), 1)); Y = (fv - fv.mean()) / fv.std()
    Xs = torch.from_numpy(rng.random((N, d)) * 9 + 1).cuda()
    outs = {}
    for rep in range(2):
        for desync in (0, 1):
            eng = Engine(0); eng.set_option('desync', desync)
            eng.fit(kern, X, Y, 1.0 + 0.01 * np.arange(S), np.full((S, 1), 2.0) + 0.01 * np.arange(S)[:, None], np.full(S, 1e-2))
            acq = 'EI' if S == 1 else 'EI_MCMC'
            f = torch.empty(N, dtype=torch.float64, device='cuda'); df = torch.empty(N, d, dtype=torch.float64, device='cuda')
            tg = t(lambda: eng.acq(acq, 0.01, Xs, grad=True, out=(f, df), stream=st))
            tv = t(lambda: eng.acq(acq, 0.01, Xs, out=(f, None), stream=st))
            eng.acq(acq, 0.01, Xs, grad=True, out=(f, df), stream=st); torch.cuda.synchronize()
            outs[desync] = (f.cpu().numpy().copy(), df.cpu().numpy().copy())
            print(json.dumps({'n': n, 'S': S, 'kernel': kern, 'desync': desync, 'rep': rep, 'grad_Mcand_s': round(N / tg / 1e3, 4),
                              'value_Mcand_s': round(N / tv / 1e3, 4)}), flush=True)
            eng.close()
    print(json.dumps({'n': n, 'S': S, 'kernel': kern, 'bitwise_equal': bool(np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]))}), flush=True)
