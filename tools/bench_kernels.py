"""Headline shape (n=2048, d=6, EI+grad and value-only sweeps) for every kernel family / ARD, device-resident."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from gpyopt_b200 import Engine
rng = np.random.default_rng(0)
st = torch.cuda.current_stream().cuda_stream
n, d, N = 2048, 6, 9472 * 16
X = rng.random((n, d)) * 9 + 1
fv = -(np.prod(np.sqrt(X), 1) * np.prod(np.sin(X), 1)); Y = (fv - fv.mean()) / fv.std()
Xs = torch.from_numpy(rng.random((N, d)) * 9 + 1).cuda()
f = torch.empty(N, dtype=torch.float64, device='cuda'); df = torch.empty(N, d, dtype=torch.float64, device='cuda')
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for kern, ls in (('rbf', [[2.0] * d]), ('rbf', [[1.5, 1.8, 2.1, 2.4, 2.7, 3.0]]), ('matern52', [[2.5] * d]), ('matern32', [[2.5] * d])):
    eng = Engine(0); eng.fit(kern, X, Y, [1.0], ls, [1e-2])
    for acq, par in (('EI', 0.01), ('MPI', 0.01), ('LCB', 2.0)):
        tg = t(lambda: eng.acq(acq, par, Xs, grad=True, out=(f, df), stream=st)); tv = t(lambda: eng.acq(acq, par, Xs, out=(f, None), stream=st))
        print(json.dumps({'kernel': kern, 'ard': len(set(ls[0])) > 1, 'acq': acq, 'grad_Mcand_s': round(N / tg / 1e3, 3), 'value_Mcand_s': round(N / tv / 1e3, 3),
                          'strict': eng.info()['strict_variance']}))
    eng.close()
