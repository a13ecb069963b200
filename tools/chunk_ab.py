"""Headline shape (n=2048, d=6, EI + gradient): candidate rate against the chunk size (candidates per GEMM launch)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from gpyopt_b200 import Engine
rng = np.random.default_rng(0)
st = torch.cuda.current_stream().cuda_stream
def t(fn, reps=4):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
n, d, N = 2048, 6, 9472 * 32
X = rng.random((n, d)) * 9 + 1
fv = -(np.prod(np.sqrt(X), 1) * np.prod(np.sin(X), 1)); Y = (fv - fv.mean()) / fv.std()
Xs = torch.from_numpy(rng.random((N, d)) * 9 + 1).cuda()
for kern in ('rbf', 'matern52'):
    for chunk in (9472, 2 * 9472, 4 * 9472, 8 * 9472):
        eng = Engine(0); eng.set_chunk(chunk)
        eng.fit(kern, X, Y, [1.0], [[2.0]], [1e-2])
        f = torch.empty(N, dtype=torch.float64, device='cuda'); df = torch.empty(N, d, dtype=torch.float64, device='cuda')
        tg = t(lambda: eng.acq('EI', 0.01, Xs, grad=True, out=(f, df), stream=st))
        tv = t(lambda: eng.acq('EI', 0.01, Xs, out=(f, None), stream=st))
        print(json.dumps({'kernel': kern, 'chunk': eng.info()['chunk'], 'grad_Mcand_s': round(N / tg / 1e3, 4), 'value_Mcand_s': round(N / tv / 1e3, 4),
                          'checksum': float(f.sum().item())}), flush=True)
        eng.close()
