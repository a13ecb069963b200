"""Short run of the fused small-model value sweep for ncu (C2 shape: n=55, d=2, Matern52, LCB)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from gpyopt_b200 import Engine
rng = np.random.default_rng(0)
n, d, N = 55, 2, 1 << 18
lo, hi = np.array([-2., -1.]), np.array([2., 1.])
X = lo + (hi - lo) * rng.random((n, d)); Y = np.sin(3 * X).sum(1); Y = (Y - Y.mean()) / Y.std()
eng = Engine(0); eng.fit('matern52', X, Y, [1.0], [[0.7, 0.7]], [1e-2])
Xs = torch.from_numpy(lo + (hi - lo) * rng.random((N, d))).cuda(); f = torch.empty(N, dtype=torch.float64, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    eng.acq('LCB', 2.0, Xs, out=(f, None), stream=st)
torch.cuda.synchronize()
print('ok', float(f.max()))
