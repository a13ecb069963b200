import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from gpyopt_b200 import Engine
rng = np.random.default_rng(0)
n, d, N = 2048, 6, 9472
X = 1 + 9 * rng.random((n, d)); Y = np.sin(X).prod(1); Y = (Y - Y.mean()) / Y.std()
eng = Engine(0); eng.fit('rbf', X, Y, [1.0], [[2.0]], [1e-2])
Xs = torch.from_numpy(1 + 9 * rng.random((N, d))).cuda()
f = torch.empty(N, dtype=torch.float64, device='cuda'); df = torch.empty(N, d, dtype=torch.float64, device='cuda')
for it in range(2):
    eng.acq('EI', 0.01, Xs, grad=True, out=(f, df), stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize(); print('ok')
