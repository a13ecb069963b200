import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from gpyopt_b200 import Engine
rng = np.random.default_rng(0)
n, d, S, N = 512, 4, 64, 4096
X = rng.random((n, d)); Y = np.sin(5*X).sum(1); Y = (Y - Y.mean()) / Y.std()
theta = np.array([1.0, 0.5, 1e-2]) * np.exp(0.1 * np.random.default_rng(2).standard_normal((S, 3)))
eng = Engine(0); eng.fit('rbf', X, Y, theta[:, 0], theta[:, 1:2], theta[:, 2])
Xs = torch.from_numpy(rng.random((N, d))).cuda()
f = torch.empty(N, dtype=torch.float64, device='cuda'); df = torch.empty(N, d, dtype=torch.float64, device='cuda')
for it in range(2):
    eng.acq('EI_MCMC', 0.01, Xs, grad=True, out=(f, df), stream=torch.cuda.current_stream().cuda_stream)
    eng.acq('EI_MCMC', 0.01, Xs, grad=False, out=(f, None), stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize(); print('ok', eng.info())
