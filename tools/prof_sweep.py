"""Short sweep for ncu: fit n=2048,d=6 then EI+grad and EI value-only over a few chunks."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from gpyopt_b200._lib import Engine
n, d = 2048, 6
N = int(sys.argv[1]) if len(sys.argv) > 1 else 9472 * 2
mode = sys.argv[2] if len(sys.argv) > 2 else 'both'
rng = np.random.default_rng(0)
X = 1 + 9 * rng.random((n, d)); Y = np.sin(X).prod(1); Y = (Y - Y.mean()) / Y.std()
eng = Engine(0); eng.fit('rbf', X, Y, [1.0], [[2.0]], [1e-2])
Xs = torch.from_numpy(1 + 9 * rng.random((N, d))).cuda()
f = torch.empty(N, dtype=torch.float64, device='cuda'); df = torch.empty(N, d, dtype=torch.float64, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for it in range(2):
    if mode in ('both', 'grad'):
        eng.acq('EI', 0.01, Xs, grad=True, out=(f, df), stream=st)
    if mode in ('both', 'value'):
        eng.acq('EI', 0.01, Xs, grad=False, out=(f, None), stream=st)
torch.cuda.synchronize()
print('ok', float(f.max()))
