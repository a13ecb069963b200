"""EI + gradient sweep rate at the headline shape with the strict-variance form forced on / off."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import Engine
n, d, N = 2048, 6, 9472 * 16
rng = np.random.default_rng(0)
X = 1 + 9 * rng.random((n, d)); Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
eng = Engine(0)
xs = torch.from_numpy(1 + 9 * rng.random((N, d))).cuda()
f = torch.empty(N, dtype=torch.float64, device='cuda'); df = torch.empty((N, d), dtype=torch.float64, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for mode in (0, 1):
    eng.set_strict_variance(mode)
    eng.fit('rbf', X, Y, [1.0], [[2.0]], [1e-2])
    for _ in range(2): eng.acq('EI', 0.01, xs, grad=True, out=(f, df), stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): eng.acq('EI', 0.01, xs, grad=True, out=(f, df), stream=st)
    e1.record(); torch.cuda.synchronize()
    print("strict=%d  %.3f M cand/s" % (mode, N / (e0.elapsed_time(e1) / 3) / 1e3))
