"""Time of one sweep chunk's kernels (K1, GEMM, K3) from CUDA events around small sweeps: value-only LCB over 9472
candidates at n=2048 isolates K3's share by difference against the profiled GEMM time."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import Engine
n, d = 2048, 6
rng = np.random.default_rng(0)
X = 1 + 9 * rng.random((n, d)); Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
eng = Engine(0); eng.fit('rbf', X, Y, [1.0], [[2.0]], [1e-2])
for N in (9472, 9472 * 16):
    xs = torch.from_numpy(1 + 9 * rng.random((N, d))).cuda()
    f = torch.empty(N, dtype=torch.float64, device='cuda'); df = torch.empty((N, d), dtype=torch.float64, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    for grad in (True, False):
        out = (f, df if grad else None)
        for _ in range(3): eng.acq('EI', 0.01, xs, grad=grad, out=out, stream=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): eng.acq('EI', 0.01, xs, grad=grad, out=out, stream=st)
        e1.record(); torch.cuda.synchronize()
        print("N=%d grad=%s  %.3f ms per sweep  (%.3f M cand/s)" % (N, grad, e0.elapsed_time(e1) / 5, N / (e0.elapsed_time(e1) / 5) / 1e3))
