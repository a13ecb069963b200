"""Where the end-to-end (host buffers) step spends its time relative to the device-resident one."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import GPModel, AcquisitionEI, kern
n, d, N = 2048, 6, int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
rng = np.random.default_rng(0)
X = 1 + 9 * rng.random((n, d)); Y = np.sin(X).sum(1, keepdims=True); Y = (Y - Y.mean()) / Y.std()
m = GPModel(kernel=kern.RBF(d, lengthscale=2.0), noise_var=1e-2, max_iters=0, verbose=False); m.updateModel(X, Y, None, None)
acq = AcquisitionEI(m, None, None, jitter=0.01)
eng = m.model.engine
xs_d = torch.from_numpy(1 + 9 * rng.random((N, d))).cuda()
f_d = torch.empty(N, dtype=torch.float64, device='cuda'); df_d = torch.empty((N, d), dtype=torch.float64, device='cuda')
hx = torch.empty((N, d), dtype=torch.float64, pin_memory=True); hx.copy_(xs_d); x_np = hx.numpy()
hf = torch.empty(N, dtype=torch.float64, pin_memory=True).numpy(); hdf = torch.empty((N, d), dtype=torch.float64, pin_memory=True).numpy()
x_pageable = x_np.copy()


def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best


st = torch.cuda.current_stream().cuda_stream
a = t(lambda: eng.acq('EI', 0.01, xs_d, grad=True, out=(f_d, df_d), stream=st))
b = t(lambda: eng.acq('EI', 0.01, x_np, grad=True, out=(hf, hdf)))
c = t(lambda: eng.acq('EI', 0.01, x_np, grad=True))
e = t(lambda: acq._compute_acq_withGradients(x_np))
g = t(lambda: acq._compute_acq_withGradients(x_pageable))
print("N=%d  device-resident %.1f ms | host pinned in+out (preallocated) %.1f | host, outputs allocated by the engine %.1f | plugin %.1f | plugin, pageable input %.1f"
      % (N, a * 1e3, b * 1e3, c * 1e3, e * 1e3, g * 1e3))
