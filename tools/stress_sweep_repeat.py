"""The same tensor-core sweep (K1 with its cp.async.bulk staging + both triangular GEMMs + K3) repeated: every call must return the
same bits (the latency kernels' bulk-copy ring did not: DESIGN.md section 5)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from gpyopt_b200 import Engine
rng = np.random.default_rng(5)
st = torch.cuda.current_stream().cuda_stream
for (n, d, S, N, kern, reps) in ((2048, 6, 1, 18944 * 2, 'rbf', 200), (700, 5, 4, 30000, 'matern52', 200), (100, 3, 1, 200000, 'rbf', 200),
                                (512, 4, 64, 20000, 'rbf', 100), (55, 2, 1, 1000000, 'matern52', 200)):
    X = rng.random((n, d)) * 9 + 1
    Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
    Xs = torch.from_numpy(rng.random((N, d)) * 9 + 1).cuda()
    eng = Engine(0)
    eng.fit(kern, X, Y, 1.0 + 0.01 * np.arange(S), np.full((S, 1), 2.0), np.full(S, 1e-2))
    acq = 'EI' if S == 1 else 'EI_MCMC'
    f = torch.empty(N, dtype=torch.float64, device='cuda'); df = torch.empty(N, d, dtype=torch.float64, device='cuda')
    eng.acq(acq, 0.01, Xs, grad=True, out=(f, df), stream=st); torch.cuda.synchronize()
    f0, df0 = f.clone(), df.clone()
    bad = 0
    for r in range(reps):
        eng.acq(acq, 0.01, Xs, grad=True, out=(f, df), stream=st); torch.cuda.synchronize()
        bad += not (torch.equal(f, f0) and torch.equal(df, df0))
    # value-only sweeps (small models: the fused shared-memory-resident kernel)
    eng.acq(acq, 0.01, Xs, out=(f, None), stream=st); torch.cuda.synchronize()
    v0 = f.clone(); bad_v = 0
    for r in range(reps):
        eng.acq(acq, 0.01, Xs, out=(f, None), stream=st); torch.cuda.synchronize()
        bad_v += not torch.equal(f, v0)
    print(json.dumps({"n": n, "S": S, "N": N, "kernel": kern, "calls": reps, "grad_calls_differing": bad, "value_calls_differing": bad_v}), flush=True)
    eng.close()
