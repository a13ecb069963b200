import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import torch
from conftest import load_golden
from gpyopt_b200 import Engine
g = load_golden('c3_rbf_alpine2_n256')
eng = Engine(0)
d = 6
eng.fit('rbf', g['X'], g['Y'], [float(g['variance'])], [[float(g['lengthscale'][0])] * d], [float(g['noise_var'])])
Xs = g['Xs']
f1, df1 = eng.acq('EI', 0.01, Xs, grad=True)
f2, df2 = eng.acq('EI', 0.01, Xs, grad=True)
print('host twice equal', np.array_equal(f1, f2), np.array_equal(df1, df2))
Xd = torch.from_numpy(Xs).cuda()
f3, df3 = eng.acq('EI', 0.01, Xd, grad=True, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
f3 = f3.cpu().numpy(); df3 = df3.cpu().numpy()
print('dev vs host equal', np.array_equal(f1, f3), np.array_equal(df1, df3), np.abs(f1 - f3).max(), np.nonzero(f1 != f3)[0][:10])
f4, df4 = eng.acq('EI', 0.01, Xd, grad=True, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print('dev twice equal', np.array_equal(f3, f4.cpu().numpy()))
print(f1[:6], f3[:6])

# BO loop debug
import scipy.optimize
from gpyopt_b200 import GPModel, AcquisitionEI, kern
from oracle import gpyopt_numpy as og
import gpyopt_b200.gp as gpmod
from fake_engine import FakeEngine
for EngCls in (Engine, FakeEngine):
    gpmod.Engine = EngCls
    np.random.seed(123)
    bounds = [(-5., 10.), (1., 15.)]
    lo, hi = np.array(bounds).T
    X = lo + (hi - lo) * np.random.rand(5, 2)
    Yraw = og.branin(X)
    model = GPModel(kernel=kern.RBF(2), exact_feval=True, optimize_restarts=3, verbose=False)
    for it in range(15):
        model.updateModel(X, og.normalize_stats(Yraw), None, None)
        acq = AcquisitionEI(model, None, None, jitter=0.01)
        cand = lo + (hi - lo) * np.random.rand(1000, 2)
        _, idx = model.model.engine.acq_topk('EI', 0.01, cand, 5)
        best = (np.inf, None)
        for a in cand[idx]:
            res = scipy.optimize.fmin_l_bfgs_b(
                lambda x: (lambda f, df: (float(f[0, 0]), df[0]))(*acq.acquisition_function_withGradients(np.atleast_2d(x))),
                x0=a, bounds=bounds, maxiter=200)
            if res[1] < best[0]:
                best = (res[1], res[0])
        X = np.vstack([X, best[1]])
        Yraw = np.vstack([Yraw, og.branin(best[1])])
        print(EngCls.__name__, it, model.get_model_parameters().ravel(), best[1], float(og.branin(best[1])), best[0])
    print(EngCls.__name__, 'min', Yraw.min())
