"""Print every comparison ratio of the randomised oracle test for given cases (debugging aid for manual_random_sweep)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from conftest import scaled_err, parity_tolerances
from gpyopt_b200 import Engine
from oracle import gpy_numpy as G, gpyopt_numpy as og

cases = json.loads(sys.argv[1])
for (n, d, N, kernel, ARD, S, noise, seed) in cases:
    rng = np.random.default_rng(seed)
    X = rng.random((n, d))
    Y = og.normalize_stats(np.sin(3 * X).sum(1, keepdims=True) + 0.1 * rng.standard_normal((n, 1))) if n > 1 else np.array([[0.3]])
    Xs = rng.random((N, d))
    var = 0.5 + rng.random(S); ls = 0.3 + rng.random((S, d if ARD else 1)); nz = noise * (0.5 + rng.random(S))
    K = {'rbf': G.RBF, 'matern52': G.Matern52, 'matern32': G.Matern32}[kernel]
    gps = [G.GPRegression(X, Y, K(d, variance=var[z], lengthscale=ls[z], ARD=ARD), noise_var=nz[z]) for z in range(S)]
    eng = Engine(0); eng.set_chunk(256)
    eng.fit(kernel, X, Y, var, np.broadcast_to(ls, (S, d)), nz)
    tols = [parity_tolerances(g, Xs) for g in gps]
    tol_v, tol_g = max(t[0] for t in tols), max(t[1] for t in tols)
    ref = [og.gp_predict_withGradients(g, Xs) for g in gps]
    m, s, dm, ds = eng.predict_grad(Xs); mv, sv = eng.predict(Xs)
    out = {"case": [n, d, N, kernel, ARD, S, noise], "tol_v": tol_v, "tol_g": tol_g, "info": eng.info()}
    for z in range(S):
        nat = np.sqrt(var[z] + nz[z]) / float(np.min(ls[z]))
        out["z%d" % z] = {"m": scaled_err(m[z], ref[z][0][:, 0]), "s": scaled_err(s[z], ref[z][1][:, 0]), "mv": scaled_err(mv[z], ref[z][0][:, 0]),
                          "sv": scaled_err(sv[z], ref[z][1][:, 0]), "dm": scaled_err(dm[z], ref[z][2]),
                          "ds": float(np.max(np.abs(ds[z] - ref[z][3])) / max(nat, np.max(np.abs(ref[z][3]))))}
    fmins = [og.gp_get_fmin(g) for g in gps]
    out["fmin"] = scaled_err(eng.get_fmin(), np.array(fmins))
    means, stds = [r[0] for r in ref], [r[1].copy() for r in ref]
    if S == 1:
        fo, dfo = og.acq_EI_withGradients(means[0], stds[0], ref[0][2], ref[0][3], fmins[0], 0.01); kind = 'EI'
    else:
        fo, dfo = og.acq_EI_MCMC_withGradients(means, stds, [r[2] for r in ref], [r[3] for r in ref], fmins, 0.01); kind = 'EI_MCMC'
    f, df = eng.acq(kind, 0.01, Xs, grad=True); fv = eng.acq(kind, 0.01, Xs)
    scale_f = max(np.max(np.abs(fo)), 1e-100)
    out["f"] = float(np.max(np.abs(f - fo[:, 0])) / scale_f); out["fv"] = float(np.max(np.abs(fv - fo[:, 0])) / scale_f)
    out["df"] = float(np.max(np.abs(df - dfo)) / max(np.max(np.abs(dfo)), 1e-100)); out["max_f"] = float(np.max(np.abs(fo)))
    k = min(3, N)
    vals, idx = eng.acq_topk(kind, 0.01, Xs, k)
    out["topk_ok"] = bool(np.array_equal(np.sort(fv)[::-1][:k], -vals))
    print(json.dumps(out)); eng.close()
