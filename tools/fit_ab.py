"""A/B of the two refit paths (gpo_set_option fit_path 0 = short critical chain, 1 = round-1 look-ahead scheme): state agreement
(L, L^-1, W^-1, alpha, log-likelihood; scaled max differences and the residuals |L L^T - Ky|, |L^-1 L - I|, |W^-1 Ky - I| of each
path by itself), bit-reproducibility of the new path, and refit latency of both.

    python tools/fit_ab.py [n ...] [--S k] [--no-check]
"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import Engine

args = [a for a in sys.argv[1:]]
S = 1
check = True
ns = []
i = 0
while i < len(args):
    if args[i] == '--S':
        S = int(args[i + 1]); i += 2
    elif args[i] == '--no-check':
        check = False; i += 1
    else:
        ns.append(int(args[i])); i += 1
ns = ns or [200, 256, 512, 1000, 2048, 4096]
d = 6
rng = np.random.default_rng(0)


def rbf(X, var, ls):
    Z = X / ls
    sq = (Z * Z).sum(1)
    return var * np.exp(-0.5 * np.maximum(sq[:, None] + sq[None, :] - 2 * Z @ Z.T, 0.0))


for n in ns:
    X = 1 + 9 * rng.random((n, d)); Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
    var = 1.0 + 0.1 * np.arange(S); ls = np.full((S, 1), 2.0) + 0.05 * np.arange(S)[:, None]; noise = np.full(S, 1e-2)
    out = {"n": n, "S": S}
    states = {}
    for path in (1, 0):
        eng = Engine(0)
        eng.set_option('fit_path', path)
        ts = []
        for it in range(10):
            t0 = time.perf_counter()
            eng.fit('rbf', X, Y, var * (1 + 0.01 * it), ls, noise)
            ts.append(time.perf_counter() - t0)
        out["fit_ms_path%d" % path] = round(min(ts[2:]) * 1e3, 4)
        if check:
            eng.fit('rbf', X, Y, var, ls, noise)
            st = eng.get_state()
            ll = eng.get_loglik()[0]
            if path == 0:   # bit-reproducibility: a second fit of the same problem
                eng.fit('rbf', X, Y, var * 1.5, ls, noise)
                eng.fit('rbf', X, Y, var, ls, noise)
                st2 = eng.get_state()
                out["bitwise_repeat"] = bool(all(np.array_equal(a, b) for a, b in zip(st, st2)))
            states[path] = (st, ll)
        eng.close()
    if check:
        (L1, X1, W1, a1), ll1 = states[1]
        (L0, X0, W0, a0), ll0 = states[0]
        sc = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
        out.update(dL=sc(L0, L1), dLinv=sc(X0, X1), dWinv=sc(W0, W1), dalpha=sc(a0, a1), dloglik=float(np.max(np.abs(ll0 - ll1) / np.abs(ll1))))
        if n <= 2048:
            z = S - 1
            Ky = rbf(X, var[z], ls[z]) + (noise[z] + 1e-8) * np.eye(n)
            for name, (L, Xi, W) in (("new", (L0[z], X0[z], W0[z])), ("old", (L1[z], X1[z], W1[z]))):
                out["res_chol_" + name] = float(np.max(np.abs(L @ L.T - Ky)))
                out["res_inv_" + name] = float(np.max(np.abs(Xi @ L - np.eye(n))))
                out["res_w_" + name] = float(np.max(np.abs(W @ Ky - np.eye(n))))
                out["upper_zero_" + name] = bool(np.all(np.triu(L, 1) == 0) and np.all(np.triu(Xi, 1) == 0))
                out["w_sym_" + name] = bool(np.array_equal(W, W.T))
    print(json.dumps(out), flush=True)
