"""Timeline of one refit (GPO_OPT_FIT_TRACE): every launch of factorize_chain with its start / end on its stream -> stderr."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1
d = 6
rng = np.random.default_rng(0)
X = 1 + 9 * rng.random((n, d)); Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
eng = Engine(0)
for it in range(4):
    if it == 3:
        eng.set_option('fit_trace', 1)
    eng.fit('rbf', X, Y, np.full(S, 1.0 + 0.01 * it), np.full((S, 1), 2.0), np.full(S, 1e-2))
