import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from gpyopt_b200 import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = 10
rng = np.random.default_rng(0)
X = -4 + 10 * rng.random((n, d)); Y = np.sin(X).sum(1); Y = (Y - Y.mean()) / Y.std()
eng = Engine(0)
eng.fit('rbf', X, Y, [1.0], [[3.0] * d], [1e-2])
x = -4 + 10 * rng.random((1, d))
for _ in range(4):
    f, df = eng.acq('EI', 0.01, x, grad=True)
print(f, df[0, :3])
