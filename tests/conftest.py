import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` on the GPU box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def scaled_err(a, ref):
    """max|a - ref| / max|ref|: the parity metric of SURVEY.md section 8(d)."""
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    den = np.max(np.abs(ref)) if ref.size else 1.0
    return float(np.max(np.abs(a - ref)) / (den if den > 0 else 1.0)) if ref.size else 0.0


GP_CASES = ['c3_rbf_alpine2_n256', 'c3_rbf_ard_exact_n200', 'c3_mat52_alpine2_n192', 'c3_mat32_ard_n130',
            'c1_branin_rbf_n20', 'c2_camel_mat52_n55', 'c1_1d_rbf_n7', 'c5_lp_gsobol_n128_gp']
MCMC_CASES = ['c4_mcmc_n64_S8', 'c4_mcmc_exact_n48_S4']


def oracle_gp_from_golden(g):
    """Rebuild the oracle GP (oracle.gpy_numpy) with the golden case's data and fixed hyper-parameters."""
    from oracle import gpy_numpy as G
    K = {'rbf': G.RBF, 'matern52': G.Matern52, 'matern32': G.Matern32}[str(g['kernel'])]
    d = g['X'].shape[1]
    kern = K(d, variance=float(g['variance']), lengthscale=g['lengthscale'], ARD=bool(g['ARD']))
    return G.GPRegression(g['X'], g['Y'], kern, noise_var=float(g['noise_var']))


@pytest.fixture(scope='session')
def engine():
    from gpyopt_b200 import Engine
    eng = Engine(0)
    yield eng
    eng.close()
