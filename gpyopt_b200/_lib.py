"""ctypes binding of the C ABI declared in include/gpo_b200.h (libgpo_b200.so, sm_100a only).

There is no CPU fallback: if the shared library is missing or no B200 is present the constructor of
:class:`Engine` raises.  NumPy arrays (host pointers) and torch CUDA tensors (device pointers) are both
accepted wherever the header says (host|device).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libgpo_b200.so')

GPO_OK, GPO_ERR_CUDA, GPO_ERR_NOT_PD, GPO_ERR_ARG, GPO_ERR_STATE, GPO_ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5
KERNELS = {'rbf': 0, 'matern52': 1, 'mat52': 1, 'matern32': 2, 'mat32': 2}
ACQ = {'EI': 0, 'MPI': 1, 'LCB': 2, 'EI_MCMC': 3, 'MPI_MCMC': 4, 'LCB_MCMC': 5}
LP_TRANSFORM = {'none': 0, 'softplus': 1}

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_ll_p = ctypes.POINTER(ctypes.c_longlong)

# symbol -> (restype, argtypes); kept in one table so tests can check the header against the library
SIGNATURES = {
    'gpo_strerror': (ctypes.c_char_p, [ctypes.c_int]),
    'gpo_last_error': (ctypes.c_char_p, [ctypes.c_void_p]),
    'gpo_version': (ctypes.c_char_p, []),
    'gpo_create': (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    'gpo_destroy': (ctypes.c_int, [ctypes.c_void_p]),
    'gpo_set_chunk': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong]),
    'gpo_fit': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                               ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'gpo_get_fmin': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'gpo_get_loglik': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'gpo_get_loglik_grad': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    'gpo_get_state': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'gpo_predict': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p]),
    'gpo_predict_grad': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'gpo_predict_mean_grad': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p]),
    'gpo_acq': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_longlong, ctypes.c_void_p,
                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'gpo_lp_set': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_int]),
    'gpo_acq_topk': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_longlong, ctypes.c_void_p,
                                    ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'gpo_design_uniform': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'gpo_acq_topk_uniform': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_longlong,
                                            ctypes.c_longlong, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p]),
    'gpo_topk': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'gpo_lbfgs': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_void_p]),
    'gpo_posterior_cov': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_void_p]),
    'gpo_set_strict_variance': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    'gpo_get_info': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    'gpo_set_option': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong]),
    'gpo_set_profiling': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    'gpo_get_profile': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'gpo_alloc': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    'gpo_state_buffers': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    'gpo_state_commit': (ctypes.c_int, [ctypes.c_void_p]),
    'gpo_launch_count': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
}

_lib = None


def load_library():
    """dlopen libgpo_b200.so and attach the prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not found: build it with `python -m gpyopt_b200.build` (nvcc, sm_100a). "
                          "This package has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class GPOError(RuntimeError):
    pass


def _ptr(a):
    """Raw address of a NumPy array (host) or a torch tensor (host or device); None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float64 or a.dtype == np.int64, a.dtype
        assert a.flags['C_CONTIGUOUS']
        return a.ctypes.data or None   # (zero-size arrays have no buffer)
    # torch tensor
    assert a.is_contiguous()
    return a.data_ptr()


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class Engine(object):
    """One posterior state (data + S hyper-parameter sets) resident on one GPU."""

    def __init__(self, device=0):
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        rc = self._lib.gpo_create(int(device), ctypes.byref(self._h))
        if rc != GPO_OK:
            msg = self._lib.gpo_last_error(None).decode()
            raise GPOError("gpo_create(device=%d) failed: %s (%s)" % (device, self._lib.gpo_strerror(rc).decode(), msg))
        self.device = int(device)
        self.n = self.d = self.S = 0

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self._lib.gpo_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc == GPO_OK:
            return
        msg = self._lib.gpo_last_error(self._h).decode()
        if rc == GPO_ERR_NOT_PD:
            # same exception type GPy's jitchol raises; GPyOpt/core/bo.py:143 catches it
            raise np.linalg.LinAlgError("not positive definite, even with jitter.")
        raise GPOError("%s failed: %s (%s)" % (what, self._lib.gpo_strerror(rc).decode(), msg))

    # -- fit -------------------------------------------------------------------------------------
    def set_chunk(self, chunk):
        self._check(self._lib.gpo_set_chunk(self._h, int(chunk)), 'gpo_set_chunk')

    def fit(self, kernel, X, Y, variance, lengthscale, noise):
        """variance [S], lengthscale [S, d] (or [S, 1] / [S] for non-ARD, broadcast), noise [S]."""
        X = _f64(X)
        Y = _f64(Y).reshape(-1)
        n, d = X.shape
        variance = _f64(variance).reshape(-1)
        S = variance.size
        ls = _f64(lengthscale).reshape(S, -1)
        if ls.shape[1] == 1:
            ls = np.repeat(ls, d, axis=1)
        ls = _f64(ls)
        noise = _f64(noise).reshape(-1)
        assert ls.shape == (S, d) and noise.size == S and Y.size == n
        k = KERNELS[kernel.lower()] if isinstance(kernel, str) else int(kernel)
        rc = self._lib.gpo_fit(self._h, k, n, d, _ptr(X), _ptr(Y), S, _ptr(variance), _ptr(ls), _ptr(noise))
        self._check(rc, 'gpo_fit')
        self.n, self.d, self.S = n, d, S

    def get_fmin(self):
        out = np.empty(self.S)
        self._check(self._lib.gpo_get_fmin(self._h, _ptr(out)), 'gpo_get_fmin')
        return out

    def get_loglik(self):
        ll, ld = np.empty(self.S), np.empty(self.S)
        self._check(self._lib.gpo_get_loglik(self._h, _ptr(ll), _ptr(ld)), 'gpo_get_loglik')
        return ll, ld

    def get_loglik_grad(self):
        g = np.empty((self.S, self.d + 2))
        self._check(self._lib.gpo_get_loglik_grad(self._h, _ptr(g)), 'gpo_get_loglik_grad')
        return g

    def get_state(self, L=True, Linv=True, Winv=True, alpha=True):
        S, n = self.S, self.n
        out = [np.empty((S, n, n)) if f else None for f in (L, Linv, Winv)] + [np.empty((S, n)) if alpha else None]
        self._check(self._lib.gpo_get_state(self._h, *[_ptr(o) for o in out]), 'gpo_get_state')
        return out

    # -- sweeps ----------------------------------------------------------------------------------
    @staticmethod
    def _rows(Xs, d):
        if isinstance(Xs, np.ndarray):
            assert Xs.ndim == 2 and Xs.shape[1] == d
            return Xs.shape[0]
        assert Xs.dim() == 2 and Xs.shape[1] == d
        return int(Xs.shape[0])

    def predict(self, Xs, with_noise=True, out=None, stream=None):
        """-> m, s of shape [S, N] (NumPy for NumPy input, torch CUDA tensors for torch CUDA input)."""
        Xs = self._as_input(Xs)
        N = self._rows(Xs, self.d)
        m, s = out if out is not None else (self._empty_like(Xs, (self.S, N)), self._empty_like(Xs, (self.S, N)))
        rc = self._lib.gpo_predict(self._h, N, _ptr(Xs), int(bool(with_noise)), _ptr(m), _ptr(s), stream)
        self._check(rc, 'gpo_predict')
        return m, s

    def predict_grad(self, Xs, out=None, stream=None):
        Xs = self._as_input(Xs)
        N = self._rows(Xs, self.d)
        if out is None:
            out = (self._empty_like(Xs, (self.S, N)), self._empty_like(Xs, (self.S, N)),
                   self._empty_like(Xs, (self.S, N, self.d)), self._empty_like(Xs, (self.S, N, self.d)))
        rc = self._lib.gpo_predict_grad(self._h, N, _ptr(Xs), *[_ptr(o) for o in out], stream)
        self._check(rc, 'gpo_predict_grad')
        return out

    def predict_mean_grad(self, Xs, out=None, stream=None):
        """-> m [S, N], dmdx [S, N, d]: posterior mean and its gradient only (no variance contraction)."""
        Xs = self._as_input(Xs)
        N = self._rows(Xs, self.d)
        if out is None:
            out = (self._empty_like(Xs, (self.S, N)), self._empty_like(Xs, (self.S, N, self.d)))
        rc = self._lib.gpo_predict_mean_grad(self._h, N, _ptr(Xs), _ptr(out[0]), _ptr(out[1]), stream)
        self._check(rc, 'gpo_predict_mean_grad')
        return out

    def acq(self, kind, par, Xs, grad=False, out=None, stream=None):
        """-> f [N] (and df [N, d] when grad)."""
        Xs = self._as_input(Xs)
        N = self._rows(Xs, self.d)
        if out is None:
            out = (self._empty_like(Xs, (N,)), self._empty_like(Xs, (N, self.d)) if grad else None)
        f, df = out
        k = ACQ[kind] if isinstance(kind, str) else int(kind)
        rc = self._lib.gpo_acq(self._h, k, float(par), N, _ptr(Xs), _ptr(f), _ptr(df) if grad else None, stream)
        self._check(rc, 'gpo_acq')
        return (f, df) if grad else f

    def lp_set(self, X_batch, r, s, transform):
        if transform is None:
            self._check(self._lib.gpo_lp_set(self._h, 0, None, None, None, -1), 'gpo_lp_set')
            return
        t = LP_TRANSFORM[transform] if isinstance(transform, str) else int(transform)
        if X_batch is None:
            self._check(self._lib.gpo_lp_set(self._h, 0, None, None, None, t), 'gpo_lp_set')
            return
        Xb, r, s = _f64(np.atleast_2d(X_batch)), _f64(r).reshape(-1), _f64(s).reshape(-1)
        self._check(self._lib.gpo_lp_set(self._h, Xb.shape[0], _ptr(Xb), _ptr(r), _ptr(s), t), 'gpo_lp_set')

    def acq_topk(self, kind, par, Xs, k, f_out=None, stream=None):
        Xs = self._as_input(Xs)
        N = self._rows(Xs, self.d)
        vals = np.empty(k)
        idx = np.empty(k, dtype=np.int64)
        kk = ACQ[kind] if isinstance(kind, str) else int(kind)
        rc = self._lib.gpo_acq_topk(self._h, kk, float(par), N, _ptr(Xs), int(k), _ptr(vals), _ptr(idx), _ptr(f_out), stream)
        self._check(rc, 'gpo_acq_topk')
        kk = min(k, N)
        return vals[:kk], idx[:kk]

    @staticmethod
    def _design_args(key, bounds):
        key = np.atleast_1d(np.asarray(key, dtype=np.uint64)).reshape(-1)
        if key.size == 1:                    # numpy.random.Philox(key=int): the integer is the low word
            key = np.array([key[0], 0], dtype=np.uint64)
        if key.size != 2:
            raise ValueError("Philox key: one integer or two 64-bit words")
        bounds = _f64(np.atleast_2d(bounds))
        if bounds.shape[1] != 2:
            raise ValueError("bounds must be [d, 2] (low, high)")
        return np.ascontiguousarray(key), bounds

    def design_uniform(self, key, N, bounds, row0=0, out=None, stream=None):
        """Rows [row0, row0+N) of numpy.random.Generator(numpy.random.Philox(key=key)).uniform(lo, hi, (rows, d)),
        generated on the device (bit-identical to NumPy).  out: optional [N, d] NumPy array or device tensor."""
        key, bounds = self._design_args(key, bounds)
        d = bounds.shape[0]
        if out is None:
            out = np.empty((int(N), d))
        rc = self._lib.gpo_design_uniform(self._h, key.ctypes.data, int(row0), int(N), d, _ptr(bounds), _ptr(out), stream)
        self._check(rc, 'gpo_design_uniform')
        return out

    def acq_topk_uniform(self, kind, par, key, N, bounds, k, row0=0, stream=None):
        """acq_topk over a device-generated uniform design -> (scores [k], local idx [k], rows [k, d])."""
        key, bounds = self._design_args(key, bounds)
        if bounds.shape[0] != self.d:
            raise ValueError("bounds must have one row per model dimension (%d)" % self.d)
        vals = np.empty(k)
        idx = np.empty(k, dtype=np.int64)
        xt = np.empty((k, self.d))
        kk = ACQ[kind] if isinstance(kind, str) else int(kind)
        rc = self._lib.gpo_acq_topk_uniform(self._h, kk, float(par), key.ctypes.data, int(row0), int(N), _ptr(bounds), int(k),
                                            _ptr(vals), _ptr(idx), _ptr(xt), stream)
        self._check(rc, 'gpo_acq_topk_uniform')
        kk = min(k, int(N))
        return vals[:kk], idx[:kk], xt[:kk]

    def topk(self, v, k, largest=False, stream=None, out=None):
        if isinstance(v, np.ndarray):
            v = _f64(v).reshape(-1)
            N = v.size
        else:
            N = int(v.numel())
        if out is not None:   # device tensors (float64 [k], int64 [k]): records stay in HBM, no synchronisation
            vals, idx = out
            rc = self._lib.gpo_topk(self._h, N, _ptr(v), int(k), int(bool(largest)), _ptr(vals), _ptr(idx), stream)
            self._check(rc, 'gpo_topk')
            return vals, idx
        vals = np.empty(k)
        idx = np.empty(k, dtype=np.int64)
        rc = self._lib.gpo_topk(self._h, N, _ptr(v), int(k), int(bool(largest)), _ptr(vals), _ptr(idx), stream)
        self._check(rc, 'gpo_topk')
        kk = min(k, N)
        return vals[:kk], idx[:kk]

    def lbfgs(self, kind, par, x0, bounds, maxiter=1000, pgtol=1e-5, factr=1e7):
        """Batched multi-start L-BFGS on the device (gpo_lbfgs): minimise acquisition_function = -acq (or the LP value when
        lp_set is active) from the rows of x0 inside the box `bounds` [d, 2].  -> (x [M, d], f [M], iterations [M], evaluations [M])."""
        x0 = _f64(np.atleast_2d(x0))
        bounds = _f64(np.atleast_2d(bounds))
        M, d = x0.shape
        assert d == self.d and bounds.shape == (d, 2)
        x, f = np.empty((M, d)), np.empty(M)
        it, nf = np.zeros(M, dtype=np.int32), np.zeros(M, dtype=np.int32)
        k = ACQ[kind] if isinstance(kind, str) else int(kind)
        rc = self._lib.gpo_lbfgs(self._h, k, float(par), M, _ptr(x0), _ptr(bounds), int(maxiter), float(pgtol), float(factr),
                                 _ptr(x), _ptr(f), it.ctypes.data, nf.ctypes.data)
        self._check(rc, 'gpo_lbfgs')
        return x, f, it, nf

    def posterior_cov(self, X1, X2=None, with_noise=False):
        """Full posterior covariance [N1, N2] (X2 None: between X1 and itself, noise on the diagonal if with_noise)."""
        X1 = _f64(np.atleast_2d(X1))
        X2c = None if X2 is None else _f64(np.atleast_2d(X2))
        out = np.empty((X1.shape[0], X1.shape[0] if X2c is None else X2c.shape[0]))
        rc = self._lib.gpo_posterior_cov(self._h, X1.shape[0], _ptr(X1), 0 if X2c is None else X2c.shape[0], _ptr(X2c),
                                         int(bool(with_noise)), _ptr(out))
        self._check(rc, 'gpo_posterior_cov')
        return out

    def set_strict_variance(self, mode):
        """1 the reference's variance form always (default), 0 fast form, -1 automatic: see gpo_set_strict_variance in
        include/gpo_b200.h."""
        self._check(self._lib.gpo_set_strict_variance(self._h, int(mode)), 'gpo_set_strict_variance')

    OPTIONS = {'serpentine': 0, 'bn_sweep': 1, 'bn_fit': 2, 'inverse': 3, 'fused_small': 4, 'prefetch_kt': 5, 'fit_path': 6, 'fit_trace': 7, 'fit_graph': 8, 'fit_chain_max': 9, 'desync': 10, 'k1_slice': 11}

    def set_option(self, key, value):
        """Tuning knobs for measurements (gpo_set_option in include/gpo_b200.h)."""
        k = self.OPTIONS[key] if isinstance(key, str) else int(key)
        self._check(self._lib.gpo_set_option(self._h, k, int(value)), 'gpo_set_option')

    def info(self):
        v = np.zeros(6)
        self._check(self._lib.gpo_get_info(self._h, _ptr(v), 6), 'gpo_get_info')
        return {'strict_variance': bool(v[0]), 'variance_discrepancy': float(v[1]), 'chunk': int(v[2]), 'n_pad': int(v[3]),
                'sm_count': int(v[4]), 'S': int(v[5])}

    # -- multi-GPU state replication ---------------------------------------------------------------
    def alloc(self, kernel, n, d, S):
        k = KERNELS[kernel.lower()] if isinstance(kernel, str) else int(kernel)
        self._check(self._lib.gpo_alloc(self._h, k, int(n), int(d), int(S)), 'gpo_alloc')
        self.n, self.d, self.S = int(n), int(d), int(S)

    def state_buffers(self):
        """[(device_pointer, nbytes), ...] of every buffer a sweep reads."""
        ptrs = (ctypes.c_void_p * 32)()
        nbytes = (ctypes.c_longlong * 32)()
        cnt = self._lib.gpo_state_buffers(self._h, ptrs, nbytes, 32)
        if cnt < 0:
            self._check(cnt, 'gpo_state_buffers')
        return [(int(ptrs[i]), int(nbytes[i])) for i in range(cnt)]

    def state_commit(self):
        self._check(self._lib.gpo_state_commit(self._h), 'gpo_state_commit')

    def set_profiling(self, enable):
        self._check(self._lib.gpo_set_profiling(self._h, int(bool(enable))), 'gpo_set_profiling')

    def get_profile(self):
        """-> (gemm_ms, gemm_launches, gemm_candidates) accumulated since the last call (synchronises)."""
        ms, nl, nc = ctypes.c_double(0), ctypes.c_longlong(0), ctypes.c_longlong(0)
        self._check(self._lib.gpo_get_profile(self._h, ctypes.byref(ms), ctypes.byref(nl), ctypes.byref(nc)), 'gpo_get_profile')
        return ms.value, nl.value, nc.value

    def launch_count(self, reset=False):
        c = ctypes.c_longlong(0)
        self._check(self._lib.gpo_launch_count(self._h, ctypes.byref(c), int(reset)), 'gpo_launch_count')
        return c.value

    # -- helpers ---------------------------------------------------------------------------------
    @staticmethod
    def _as_input(Xs):
        if isinstance(Xs, np.ndarray) or not hasattr(Xs, 'data_ptr'):
            Xs = _f64(Xs)
            if Xs.ndim == 1:
                Xs = Xs[None, :]
            return Xs
        import torch
        assert Xs.dtype == torch.float64
        return Xs.contiguous()

    @staticmethod
    def _empty_like(Xs, shape):
        if isinstance(Xs, np.ndarray):
            if int(np.prod(shape)) >= (1 << 17):
                # large host results: page-locked memory (torch's caching host allocator) so the chunk-by-chunk
                # device->host copies run asynchronously at full PCIe rate and overlap the next chunk's kernels
                try:
                    import torch
                    return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
                except Exception:
                    pass
            return np.empty(shape, dtype=np.float64)
        import torch
        return torch.empty(shape, dtype=torch.float64, device=Xs.device)
