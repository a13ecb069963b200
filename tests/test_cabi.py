"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/gpo_b200.h declares,
and refuses to run without a GPU (no CPU fallback).  No compute calls are made here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'gpo_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gpo_[a-z_0-9]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
    from gpyopt_b200 import build
    build.build()
    from gpyopt_b200 import _lib
    return _lib.load_library()


def test_header_declares_the_path():
    syms = _declared_symbols()
    for s in ('gpo_create', 'gpo_destroy', 'gpo_fit', 'gpo_get_fmin', 'gpo_predict', 'gpo_predict_grad', 'gpo_acq',
              'gpo_lp_set', 'gpo_acq_topk', 'gpo_topk', 'gpo_strerror'):
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    from gpyopt_b200 import _lib
    for s in _declared_symbols():
        assert hasattr(lib, s), "libgpo_b200.so does not export %s" % s
        assert s in _lib.SIGNATURES, "ctypes binding lacks a prototype for %s" % s
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_option_keys_of_the_binding_match_the_header():
    """gpo_set_option keys: the ctypes binding's table (Engine.OPTIONS) names every GPO_OPT_* enumerator of the header with its
    value, and nothing else."""
    from gpyopt_b200 import _lib
    text = open(os.path.join(ROOT, 'include', 'gpo_b200.h')).read()
    enum = re.search(r'typedef enum \{([^}]*)\} gpo_option;', text, flags=re.S).group(1)
    declared = {m.group(1).lower(): int(m.group(2)) for m in re.finditer(r'GPO_OPT_([A-Z0-9_]+)\s*=\s*(\d+)', enum)}
    assert declared and declared == _lib.Engine.OPTIONS


def test_strerror_and_version(lib):
    assert lib.gpo_strerror(0) == b'ok'
    assert b'positive definite' in lib.gpo_strerror(-2)
    assert b'sm_100a' in lib.gpo_version()


def test_library_is_sm100a_only():
    """The shipped cubin must be sm_100a (no other architectures, no PTX-JIT fallback path)."""
    import subprocess
    from gpyopt_b200 import _lib
    out = subprocess.run(['cuobjdump', '-lelf', _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
    assert not re.search(r'sm_(?!100a)\d+', out)


def test_sass_uses_dmma_and_tma():
    import subprocess
    from gpyopt_b200 import _lib
    sass = subprocess.run(['cuobjdump', '-sass', _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'DMMA.8x8x4' in sass          # FP64 tensor-core path
    assert 'UTMALDG' in sass             # cp.async.bulk.tensor (TMA tiles of the GEMM operands)
    assert 'UBLKCP' in sass              # cp.async.bulk (TMA staging of K1's training tiles)
    assert 'SYNCS' in sass               # mbarrier pipeline


@pytest.mark.skipif(os.path.exists('/dev/nvidia0'), reason="GPU present")
def test_no_cpu_fallback(lib):
    """Without a B200 the engine must fail loudly rather than compute on the host."""
    from gpyopt_b200 import Engine, GPOError
    with pytest.raises(GPOError) as e:
        Engine(0)
    assert 'no usable' in str(e.value) or 'device' in str(e.value)
    from gpyopt_b200 import GPModel
    m = GPModel(max_iters=0)
    with pytest.raises(GPOError):
        m.updateModel(np.random.rand(5, 2), np.random.rand(5, 1), None, None)


def test_product_package_never_imports_the_oracle():
    """gpyopt_b200/ must not reference oracle/ (the oracle is test infrastructure)."""
    for top in ('gpyopt_b200', 'tools', 'include'):
      for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
                assert 'gpy_numpy' not in src and 'gpyopt_numpy' not in src and 'fake_engine' not in src, f
