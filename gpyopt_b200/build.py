"""Build the sm_100a shared library in-tree with nvcc (no torch dependency in the .so).

    python -m gpyopt_b200.build        # -> gpyopt_b200/lib/libgpo_b200.so
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'gpo_api.cu')
DEPS = [os.path.join(HERE, 'csrc', f) for f in ('gpo_api.cu', 'dgemm_tma.cuh', 'kernels.cuh', 'refit_kernels.cuh', 'kernfun.cuh', 'common.cuh')]
DEPS.append(os.path.join(HERE, '..', 'include', 'gpo_b200.h'))
OUT_DIR = os.path.join(HERE, 'lib')
OUT = os.path.join(OUT_DIR, 'libgpo_b200.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-shared',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v', '--use_fast_math=false']


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(f) > t for f in DEPS)


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> lib/libgpo_b200.so for sm_100a.  Returns the path of the library."""
    if not force and not needs_build():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = os.environ.get('NVCC', 'nvcc')
    flags = [f for f in NVCC_FLAGS if not f.startswith('--use_fast_math')] + os.environ.get('GPO_NVCC_EXTRA', '').split()
    cmd = [nvcc] + flags + ['-o', OUT, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(OUT_DIR, 'build.log'), 'w') as fh:
        fh.write(' '.join(cmd) + '\n' + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError('nvcc failed (%d)' % res.returncode)
    if verbose:
        print(log)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
