"""Builds libglow_b200.so (sm_100a only) in-tree with nvcc.  No JIT cache: the .so travels with the repo."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
OBJ = os.path.join(PKG, 'build')
LIB_DIR = os.path.join(PKG, 'lib')
LIB = os.path.join(LIB_DIR, 'libglow_b200.so')

SOURCES = ['api.cu', 'aux.cu', 'gram_f64.cu', 'gram_f32.cu', 'gram_i8.cu', 'gram_p2.cu', 'split8.cu', 'mask8.cu', 'masked_sparse.cu', 'logreg.cu']
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC',
    '--expt-relaxed-constexpr'
]


def _nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: glow_b200 needs the CUDA toolkit to build its kernels')
    return nvcc


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=()):
    """Compiles every CUDA source for sm_100a and links the C-ABI shared library."""
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(os.path.dirname(PKG), 'include', 'glow_b200.h'))
    nvcc = _nvcc()
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace('.cu', '.o'))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append([nvcc] + NVCC_FLAGS + list(extra_flags) + ['-c', s, '-o', o])

    def run(cmd):
        if verbose:
            print(' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed: {" ".join(cmd)}\n{r.stdout}\n{r.stderr}')
        if verbose and (r.stdout or r.stderr):
            print(r.stdout, r.stderr, flush=True)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if jobs or force or _stale(LIB, objs):
        run([nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    return LIB


if __name__ == '__main__':
    flags = [a for a in sys.argv[1:] if a not in ('--force', '--verbose')]
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv, extra_flags=flags))
