"""Build libhalob200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libhalob200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC']   # never --use_fast_math: parity is rtol 1e-10 in fp64


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: cannot build libhalob200.so')
    return exe


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources()+glob.glob(os.path.join(CSRC, '*.cuh'))+glob.glob(os.path.join(CSRC, '*.h')) + \
        [os.path.join(os.path.dirname(HERE), 'include', 'halob200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """Compile every .cu under csrc/ into one shared library; returns its path.  `defines` / `out` build an
    experimental variant beside the product library (kernel A/B runs on the GPU box: HALOB200_LIB selects it)."""
    lib = out or LIB
    if out is None and not force and not is_stale():
        return LIB
    objs = []
    procs = []
    builddir = os.path.join(HERE, 'build', os.path.basename(lib)[:-3] if out else '')
    os.makedirs(builddir, exist_ok=True)
    for src in sources():
        obj = os.path.join(builddir, os.path.basename(src)[:-3]+'.o')
        objs.append(obj)
        cmd = [_nvcc()]+NVCC_FLAGS+['-D'+d for d in defines]+['-c', src, '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
            print(' '.join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s' % (src, out))
        if verbose and out:
            print(out)
    link = [_nvcc(), '-shared', '-o', lib]+objs+['-lcudart']
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n'+r.stdout)
    return lib


if __name__ == '__main__':
    import sys
    defs = [a[2:] for a in sys.argv[1:] if a.startswith('-D')]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith('--out=')]
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, defines=defs, out=outs[0] if outs else None))
