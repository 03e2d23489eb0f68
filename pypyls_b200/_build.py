# -*- coding: utf-8 -*-
"""
Compiles the sm_100a CUDA sources of this package into ``libplsb200.so``
(in-tree, next to this file) with nvcc.  There is no CPU build: without the
library every entry point of the package raises.

Every ``csrc/*.cu`` is compiled to an object under ``csrc/build/`` (in
parallel, only when it is older than its source or a shared header) and the
objects are linked into the shared library.
"""

import glob
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, 'build')
LIB = os.path.join(HERE, 'libplsb200.so')

NVCC_FLAGS = [
    '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo', '-Xcompiler', '-fPIC',
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def headers():
    return glob.glob(os.path.join(CSRC, '*.cuh')) + [
        os.path.join(os.path.dirname(HERE), 'include', 'plsb200.h')]


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in sources() + headers())


def build(force=False, verbose=False):
    """Builds libplsb200.so if it is missing or older than its sources."""
    if not force and not is_stale():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    extra = os.environ.get('PLSB_NVCC_EXTRA', '').split()
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in headers())
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        if force or not os.path.exists(obj) or \
                os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append([nvcc] + NVCC_FLAGS + extra + ['-c', src, '-o', obj])

    def run(cmd):
        if verbose:
            print(' '.join(cmd), flush=True)
        subprocess.run(cmd, check=True)

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(run, jobs))
    run([nvcc, '-shared', '-o', LIB] + objs)
    return LIB


if __name__ == '__main__':
    print(build(force=True, verbose=True))
