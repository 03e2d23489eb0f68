# -*- coding: utf-8 -*-
"""
Compiles the sm_100a CUDA sources of this package into ``libplsb200.so``
(in-tree, next to this file) with nvcc.  There is no CPU build: without the
library every entry point of the package raises.
"""

import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libplsb200.so')

NVCC_FLAGS = [
    '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo', '-shared', '-Xcompiler', '-fPIC',
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + [
        os.path.join(os.path.dirname(HERE), 'include', 'plsb200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Builds libplsb200.so if it is missing or older than its sources."""
    if not force and not is_stale():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + ['-o', LIB] + sources()
    if verbose:
        print(' '.join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    print(build(force=True, verbose=True))
