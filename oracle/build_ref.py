# -*- coding: utf-8 -*-
"""
Recipe that materialises the UNMODIFIED reference package under
``oracle/_ref/`` -- TEST / BASELINE INFRASTRUCTURE ONLY.

    python oracle/build_ref.py            (also run by __graft_entry__.build())

The reference (netneurolab/pypyls, /root/reference) is a pure-Python package
whose ``setup.py`` cannot run under Python 3.12 (its vendored versioneer calls
``configparser.SafeConfigParser``), so ``pip install --target`` fails at
metadata generation (outcome recorded in DESIGN.md section 8).  What pip would
have put into the target directory for a pure-Python package is the package
tree itself; this recipe lays down exactly that: the ``pyls/`` package as it
lies under /root/reference (fixtures, docs and example data left out), plus the
one-function ``h5py`` stand-in that lets ``import pyls`` succeed in an image
without h5py (only ``pyls/io.py`` touches it).

``oracle/_ref/`` is listed in .gitignore (nothing of the reference enters the
history) but not in .gpurunignore, so the installed tree travels to the GPU box
where ``bench.py --impl reference`` and the ``cpu_baseline`` leg drive
``pyls.behavioral_pls(..., n_proc=<cores>)`` through the reference's own stock
code path (oracle/ref_runner.py).  Nothing under ``pypyls_b200/`` reads it.
"""

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference'
DST = os.path.join(HERE, '_ref')

H5PY_STUB = '''"""Stand-in so that `import pyls` works without h5py (pyls/io.py:6 is the only
user; save_results / load_results are not on the resampling path)."""


def is_hdf5(fname):
    return False
'''


def build(verbose=True):
    """(Re)creates oracle/_ref from /root/reference; returns its path or None
    when the reference is not present (GPU box: the prebuilt tree is used)."""
    if not os.path.isdir(os.path.join(SRC, 'pyls')):
        if verbose:
            print('oracle/build_ref: %s not present, keeping %s' % (
                SRC, DST if os.path.isdir(DST) else '(nothing)'))
        return DST if os.path.isdir(os.path.join(DST, 'pyls')) else None
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    # the whole package tree as it lies in the reference; only its test-suite
    # (3.4 MB of Matlab fixtures) is left out -- nothing imports it
    shutil.copytree(
        os.path.join(SRC, 'pyls'), os.path.join(DST, 'pyls'),
        ignore=shutil.ignore_patterns('tests', '__pycache__', '*.pyc'))
    shim = os.path.join(DST, '_shim')
    os.makedirs(shim)
    with open(os.path.join(shim, 'h5py.py'), 'w') as f:
        f.write(H5PY_STUB)
    with open(os.path.join(DST, 'PROVENANCE.txt'), 'w') as f:
        f.write('installed from %s by oracle/build_ref.py (unmodified pyls/ '
                'package tree; _shim/h5py.py is a stand-in module)\n' % SRC)
    if verbose:
        print('oracle/build_ref: installed reference package into', DST)
    return DST


def paths():
    """sys.path entries that make ``import pyls`` resolve to oracle/_ref."""
    out = [DST]
    try:
        import h5py  # noqa: F401
    except Exception:
        out.insert(0, os.path.join(DST, '_shim'))
    return out


def available():
    return os.path.isfile(os.path.join(DST, 'pyls', '__init__.py'))


if __name__ == '__main__':
    sys.exit(0 if build() else 1)
