"""Stub so that `import pyls` works without h5py (only pyls/io.py uses it).
Used by make_golden.py in the build container only."""


def is_hdf5(fname):
    return False
