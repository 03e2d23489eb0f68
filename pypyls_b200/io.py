# -*- coding: utf-8 -*-
"""
Results interchange with the reference: the HDF5 layout of ``pyls.save_results``
/ ``pyls.load_results`` (pyls/io.py:12-122) -- every (nested) results
dictionary is a group under ``/results``, arrays are datasets, scalars and
strings are attributes of their group and ``None`` is stored as the string
``'None'``.  Files written here load in the reference and vice versa.

Host glue, not part of the accelerated path.  ``h5py`` is imported when a
function is called; it is not a dependency of the engine.
"""

import numpy as np

from .structures import PLSResults

ROOT = '/results'


def _h5py():
    try:
        import h5py
    except ImportError as err:          # pragma: no cover
        raise ImportError('save_results / load_results need the `h5py` '
                          'package') from err
    return h5py


def _filename(fname):
    fname = str(fname)
    return fname if fname.endswith('.hdf5') else fname + '.hdf5'


def save_results(fname, results):
    """
    Saves PLS `results` to the HDF5 file `fname` ('.hdf5' is appended if
    missing) and returns the file name.
    """
    h5py = _h5py()
    fname = _filename(fname)
    with h5py.File(fname, 'w') as out:
        todo = [(ROOT, results)]
        while todo:
            path, mapping = todo.pop()
            group = out.create_group(path)
            for key, value in mapping.items():
                if isinstance(value, dict):
                    todo.append((path + '/' + key, value))
                elif isinstance(value, np.ndarray):
                    group.create_dataset(key, data=value)
                else:
                    group.attrs[key] = 'None' if value is None else value
    return fname


def load_results(fname):
    """
    Loads PLS results stored by :func:`save_results` (or by the reference's
    ``pyls.save_results``) -> :obj:`pypyls_b200.structures.PLSResults`.
    """
    h5py = _h5py()
    fname = _filename(fname)
    if not h5py.is_hdf5(fname):
        raise TypeError('Provided file {} is not valid HDF5 format.'
                        .format(fname))

    def read(group):
        found = {}
        for key, node in group.items():
            found[key] = read(node) if isinstance(node, h5py.Group) \
                else node[()]
        for key, value in group.attrs.items():
            found[key] = None if isinstance(value, str) and value == 'None' \
                else value
        return found

    with h5py.File(fname, 'r') as src:
        return PLSResults(**read(src[ROOT]))
