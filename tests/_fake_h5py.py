# -*- coding: utf-8 -*-
"""
A small stand-in for the parts of the h5py API that save_results /
load_results use (File, Group, Dataset, attrs, is_hdf5), for images without
h5py: groups are nested dictionaries pickled to the file.  It checks the
LAYOUT logic of pypyls_b200/io.py (what becomes a group, a dataset, an
attribute; None -> 'None'); byte-level HDF5 compatibility needs the real h5py,
which the same test uses whenever it is installed.
"""

import pickle

import numpy as np

MAGIC = b'FAKEHDF5'


class Dataset:
    def __init__(self, data):
        self._data = np.array(data)

    def __getitem__(self, key):
        return self._data[key]

    def __setitem__(self, key, value):
        self._data[key] = value

    @property
    def shape(self):
        return self._data.shape


class Group:
    def __init__(self):
        self._nodes = {}
        self.attrs = {}

    def _walk(self, path, create=False):
        node = self
        for part in [p for p in path.split('/') if p]:
            if part not in node._nodes:
                if not create:
                    raise KeyError(path)
                node._nodes[part] = Group()
            node = node._nodes[part]
        return node

    def create_group(self, path):
        parts = [p for p in path.split('/') if p]
        parent = self._walk('/'.join(parts[:-1]), create=True)
        if parts[-1] in parent._nodes:
            raise ValueError('group exists: ' + path)
        parent._nodes[parts[-1]] = Group()
        return parent._nodes[parts[-1]]

    def create_dataset(self, name, shape=None, dtype=None, data=None):
        if data is None:
            data = np.zeros(shape, dtype)
        self._nodes[name] = Dataset(data)
        return self._nodes[name]

    def __getitem__(self, path):
        return self._walk(path)

    def items(self):
        return self._nodes.items()


class File(Group):
    def __init__(self, fname, mode='r'):
        super().__init__()
        self._fname, self._mode = str(fname), mode
        if mode == 'r':
            with open(self._fname, 'rb') as f:
                assert f.read(len(MAGIC)) == MAGIC
                root = pickle.load(f)
            self._nodes, self.attrs = root._nodes, root.attrs

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self._mode != 'r' and exc[0] is None:
            root = Group()
            root._nodes, root.attrs = self._nodes, self.attrs
            with open(self._fname, 'wb') as f:
                f.write(MAGIC)
                pickle.dump(root, f)
        return False


def is_hdf5(fname):
    try:
        with open(str(fname), 'rb') as f:
            return f.read(len(MAGIC)) == MAGIC
    except OSError:
        return False
