"""Checkpoint containers: HDF5 through h5py when it is installed, and a small
built-in hierarchical container ('.npz' on disk) otherwise.

The reference checkpoints into HDF5 (nautilus/sampler.py:1253-1377); every
``write`` / ``read`` / ``update`` method of the reference touches an
``h5py.Group`` through a handful of calls only:

    group.attrs[key] = value          value = group.attrs[key]
    group.create_group(name)          group[name]        name in group
    group.create_dataset(name, data=..., maxshape=...)
    dataset.resize(shape); dataset[...] = array; numpy.array(dataset)

``open_store`` returns an object with exactly that surface: an ``h5py.File``
for '.h5' / '.hdf5' paths (the file is then laid out as the reference lays it
out, see Sampler.write), or a ``NpzFile`` for '.npz' paths -- the same tree
kept in memory and stored as ONE ``numpy.savez`` archive (no pickling:
datasets under '<path>', attributes under '<path>@<key>'), written to a
temporary file and renamed so that an interrupted write never destroys the
previous checkpoint.  h5py is not part of this image; the HDF5 branch is
what a maintainer with h5py gets, the '.npz' branch is what the tests run.
"""

import os
from pathlib import Path

import numpy as np

SUFFIXES_HDF5 = ('.h5', '.hdf5')
SUFFIX_NPZ = '.npz'


def check_suffix(filepath):
    """The reference accepts '.h5' / '.hdf5' only (sampler.py:1273-1274);
    '.npz' selects the built-in container."""
    if Path(filepath).suffix not in SUFFIXES_HDF5 + (SUFFIX_NPZ, ):
        raise ValueError("File ending must '.h5', '.hdf5' or '.npz'.")


class _Attrs(dict):
    """Attribute values come back the way h5py returns them: NumPy scalars,
    ``str`` for strings, arrays otherwise."""

    def __setitem__(self, key, value):
        value = np.asarray(value)
        if value.dtype == object:
            raise TypeError('cannot store an object attribute: ' + str(key))
        super().__setitem__(str(key), value)

    def __getitem__(self, key):
        value = super().__getitem__(key)
        if value.ndim == 0:
            return str(value[()]) if value.dtype.kind == 'U' else value[()]
        return value.copy()

    def raw(self, key):
        return super().__getitem__(key)


class Dataset:
    """An array with h5py's resize / assignment surface."""

    def __init__(self, data):
        self.data = np.array(data)

    @property
    def shape(self):
        return self.data.shape

    @property
    def dtype(self):
        return self.data.dtype

    def resize(self, shape):
        shape = tuple(np.atleast_1d(shape))
        new = np.zeros(shape, dtype=self.data.dtype)
        common = tuple(slice(0, min(a, b))
                       for a, b in zip(shape, self.data.shape))
        new[common] = self.data[common]
        self.data = new

    def __setitem__(self, key, value):
        self.data[key] = value

    def __getitem__(self, key):
        return self.data[key]

    def __array__(self, dtype=None, copy=None):
        return np.array(self.data, dtype=dtype)

    def __len__(self):
        return len(self.data)


class Group:
    def __init__(self):
        self.attrs = _Attrs()
        self._children = {}

    def create_group(self, name):
        if name in self._children:
            raise ValueError('name already exists: ' + name)
        group = Group()
        self._children[name] = group
        return group

    def create_dataset(self, name, data=None, maxshape=None, **kwargs):
        if name in self._children:
            raise ValueError('name already exists: ' + name)
        data = np.asarray(data)
        if data.dtype == object:
            raise TypeError('cannot store an object dataset: ' + name)
        dataset = Dataset(data)
        self._children[name] = dataset
        return dataset

    def __contains__(self, name):
        node = self
        for part in str(name).strip('/').split('/'):
            if not isinstance(node, Group) or part not in node._children:
                return False
            node = node._children[part]
        return True

    def __getitem__(self, name):
        node = self
        for part in str(name).strip('/').split('/'):
            node = node._children[part]
        return node

    def keys(self):
        return self._children.keys()

    # -- flattening ------------------------------------------------------
    def _flatten(self, prefix, out):
        if prefix:
            out[prefix + '/'] = np.zeros(0)      # keeps empty groups
        for key in self.attrs:
            out[prefix + '@' + key] = self.attrs.raw(key)
        for name, child in self._children.items():
            path = prefix + '/' + name if prefix else name
            if isinstance(child, Group):
                child._flatten(path, out)
            else:
                out[path] = child.data

    def _insert(self, parts, value, is_group):
        node = self
        for part in parts[:-1]:
            if part not in node._children:
                node._children[part] = Group()
            node = node._children[part]
        if is_group:
            if parts[-1] not in node._children:
                node._children[parts[-1]] = Group()
        else:
            node._children[parts[-1]] = Dataset(value)


class NpzFile(Group):
    """The root group.  Modes as in h5py: 'r' read, 'r+' read / write back on
    close, 'x' create (fails if the file exists), 'w' create / truncate."""

    def __init__(self, filepath, mode='r'):
        super().__init__()
        self.filepath = Path(filepath)
        self.mode = mode
        if mode not in ('r', 'r+', 'x', 'w'):
            raise ValueError('unknown mode ' + repr(mode))
        if mode == 'x' and self.filepath.exists():
            raise FileExistsError(str(self.filepath))
        if mode in ('r', 'r+'):
            with np.load(self.filepath, allow_pickle=False) as archive:
                for key in archive.files:
                    self._load(key, archive[key])

    def _load(self, key, value):
        if '@' in key:
            path, attr = key.split('@', 1)
            node = self
            for part in [p for p in path.split('/') if p]:
                if part not in node._children:
                    node._children[part] = Group()
                node = node._children[part]
            dict.__setitem__(node.attrs, attr, value)
            return
        parts = [p for p in key.split('/') if p]
        if not parts:
            return
        self._insert(parts, value, is_group=key.endswith('/'))

    def flush(self):
        if self.mode == 'r':
            return
        out = {}
        self._flatten('', out)
        tmp = self.filepath.with_name(self.filepath.name + '.tmp')
        with open(tmp, 'wb') as stream:
            np.savez(stream, **out)
        os.replace(tmp, self.filepath)

    def close(self):
        self.flush()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if exc[0] is None:
            self.close()
        return False


def _h5py(filepath):
    try:
        import h5py
    except ImportError as error:
        raise ImportError(
            "HDF5 checkpoints ('.h5' / '.hdf5') need the h5py package, which "
            "is not installed; use a '.npz' path for the built-in container "
            "(same layout, one numpy archive).") from error
    return h5py


def require_backend(filepath):
    """Fail when the Sampler is made, not at the first write hours later."""
    check_suffix(filepath)
    if Path(filepath).suffix != SUFFIX_NPZ:
        _h5py(filepath)


def open_store(filepath, mode='r'):
    """``h5py.File(filepath, mode)`` for HDF5 paths, ``NpzFile`` for '.npz'.
    """
    filepath = Path(filepath)
    check_suffix(filepath)
    if filepath.suffix == SUFFIX_NPZ:
        return NpzFile(filepath, mode)
    return _h5py(filepath).File(filepath, mode)
