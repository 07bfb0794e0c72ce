"""A minimal in-memory stand-in for the part of h5py that the DESI-DR2 reader / writer uses (h5py is not installed in
this image, so the HDF5 branches of cupix_b200/px_data/data_DESI_DR2.py -- and the reference's own reader,
cupix/px_data/data_DESI_DR2.py:66-99 -- would otherwise never execute).  Files live in a module-level dict keyed by path;
groups, datasets and attributes follow h5py's access patterns: ``f['a/b/c/']``, ``g.attrs[k]``, ``create_group``,
``create_dataset(path, data=...)`` with intermediate groups, ``dataset[:]``, context manager, modes 'r' / 'w' / 'a'."""
import numpy as np

_FILES = {}


class _Dataset(object):
    def __init__(self, data):
        self._data = np.array(data)
        self.attrs = {}

    def __getitem__(self, key):
        return self._data[key] if key is not Ellipsis and key != slice(None) else self._data.copy()

    @property
    def shape(self):
        return self._data.shape


class _Group(object):
    def __init__(self):
        self.attrs = {}
        self._children = {}

    def _walk(self, path, create=False):
        node = self
        for part in [p for p in path.split("/") if p]:
            if part not in node._children:
                if not create:
                    raise KeyError(path)
                node._children[part] = _Group()
            node = node._children[part]
        return node

    def __getitem__(self, path):
        return self._walk(path)

    def __contains__(self, path):
        try:
            self._walk(path)
            return True
        except KeyError:
            return False

    def create_group(self, path):
        return self._walk(path, create=True)

    def create_dataset(self, path, data=None):
        parts = [p for p in path.split("/") if p]
        parent = self._walk("/".join(parts[:-1]), create=True)
        parent._children[parts[-1]] = _Dataset(data)
        return parent._children[parts[-1]]

    def keys(self):
        return self._children.keys()


class File(_Group):
    def __init__(self, path, mode="r"):
        path = str(path)
        if mode == "w":
            _FILES[path] = _Group()
        elif path not in _FILES:
            raise OSError("no such (fake) HDF5 file: %s" % path)
        root = _FILES[path]
        self.attrs, self._children = root.attrs, root._children

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
