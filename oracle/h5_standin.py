"""An in-memory stand-in for the slice of the h5py API that cosmogrb's save / load code touches.

TEST INFRASTRUCTURE ONLY (never imported by the product package).  h5py is not installable in the authoring
image, so the reference's own ``GRB.save`` (grb/grb.py:205-314), ``GRBSave.from_file`` (io/grb_save.py:146-255)
and ``recursively_{save,load}_dict_contents`` (utils/hdf5_utils.py) cannot touch a real HDF5 file here.  With this
module registered as ``sys.modules['h5py']`` that code runs UNMODIFIED against a tree of ``Group`` / ``Dataset``
objects, which is what pins SURVEY.md §8 row f1: the layout the reference writes (group / dataset / attribute paths,
dtypes, shapes) is recorded from its own writer, and files written by ``cosmogrb_b200`` are read by its own reader.

Value semantics follow h5py 3: numbers come back as numpy scalars / arrays, ``str`` attributes as ``str``, ``str``
datasets as ``bytes`` from ``ds[()]``, variable-length string arrays as object arrays of ``bytes``.  A "file" is a
pickle behind the 8-byte HDF5 signature, so ``cosmogrb_b200.io.store.open_tree`` takes its HDF5 branch for it.
"""
from __future__ import annotations

import pickle
import types

import numpy as np

MAGIC = b"\x89HDF\r\n\x1a\n"
__version__ = "standin-3"


def string_dtype(encoding="utf-8", length=None):
    """h5py's variable-length string dtype is an object dtype with metadata; ``astype`` of it yields ``str`` objects,
    which ``Dataset`` stores as ``bytes`` (see ``_to_stored``)."""
    return np.dtype("O")


def _to_stored(value):
    if isinstance(value, str):
        return value.encode()
    if isinstance(value, bytes):
        return value
    a = np.asarray(value)
    if a.dtype.kind == "U":
        raise TypeError("No conversion path for dtype: %r" % a.dtype)          # what h5py raises
    if a.dtype.kind == "O":
        return np.array([x.encode() if isinstance(x, str) else x for x in a.ravel()], dtype=object).reshape(a.shape)
    return a.copy()


class AttributeManager(dict):
    def __setitem__(self, key, value):
        if isinstance(value, (str, bytes)):
            dict.__setitem__(self, key, value if isinstance(value, str) else value.decode())
            return
        a = np.asarray(value)
        if a.dtype.kind == "U":
            a = a.astype(object)
        dict.__setitem__(self, key, a[()] if a.shape == () else a.copy())


class Dataset(object):
    def __init__(self, name, value, compression=None):
        self.name = name
        self._v = _to_stored(value)
        self.compression = compression
        self.attrs = AttributeManager()

    @property
    def shape(self):
        return () if isinstance(self._v, bytes) else self._v.shape

    @property
    def dtype(self):
        return np.dtype("O") if isinstance(self._v, bytes) else self._v.dtype

    def __getitem__(self, idx):
        if isinstance(self._v, bytes):
            if idx != ():
                raise ValueError("scalar dataset")
            return self._v
        out = self._v[idx]
        return out.copy() if isinstance(out, np.ndarray) else out

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._v, dtype=dtype)

    def __len__(self):
        return len(self._v)


class Group(object):
    def __init__(self, name="/"):
        self.name = name
        self._children = {}
        self.attrs = AttributeManager()

    # -- path handling: "a//b/" == "a/b" (hdf5_utils.py builds such paths) --------------------------------------------
    @staticmethod
    def _parts(path):
        return [p for p in str(path).split("/") if p]

    def _child_name(self, part):
        return (self.name.rstrip("/") + "/" + part)

    def _walk(self, parts, create):
        g = self
        for p in parts:
            if p not in g._children:
                if not create:
                    raise KeyError("Unable to open object (component not found: %s)" % p)
                g._children[p] = Group(g._child_name(p))
            g = g._children[p]
            if not isinstance(g, Group) and p is not parts[-1]:
                raise KeyError("%s is not a group" % p)
        return g

    def create_group(self, name):
        parts = self._parts(name)
        parent = self._walk(parts[:-1], True)
        if parts[-1] in parent._children:
            raise ValueError("Unable to create group (name already exists)")
        g = parent._children[parts[-1]] = Group(parent._child_name(parts[-1]))
        return g

    def require_group(self, name):
        g = self._walk(self._parts(name), True)
        if not isinstance(g, Group):
            raise TypeError("Incompatible object (Dataset) already exists")
        return g

    def create_dataset(self, name, data=None, compression=None, **kw):
        parts = self._parts(name)
        parent = self._walk(parts[:-1], True)
        if parts[-1] in parent._children:
            raise ValueError("Unable to create dataset (name already exists)")
        if compression is not None and np.ndim(data) == 0:
            raise TypeError("Scalar datasets don't support chunk/filter options")     # h5py's rule
        d = parent._children[parts[-1]] = Dataset(parent._child_name(parts[-1]), data, compression)
        return d

    def __setitem__(self, name, value):
        self.create_dataset(name, data=value)

    def __getitem__(self, name):
        parts = self._parts(name)
        return self._walk(parts, False) if parts else self

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    # h5py iterates a group in name order (no track_order), not in creation order
    def keys(self):
        return sorted(self._children)

    def values(self):
        return [self._children[k] for k in sorted(self._children)]

    def items(self):
        return [(k, self._children[k]) for k in sorted(self._children)]

    def __iter__(self):
        return iter(sorted(self._children))

    def __len__(self):
        return len(self._children)

    def visititems(self, func):
        def rec(g, prefix):
            for k, v in g.items():
                p = prefix + k
                r = func(p, v)
                if r is not None:
                    return r
                if isinstance(v, Group):
                    r = rec(v, p + "/")
                    if r is not None:
                        return r
            return None

        return rec(self, "")


class File(Group):
    def __init__(self, name, mode="r", **kw):
        Group.__init__(self, "/")
        self.filename = str(name)
        self.mode = mode
        if mode in ("r", "r+", "a"):
            try:
                with open(self.filename, "rb") as fh:
                    if fh.read(len(MAGIC)) != MAGIC:
                        raise OSError("Unable to open file (file signature not found)")
                    root = pickle.load(fh)
            except FileNotFoundError:
                if mode != "a":
                    raise
                root = None
            if root is not None:
                self._children = root._children
                self.attrs = root.attrs
        elif mode not in ("w", "x", "w-"):
            raise ValueError(mode)

    def close(self):
        if self.mode != "r":
            root = Group("/")
            root._children = self._children
            root.attrs = self.attrs
            with open(self.filename, "wb") as fh:
                fh.write(MAGIC)
                pickle.dump(root, fh, protocol=4)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


# the reference tests ``isinstance(item, h5py._hl.dataset.Dataset)`` (utils/hdf5_utils.py:47-49)
_hl = types.SimpleNamespace(dataset=types.SimpleNamespace(Dataset=Dataset), group=types.SimpleNamespace(Group=Group),
                            files=types.SimpleNamespace(File=File))


def listing(group, with_values=False):
    """Flat description of a tree: {path: {"kind", "dtype", "shape", "compression", "attrs": {name: dtype}}}.
    String payloads read "str"; this is the fixture format of tests/golden/grb_save_layout.json."""
    def kind_of(v):
        if isinstance(v, (str, bytes)):
            return "str", []
        a = np.asarray(v)
        if a.dtype.kind in "OSU":
            return "str", list(a.shape)
        return a.dtype.str.lstrip("<>|=") if a.dtype.kind in "iufb" else a.dtype.str, list(a.shape)

    out = {}

    def add(path, obj):
        e = {"kind": "group" if isinstance(obj, Group) else "dataset",
             "attrs": {k: kind_of(v)[0] for k, v in sorted(obj.attrs.items())}}
        if isinstance(obj, Dataset):
            e["dtype"], e["shape"] = kind_of(obj._v)
            e["compression"] = obj.compression
        out[path] = e

    add("/", group)
    group.visititems(lambda p, o: add("/" + p, o))
    return out
