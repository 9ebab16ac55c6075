"""Hierarchical store behind ``GRB.save`` / ``GRBSave.from_file``.

The reference writes HDF5 (grb/grb.py:205-314).  h5py may be absent (it is in the build image), so
the same tree -- identical group / dataset / attribute paths -- has two backends:

  * ``.h5`` / ``.hdf5`` with h5py importable: a real HDF5 file, readable by the reference's
    ``GRBSave.from_file`` (lzf compression for the event arrays like the reference);
  * otherwise: a NumPy ``.npz`` twin.  Datasets are stored under their path, attributes under
    ``<path>@<name>``, (empty) groups under ``<path>/``.

``open_tree`` picks the backend from the file's magic bytes, so a file written with one backend under
any name is read back correctly.
"""
import os

import numpy as np


def have_h5py():
    try:
        import h5py  # noqa: F401

        return True
    except Exception:
        return False


def _norm(path):
    return "/" + path.strip("/") if path.strip("/") else "/"


class Tree(object):
    """In-memory tree: datasets, attributes and groups by absolute path."""

    def __init__(self):
        self._datasets = {}
        self._attrs = {"/": {}}
        self._compress = set()

    # -- building -----------------------------------------------------------------------------------
    def group(self, path):
        p = _norm(path)
        parts = p.strip("/").split("/") if p != "/" else []
        for i in range(1, len(parts) + 1):
            self._attrs.setdefault("/" + "/".join(parts[:i]), {})
        return p

    def attrs(self, path):
        return self._attrs[self.group(path)]

    def dataset(self, path, value, compress=False):
        p = _norm(path)
        self.group(os.path.dirname(p))
        if isinstance(value, dict):
            for k, v in value.items():
                self.dataset(f"{p}/{k}", v, compress)
            return
        if not isinstance(value, (np.ndarray, np.generic, str, bytes, float, int, list, tuple)):
            raise ValueError("Cannot save %s type" % type(value))          # hdf5_utils.py:28
        self._datasets[p] = np.asarray(value) if not isinstance(value, (str, bytes)) else value
        if compress:
            self._compress.add(p)

    # -- reading ------------------------------------------------------------------------------------
    def __contains__(self, path):
        p = _norm(path)
        return p in self._datasets or p in self._attrs

    def get(self, path):
        v = self._datasets[_norm(path)]
        if isinstance(v, np.ndarray) and v.shape == ():
            v = v[()]
            if isinstance(v, (bytes, np.bytes_)):
                v = v.decode()
            elif isinstance(v, np.str_):
                v = str(v)
        return v

    def children(self, path):
        p = _norm(path)
        prefix = p if p.endswith("/") else p + "/"
        names = []
        for q in list(self._datasets) + list(self._attrs):
            if q.startswith(prefix) and q != p:
                n = q[len(prefix):].split("/")[0]
                if n and n not in names:
                    names.append(n)
        return sorted(names)                    # h5py lists a group's members by name, whatever the backend

    def is_group(self, path):
        return _norm(path) in self._attrs

    def load_dict(self, path):
        """recursively_load_dict_contents_from_group (utils/hdf5_utils.py:33-53)"""
        if not self.is_group(path):
            raise KeyError(path)
        out = {}
        for n in self.children(path):
            q = f"{_norm(path).rstrip('/')}/{n}"
            out[n] = self.load_dict(q) if self.is_group(q) and _norm(q) not in self._datasets else self.get(q)
        return out

    # -- backends -----------------------------------------------------------------------------------
    def write(self, file_name, backend=None):
        file_name = str(file_name)
        if backend is None:
            backend = "h5" if have_h5py() else "npz"
        if backend == "h5":
            import h5py

            with h5py.File(file_name, "w") as f:
                for p in sorted(self._attrs):
                    g = f if p == "/" else f.require_group(p)
                    for k, v in self._attrs[p].items():
                        g.attrs[k] = v
                for p, v in self._datasets.items():
                    if isinstance(v, np.ndarray) and v.dtype.kind == "U":
                        # h5py has no conversion path for numpy '<U' arrays: variable-length UTF-8 strings
                        v = v.astype(h5py.string_dtype())
                    if p in self._compress and isinstance(v, np.ndarray) and v.ndim > 0:
                        f.create_dataset(p, data=v, compression="lzf")
                    else:
                        f[p] = v
            return
        flat = {}
        for p, a in self._attrs.items():
            flat[p.rstrip("/") + "/"] = np.zeros(0)
            for k, v in a.items():
                flat[f"{p.rstrip('/')}@{k}"] = np.asarray(v)
        for p, v in self._datasets.items():
            flat[p] = np.asarray(v)
        with open(file_name, "wb") as fh:                   # keep the caller's file name (no .npz suffix added)
            np.savez_compressed(fh, **flat)


def open_tree(file_name):
    """Read a file written by ``Tree.write`` (either backend) back into a ``Tree``."""
    file_name = str(file_name)
    with open(file_name, "rb") as fh:
        magic = fh.read(8)
    t = Tree()
    if magic.startswith(b"\x89HDF"):
        try:
            import h5py
        except ImportError:                      # a real HDF5 file and no h5py: the engine's own read-only decoder
            from ..utils import hdf5_lite as h5py

        with h5py.File(file_name, "r") as f:
            t._attrs["/"] = dict(f.attrs)

            def visit(name, obj):
                p = "/" + name
                if isinstance(obj, h5py.Group):
                    t._attrs[p] = dict(obj.attrs)
                else:
                    t.group(os.path.dirname(p))
                    v = obj[()]
                    if isinstance(v, np.ndarray) and v.dtype.kind in "OS":      # string datasets come back as bytes
                        v = np.array([x.decode() if isinstance(x, bytes) else str(x) for x in v.ravel()]).reshape(v.shape)
                    t._datasets[p] = v.decode() if isinstance(v, bytes) else v

            f.visititems(visit)
        return t
    with np.load(file_name, allow_pickle=False) as z:
        for key in z.files:
            if key.endswith("/"):
                t.group(key)
            elif "@" in key.rsplit("/", 1)[-1]:
                p, k = key.rsplit("@", 1)
                v = z[key]
                t.attrs(p or "/")[k] = v[()] if v.shape == () else v
                if isinstance(t.attrs(p or "/")[k], np.str_):
                    t.attrs(p or "/")[k] = str(t.attrs(p or "/")[k])
            else:
                t.group(os.path.dirname(key))
                t._datasets[_norm(key)] = z[key]
    return t
