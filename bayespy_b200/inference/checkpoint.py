"""Checkpoints of a VB run: ``VB.save`` / ``VB.load`` / autosave (vmp.py:237-356, 750-758; node side
stochastic.py:305-354, expfamily.py:507-535).

On-disk hierarchy — the reference's:

    nodes/<name>/u0, u1, ...      moments                  nodes/<name>/phi0, ...   natural parameters
    nodes/<name>/observed, f, g                            boundterms/<name>        per-node lower-bound history
    L, cputime, iter, converged, callback_output           user_data/<key>

Container: HDF5 through ``h5py`` when a real h5py is importable (the files are then interchangeable with the
reference's); otherwise a NumPy ``.npz`` archive whose member names are exactly the HDF5 paths above (this image has
no h5py and no network, so this is the container the tests exercise).  ``load`` recognises either.

Saving materialises device-resident state on the host: virtual quantities of the fused sweeps (factored second
moments, lazily formed natural parameters) are formed at that point — a checkpoint of a plated node costs one pass.
"""
import zipfile

import numpy as np

from ..darray import DArray
from ..engine.gaussian import dense


def _real_h5py():
    try:
        import h5py
        if hasattr(h5py, "Dataset") and hasattr(h5py, "Group"):
            return h5py
    except Exception:
        pass
    return None


def _host(x):
    """Device / virtual array -> host ndarray."""
    if x is None:
        return np.array(np.nan)
    x = dense(x) if hasattr(x, "materialize") else x
    if isinstance(x, DArray):
        return x.numpy()
    return np.asarray(x)


class _NpzWriter:
    def __init__(self, filename):
        self.filename, self.items = filename, {}

    def put(self, path, value):
        self.items[path] = np.asarray(value)

    def close(self):
        with open(self.filename, "wb") as f:              # exact name (np.savez would append .npz)
            np.savez_compressed(f, **self.items)


class _H5Writer:
    def __init__(self, filename, h5py):
        self.f = h5py.File(filename, "w")

    def put(self, path, value):
        value = np.asarray(value)
        try:
            self.f.create_dataset(path, data=value, compression="gzip")      # misc.py:456-469
        except TypeError:
            self.f.create_dataset(path, data=value)

    def close(self):
        self.f.close()


class _Reader:
    def __init__(self, filename):
        self.npz = self.h5 = None
        if zipfile.is_zipfile(filename):
            self.npz = np.load(filename, allow_pickle=False)
        else:
            h5py = _real_h5py()
            if h5py is None:
                raise RuntimeError("%s is not a NumPy checkpoint and h5py is not available to read HDF5" % filename)
            self.h5 = h5py.File(filename, "r")

    def has(self, path):
        return (path in self.npz.files) if self.npz is not None else (path in self.h5)

    def get(self, path):
        if not self.has(path):
            raise KeyError(path)
        return np.asarray(self.npz[path] if self.npz is not None else self.h5[path][...])

    def close(self):
        (self.npz if self.npz is not None else self.h5).close()


def open_writer(filename):
    h5py = _real_h5py()
    return _H5Writer(filename, h5py) if h5py is not None else _NpzWriter(filename)


def save_node(w, node, prefix):
    """stochastic.py:320-330 + expfamily.py:507-521."""
    for i, u in enumerate(node.u):
        w.put("%s/u%d" % (prefix, i), _host(u))
    w.put("%s/observed" % prefix, np.asarray(node.observed))
    if hasattr(node, "phi") and node.phi is not None:
        for i, p in enumerate(node.phi):
            w.put("%s/phi%d" % (prefix, i), _host(p))
        w.put("%s/f" % prefix, _host(getattr(node, "f", np.nan)))
        w.put("%s/g" % prefix, _host(getattr(node, "g", np.nan)))


def load_node(r, node, prefix):
    """stochastic.py:341-354 + expfamily.py:524-535; shapes are broadcast against the node's own."""
    if not r.has("%s/u0" % prefix):
        raise Exception("File does not contain variable %s" % node.name)
    node.u = [DArray.from_numpy(np.array(r.get("%s/u%d" % (prefix, i)), dtype=np.float64)) for i in range(len(node.u))]
    if hasattr(node, "phi") and node.phi is not None and r.has("%s/phi0" % prefix):
        node.phi = [DArray.from_numpy(np.array(r.get("%s/phi%d" % (prefix, i)), dtype=np.float64))
                    for i in range(len(node.phi))]
        node.f = float(np.asarray(r.get("%s/f" % prefix)).reshape(-1)[0]) if np.asarray(r.get("%s/f" % prefix)).size == 1 \
            else DArray.from_numpy(np.array(r.get("%s/f" % prefix), dtype=np.float64))
        node.g = DArray.from_numpy(np.array(r.get("%s/g" % prefix), dtype=np.float64))
    old = node.observed
    obs = np.asarray(r.get("%s/observed" % prefix))
    node.observed = bool(obs) if obs.ndim == 0 else obs.astype(bool)
    node._version = getattr(node, "_version", 0) + 1
    if hasattr(node, "_fused"):
        node._fused = None
    if np.any(np.asarray(old) != np.asarray(node.observed)):
        node._update_mask()
