"""Stub so the read-only reference imports here (h5py is absent; only VB.save/load use it)."""
class _V:  # reference checks h5py.version.hdf5_version_tuple at import
    hdf5_version_tuple = (1, 14, 0)
    version = "0.0-stub"
version = _V()
def File(*a, **k):
    raise RuntimeError("h5py stub: HDF5 IO is not available in this container")
