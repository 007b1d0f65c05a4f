"""Empty h5py stub: the hot path never opens an HDF5 file (see oracle/refshim/README.md)."""


def File(*a, **k):
    raise NotImplementedError('refshim: h5py is not available in this container')
