"""Import shim (TEST INFRASTRUCTURE ONLY): the reference's io_video.py imports h5py at module
level; h5py is not installed here and only HDF5Reader uses it.  Nothing is implemented."""


class File:   # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("h5py is not available in this environment")
