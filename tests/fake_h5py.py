"""A stand-in for h5py (not installed in this image) that records what niwqg_b200/Saving.py writes:
``File(path, 'w')`` creates the file on disk (so the overwrite logic of Saving.file_exist is exercised) and keeps every
``create_dataset(name, data=..., dtype=...)`` in ``WRITTEN[path][name]``.  Test infrastructure only."""
import numpy as np

WRITTEN = {}


class File(object):
    def __init__(self, path, mode="r"):
        assert mode == "w", "Saving.py only ever writes"
        self.path = path
        WRITTEN[path] = {}
        open(path, "wb").close()

    def create_dataset(self, name, data=None, dtype=None):
        a = np.array(data, dtype=dtype) if dtype is not None else np.array(data)
        assert name not in WRITTEN[self.path], "dataset written twice: %s" % name
        WRITTEN[self.path][name] = a
        return a

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass
