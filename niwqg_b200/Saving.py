"""Disk output with the reference's HDF5 layout (niwqg/Saving.py):

    <path>/setup.h5          grid/{nx,x,y,wv,k,l}
    <path>/snapshots/{t:015.0f}.h5   datasets named by ``fields`` (t, q, phi | c)
    <path>/diagnostics.h5    one dataset per diagnostic

h5py is imported lazily (it is optional in this image); fields are read from the
device only at the snapshot cadence.
"""
import os


def _h5py():
    try:
        import h5py
    except Exception as e:            # pragma: no cover - depends on the image
        raise ImportError("save_to_disk=True needs h5py (%s)" % e)
    return h5py


def initialize_save_snapshots(self, path):
    """niwqg/Saving.py:6-21."""
    self.fno = path
    if (not os.path.isdir(self.fno)) and self.save_to_disk:
        os.makedirs(self.fno)
        os.makedirs(self.fno + "/snapshots/")


def file_exist(fno, overwrite=True):
    """niwqg/Saving.py:23-36."""
    if os.path.exists(fno):
        if overwrite:
            os.remove(fno)
        else:
            raise IOError("File exists: {0}".format(fno))


def save_setup(self):
    """niwqg/Saving.py:38-57."""
    if self.save_to_disk:
        fno = self.fno + '/setup.h5'
        file_exist(fno, overwrite=self.overwrite)
        with _h5py().File(fno, 'w') as h5file:
            h5file.create_dataset("grid/nx", data=(self.nx), dtype=int)
            h5file.create_dataset("grid/x", data=(self.x))
            h5file.create_dataset("grid/y", data=(self.y))
            h5file.create_dataset("grid/wv", data=self.wv)
            h5file.create_dataset("grid/k", data=self.kk)
            h5file.create_dataset("grid/l", data=self.ll)


def save_snapshots(self, fields=['t', 'q', 'p']):
    """niwqg/Saving.py:59-86."""
    if (not (self.tc % self.tsnaps)) and self.save_to_disk:
        fno = self.fno + '/snapshots/{:015.0f}'.format(self.t) + '.h5'
        file_exist(fno)
        with _h5py().File(fno, 'w') as h5file:
            for field in fields:
                if field == 't':
                    h5file.create_dataset(field, data=(self.t))
                    continue
                val = getattr(self, field, None)       # ONE device-to-host copy per field (hasattr would be a second one)
                if val is not None:
                    h5file.create_dataset(field, data=val)


def save_diagnostics(self):
    """niwqg/Saving.py:88-101."""
    fno = self.fno + '/diagnostics.h5'
    file_exist(fno, overwrite=self.overwrite)
    with _h5py().File(fno, 'w') as h5file:
        for key in self.diagnostics.keys():
            h5file.create_dataset(key, data=(self.diagnostics[key]['value']))
