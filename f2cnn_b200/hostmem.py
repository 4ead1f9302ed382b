"""Host memory shared by the processes of one box: the ONE (N, 2R+1, C) float32 input tensor that
every rank of a multi-GPU run places its rows into (BASELINE config 3; the reference builds one
array in one process, scripts/processing/InputGenerator.py:63,67-82).

torchrun starts one process per GPU, so "one host buffer" means a shared mapping: a file on a
RAM-backed filesystem (/dev/shm) mapped by every rank.  Rows are written by each rank's placement
threads (engine.Placer) at offsets known before launch -- a prefix sum over the label grid -- so
no rank ever touches another rank's rows and there is no collective."""
import mmap
import os

import numpy as np


def shm_dir():
    """A RAM-backed directory with room for corpus-sized tensors, or None."""
    for d in (os.environ.get("F2CNN_B200_SHM"), "/dev/shm"):
        if d and os.path.isdir(d) and os.access(d, os.W_OK):
            return d
    return None


def shm_free_bytes(d=None):
    d = d or shm_dir()
    if d is None:
        return 0
    st = os.statvfs(d)
    return int(st.f_bavail) * int(st.f_frsize)


class SharedArray:
    """numpy view of a file-backed shared mapping.  The creating process sizes the file; the others
    open it after a barrier.  unlink() removes the name (the memory lives on while mapped)."""

    def __init__(self, path, shape, dtype=np.float32, create=False):
        self.path = path
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        nbytes = max(int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize, 1)
        flags = os.O_RDWR | (os.O_CREAT | os.O_TRUNC if create else 0)
        fd = os.open(path, flags, 0o600)
        try:
            if create:
                os.ftruncate(fd, nbytes)
            elif os.fstat(fd).st_size < nbytes:
                raise ValueError("%s holds %d bytes, expected %d" % (path, os.fstat(fd).st_size, nbytes))
            self._map = mmap.mmap(fd, nbytes, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
        finally:
            os.close(fd)
        try:  # tmpfs supports huge pages when the mount allows them; harmless otherwise
            self._map.madvise(mmap.MADV_HUGEPAGE)
        except (AttributeError, OSError, ValueError):
            pass
        self.array = np.frombuffer(self._map, dtype=self.dtype,
                                   count=int(np.prod(self.shape, dtype=np.int64))).reshape(self.shape)

    def unlink(self):
        try:
            os.unlink(self.path)
        except FileNotFoundError:
            pass
