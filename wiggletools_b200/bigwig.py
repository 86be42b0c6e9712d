"""BigWig files -> SoA tracks through the library's host batch decoder
(wb200_bw_*, wiggletools_b200/csrc/bigwig.cpp)."""
import ctypes as C

import numpy as np

from . import _lib
from .api import Track


class BigWig:
    def __init__(self, path, stretch=10000):
        self.lib = _lib.load()
        self.h = self.lib.wb200_bw_open(str(path).encode())
        if not self.h:
            raise IOError(self.lib.wb200_bw_last_error(None).decode())
        self.lib.wb200_bw_set_stretch(self.h, int(stretch))
        n = self.lib.wb200_bw_nchrom(self.h)
        self.chroms = [(self.lib.wb200_bw_chrom_name(self.h, i).decode(), self.lib.wb200_bw_chrom_length(self.h, i)) for i in range(n)]

    def close(self):
        if self.h:
            self.lib.wb200_bw_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode(self, i):
        """(start, end, value) of chromosome i (strcmp order), 1-based half-open."""
        t = _lib.CTrack()
        rc = self.lib.wb200_bw_decode_chrom(self.h, i, C.addressof(t))
        if rc != 0:
            raise IOError("wb200_bw_decode_chrom: %d %s" % (rc, self.lib.wb200_bw_last_error(self.h).decode()))
        n = t.n
        if not n:
            return np.empty(0, np.int32), np.empty(0, np.int32), np.empty(0, np.float64)
        s = np.ctypeslib.as_array(C.cast(t.start, C.POINTER(C.c_int32)), (n,)).copy()
        e = np.ctypeslib.as_array(C.cast(t.end, C.POINTER(C.c_int32)), (n,)).copy()
        v = np.ctypeslib.as_array(C.cast(t.value, C.POINTER(C.c_double)), (n,)).copy()
        return s, e, v

    def track(self, names):
        """Whole-file Track; chromosome indices refer to `names`."""
        cs, ss, es, vs = [], [], [], []
        for i, (name, _) in enumerate(self.chroms):
            s, e, v = self.decode(i)
            cs.append(np.full(len(s), names.index(name), np.int32)); ss.append(s); es.append(e); vs.append(v)
        return Track(np.concatenate(cs), np.concatenate(ss), np.concatenate(es), np.concatenate(vs))
